"""Launch the large-batch fused NLL (B=262144, P=13, K=10, diag) a few times (for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from bayes_sim_ig_b200 import _lib  # noqa: E402

dev = 'cuda:0'
b, p, k = 262144, 13, 10
full = int(os.environ.get('NLL_FULL', '0'))
nh = k * (1 + 2 * p + (p * (p - 1) // 2 if full else 0))
g = torch.Generator(dev).manual_seed(0)
z = 0.4 * torch.randn(b, nh, device=dev, generator=g)
noise = torch.rand(b, p, k, device=dev, generator=g)
y = torch.rand(b, p, device=dev, generator=g)
dz = torch.empty_like(z)
loss = torch.zeros(1, device=dev)
flag = torch.zeros(1, dtype=torch.int32, device=dev)
ws = torch.zeros(_lib.load().bsig_mdn_ws_bytes(b), dtype=torch.uint8, device=dev)
for _ in range(3):
    _lib.call('bsig_mdn_nll_fused', z.data_ptr(), noise.data_ptr(), y.data_ptr(), None,
              loss.data_ptr(), dz.data_ptr(), b, p, k, full, ws.data_ptr(), ws.numel(),
              flag.data_ptr(), _lib.stream_ptr(dev))
torch.cuda.synchronize()
print(loss.item(), int(flag.item()))
