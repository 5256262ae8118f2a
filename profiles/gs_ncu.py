"""A few eager launches of the reference-sized first-layer forward GEMM (S = 5 cluster) for one
`ncu --set full` capture of gemm_small_kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from bayes_sim_ig_b200 import _lib  # noqa: E402

dev = torch.device('cuda', 0)
torch.cuda.set_device(dev)
lib = _lib.load()
x = torch.randn(800, 302, device=dev)
rows = torch.randint(0, 800, (100,), device=dev)
w1 = torch.randn(128, 302, device=dev)
b1 = torch.randn(128, device=dev)
h1 = torch.empty(100, 128, device=dev)
ws = torch.empty(1 << 22, dtype=torch.uint8, device=dev)
for _ in range(30):
    _lib.call('bsig_linear_fwd', x.data_ptr(), 302, rows.data_ptr(), w1.data_ptr(), b1.data_ptr(),
              h1.data_ptr(), 100, 128, 302, 1, 0, ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
torch.cuda.synchronize()
print('ok')
