"""One scaled-mode run_training call (Cartpole widths, minibatch B_g rows) for ncu launch lists:
    python profiles/scaled_step.py [B_g] [n_updates] [engine]"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from bayes_sim_ig.models.mdnn import MDNN  # noqa: E402

b_g = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
n_updates = int(sys.argv[2]) if len(sys.argv) > 2 else 6
engine = int(sys.argv[3]) if len(sys.argv) > 3 else -1
dev = 'cuda:0'
f, p, k, n = 302, 13, 10, 1 << 17
rs = np.random.RandomState(0)
x = torch.from_numpy(rs.randn(n, f).astype(np.float32)).to(dev)
y = torch.from_numpy((0.1 + 1.9 * rs.rand(n, p)).astype(np.float32)).to(dev)
torch.manual_seed(0)
model = MDNN(f, p, np.full(p, 0.1), np.full(p, 2.0), k, False, (128, 128), torch.nn.Tanh, 1e-4,
             device=dev)
model.gemm_engine = engine
with contextlib.redirect_stdout(io.StringIO()):
    for _ in range(2):
        model.run_training(x, y, n_updates, b_g, 0.2)
torch.cuda.synchronize()
print('done')
