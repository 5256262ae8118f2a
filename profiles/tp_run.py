"""Two reference-sized run_training calls at the bench shape (for ncu captures of the
persistent training kernel)."""
import contextlib
import io
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from bayes_sim_ig.models.mdnn import MDNN  # noqa: E402

dev = 'cuda:0'
f, p, k, b, n = 302, 13, 10, 100, 1000
rs = np.random.RandomState(0)
x = torch.from_numpy(rs.randn(n, f).astype(np.float32)).to(dev)
y = torch.from_numpy((0.1 + 1.9 * rs.rand(n, p)).astype(np.float32)).to(dev)
torch.manual_seed(0)
model = MDNN(f, p, np.full(p, 0.1), np.full(p, 2.0), k, False, (128, 128), torch.nn.Tanh, 1e-4,
             device=dev)
with contextlib.redirect_stdout(io.StringIO()):
    for _ in range(2):
        model.run_training(x, y, 100, b, 0.2)
torch.cuda.synchronize()
print('done')
