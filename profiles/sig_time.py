"""Event-timed depth-3 signature kernel at the bench roofline shape (C = 6, 1 M trajectories)."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import bench  # noqa: E402
from bayes_sim_ig_b200 import _lib  # noqa: E402

dev = torch.device('cuda', 0)
n, t1, d, a = 1 << 20, 21, 4, 1
s = torch.randn(n, t1, d, device=dev)
ac = torch.rand(n, t1, a, device=dev)
sig = torch.empty((n, 258), device=dev)
buf = torch.zeros(64 * 1024 * 1024, device=dev)
ms = bench.time_kernel(lambda: _lib.call('bsig_signature_fwd', s.data_ptr(), ac.data_ptr(), sig.data_ptr(),
                                         n, t1, t1, t1, d, a, 3, _lib.stream_ptr(dev)),
                       lambda: buf.add_(1.0))
by = n * 4 * (t1 * (d + a) + 258)
print('BSIG_SIG_TPB=%s: %.4f ms  %.1f GB/s  frac %.4f' % (os.environ.get('BSIG_SIG_TPB', 'default'), ms,
                                                       by / ms / 1e6, by / ms / 1e6 / 6554.2))
