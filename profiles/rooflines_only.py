"""Event-timed rooflines of the streaming kernels (same code path as bench.py, 20 reps)."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import bench  # noqa: E402

dev = torch.device('cuda', 0)
torch.cuda.set_device(dev)
buf = torch.zeros(64 * 1024 * 1024, device=dev)
peak, src = bench.measured_peaks()
for name, r in bench.kernel_rooflines(dev, lambda: buf.add_(1.0), peak, src).items():
    print('%-44s %8.1f GB/s  %.4f  %.4f ms' % (name, r['achieved'], r['frac'], r['ms']))
