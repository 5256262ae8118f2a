"""Launch each streaming kernel a few times at its roofline size (for ncu):

    ncu --set full --clock-control none --import-source on -k regex:<name> -c N \
        -o gpurun_out/prof python profiles/profile_kernels.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import bench  # noqa: E402

if __name__ == '__main__':
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    buf = torch.zeros(64 * 1024 * 1024, device=dev)
    orig = bench.time_kernel
    bench.time_kernel = lambda fn, flush, reps=2, warm=1: orig(fn, flush, reps=reps, warm=warm)
    peak, src = bench.measured_peaks()
    res = bench.kernel_rooflines(dev, lambda: buf.add_(1.0), peak, src)
    for name, r in res.items():
        print(name, r['achieved'], 'GB/s', r['frac'])
