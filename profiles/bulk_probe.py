"""What the 1-D cp.async.bulk pipeline sustains (GB/s, read + write bytes) for a range of
tile sizes / stages / CTAs per SM, next to torch's copy_ on the same buffers."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from bayes_sim_ig_b200 import _lib  # noqa: E402

dev = torch.device('cuda', 0)
n = 1 << 30                                     # 1 GiB each way
src = torch.empty(n, dtype=torch.uint8, device=dev).random_(0, 255)
dst = torch.empty_like(src)
flush = torch.zeros(64 << 20, device=dev)


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.add_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


ms = timed(lambda: dst.copy_(src))
print('torch copy_                         %7.1f GB/s' % (2 * n / ms / 1e6))
for tile in (4096, 16384, 32768, 65536):
    for stages in (2, 3, 4):
        for ctas in (1, 2, 3, 4):
            if stages * tile * ctas > 200 * 1024:
                continue
            ms = timed(lambda: _lib.call('bsig_bulk_copy_probe', src.data_ptr(), dst.data_ptr(), n,
                                         tile, stages, ctas, _lib.stream_ptr(dev)))
            print('bulk tile %6d stages %d ctas/SM %d  %7.1f GB/s' % (tile, stages, ctas,
                                                                      2 * n / ms / 1e6))
assert torch.equal(src, dst)
