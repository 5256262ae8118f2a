"""%globaltimer marks of every CTA of one minibatch GEMM launch inside a graph of dependent launches
(instrumented build: BSIG_NVCC_EXTRA=-DBSIG_GS_PROF python -m bayes_sim_ig_b200.build).
Marks: 0 start, 1 PDL wait returned, 2 operands staged (last pass), 3 CTA's tile reduced,
4 first cluster barrier passed, 5 remote tiles read (rank 0) / about to arrive (peers), 6 end."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from bayes_sim_ig_b200 import _lib  # noqa: E402
from bayes_sim_ig_b200.build import LIB_PATH  # noqa: E402

dev = torch.device('cuda', 0)
torch.cuda.set_device(dev)
lib = _lib.load()
raw = ctypes.CDLL(LIB_PATH)
x = torch.randn(800, 302, device=dev)
rows = torch.randint(0, 800, (100,), device=dev)
w1 = torch.randn(128, 302, device=dev)
b1 = torch.randn(128, device=dev)
h1 = torch.empty(100, 128, device=dev)
wh = torch.randn(270, 128, device=dev)
bh = torch.randn(270, device=dev)
z = torch.empty(100, 270, device=dev)
ws = torch.empty(1 << 22, dtype=torch.uint8, device=dev)
st = lambda: _lib.stream_ptr(dev)
cases = {
    'fwd L1 gather 100x128x302 (S=5, 16 tiles)': (5, lambda: _lib.call(
        'bsig_linear_fwd', x.data_ptr(), 302, rows.data_ptr(), w1.data_ptr(), b1.data_ptr(),
        h1.data_ptr(), 100, 128, 302, 1, 0, ws.data_ptr(), ws.numel(), st())),
    'fwd heads 100x270x128 (S=2, 36 tiles)': (2, lambda: _lib.call(
        'bsig_linear_fwd', h1.data_ptr(), 128, None, wh.data_ptr(), bh.data_ptr(), z.data_ptr(),
        100, 270, 128, 0, 0, ws.data_ptr(), ws.numel(), st())),
}
for name, (S, fn) in cases.items():
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(200):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    buf = np.zeros(256 * 8, dtype=np.uint64)
    assert raw.dbg_gs_prof_read(buf.ctypes.data_as(ctypes.c_void_p)) == 0
    m = buf.reshape(256, 8).astype(np.int64)
    n_cta = int((m[:, 0] > 0).sum())
    m = m[:n_cta, :7]
    t0 = m[:, 1].min()                      # earliest PDL-wait return of the launch
    print('==', name, 'CTAs', n_cta, '(ns after the earliest PDL-wait return of the launch)')
    for c in range(min(n_cta, 2 * S)):
        print('  tile %d rank %d:' % (c // S, c % S), ' '.join('%6d' % (v - t0) for v in m[c]))
    rel = m - t0
    for i, lab in enumerate(['start', 'pdl', 'staged', 'reduced', 'barrier1', 'read/arrive', 'end']):
        col = rel[:, i]
        if i in (5,):
            col = rel[::S, i]
        print('  %-12s min %6d  median %6d  max %6d' % (lab, col.min(), np.median(col), col.max()))
    r0 = rel[::S]
    print('  rank 0 phases (median ns): staged-pdl %d, reduced-staged %d, barrier1-reduced %d, read-barrier1 %d, end-read %d' % tuple(
        np.median(r0[:, b] - r0[:, a]) for a, b in [(1, 2), (2, 3), (3, 4), (4, 5), (5, 6)]))
    peers = np.delete(rel, np.s_[::S], axis=0)
    print('  peers: reduced (median/max) %d / %d ; rank0 reduced (median/max) %d / %d' % (
        np.median(peers[:, 3]), peers[:, 3].max(), np.median(r0[:, 3]), r0[:, 3].max()))
