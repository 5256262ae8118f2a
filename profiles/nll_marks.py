"""%globaltimer marks of the minibatch NLL kernel inside a graph of dependent launches
(instrumented build: BSIG_NVCC_EXTRA=-DBSIG_NLL_PROF python -m bayes_sim_ig_b200.build)."""
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
p, k = 13, 10
rows = torch.randint(0, 800, (100,), device=dev)
noise = torch.rand(100, p, k, device=dev)
y = torch.rand(800, p, device=dev)
loss = torch.zeros(1, device=dev)
flag = torch.zeros(1, dtype=torch.int32, device=dev)
wsm = torch.zeros(lib.bsig_mdn_ws_bytes(100), dtype=torch.uint8, device=dev)
zz = torch.randn(100, 270, device=dev) * 0.3
dz = torch.empty(100, 270, device=dev)
fn = lambda: _lib.call(
    'bsig_mdn_nll_fused', zz.data_ptr(), noise.data_ptr(), y.data_ptr(), rows.data_ptr(),
    loss.data_ptr(), dz.data_ptr(), 100, p, k, 0, wsm.data_ptr(), wsm.numel(), flag.data_ptr(),
    _lib.stream_ptr(dev))
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
buf = np.zeros(8 * 16, dtype=np.uint64)
assert raw.dbg_nll_prof_read(buf.ctypes.data_as(ctypes.c_void_p)) == 0
m = buf.reshape(8, 16).astype(np.int64)
m = m[m[:, 0] > 0][:, :12]
t0 = m[:, 1].min()
labels = ['start', 'pdl', 'loads+exp', 'block_sum', 'pushed+arrive', 'softmax', 'barrier A', 'density+lse',
          'backward', 'block_sum2', 'barrier B', 'end']
print('CTAs', len(m))
for r in range(len(m)):
    print('rank %d:' % r, ' '.join('%6d' % (v - t0) for v in m[r]))
print('median phase lengths (ns):')
for i in range(1, 12):
    print('  %-14s %6d' % (labels[i], np.median(m[:, i] - m[:, i - 1])))
