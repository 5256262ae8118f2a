"""Per-kernel latency inside a CUDA graph for the reference-sized layer GEMMs
(graph of 200 back-to-back launches; time per launch = replay time / 200)."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from bayes_sim_ig_b200 import _lib  # noqa: E402

dev = torch.device('cuda', 0)
torch.cuda.set_device(dev)
lib = _lib.load()
N = 200


def graph_time(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(N):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / 5 / N * 1e3


x = torch.randn(800, 302, device=dev)
rows = torch.randint(0, 800, (100,), device=dev)
w1 = torch.randn(128, 302, device=dev)
b1 = torch.randn(128, device=dev)
h1 = torch.empty(100, 128, device=dev)
wh = torch.randn(270, 128, device=dev)
bh = torch.randn(270, device=dev)
z = torch.empty(100, 270, device=dev)
dz = torch.randn(100, 270, device=dev)
dwh = torch.empty(270, 128, device=dev)
dbh = torch.empty(270, device=dev)
dh = torch.empty(100, 128, device=dev)
dw1 = torch.empty(128, 302, device=dev)
db1 = torch.empty(128, device=dev)
ws = torch.empty(1 << 22, dtype=torch.uint8, device=dev)
st = lambda: _lib.stream_ptr(dev)

tests = {
    'fwd L1 gather 100x128x302 tanh': lambda: _lib.call(
        'bsig_linear_fwd', x.data_ptr(), 302, rows.data_ptr(), w1.data_ptr(), b1.data_ptr(),
        h1.data_ptr(), 100, 128, 302, 1, 0, ws.data_ptr(), ws.numel(), st()),
    'fwd heads 100x270x128': lambda: _lib.call(
        'bsig_linear_fwd', h1.data_ptr(), 128, None, wh.data_ptr(), bh.data_ptr(), z.data_ptr(),
        100, 270, 128, 0, 0, ws.data_ptr(), ws.numel(), st()),
    'wgrad heads 270x128x100 +db': lambda: _lib.call(
        'bsig_linear_wgrad', dz.data_ptr(), h1.data_ptr(), 128, None, dwh.data_ptr(),
        dbh.data_ptr(), 100, 270, 128, 0, ws.data_ptr(), ws.numel(), st()),
    'dgrad heads 100x128x270 dtanh': lambda: _lib.call(
        'bsig_linear_dgrad', dz.data_ptr(), wh.data_ptr(), h1.data_ptr(), dh.data_ptr(), 100, 270,
        128, 1, 0, ws.data_ptr(), ws.numel(), st()),
    'wgrad L1 gather 128x302x100 +db': lambda: _lib.call(
        'bsig_linear_wgrad', dh.data_ptr(), x.data_ptr(), 302, rows.data_ptr(), dw1.data_ptr(),
        db1.data_ptr(), 100, 128, 302, 0, ws.data_ptr(), ws.numel(), st()),
}
p, k = 13, 10
noise = torch.rand(100, p, k, device=dev)
y = torch.rand(800, p, device=dev)
loss = torch.zeros(1, device=dev)
flag = torch.zeros(1, dtype=torch.int32, device=dev)
wsm = torch.zeros(lib.bsig_mdn_ws_bytes(100), dtype=torch.uint8, device=dev)
zz = torch.randn(100, 270, device=dev) * 0.3
tests['nll fused fwd+bwd B=100'] = lambda: _lib.call(
    'bsig_mdn_nll_fused', zz.data_ptr(), noise.data_ptr(), y.data_ptr(), rows.data_ptr(),
    loss.data_ptr(), dz.data_ptr(), 100, p, k, 0, wsm.data_ptr(), wsm.numel(), flag.data_ptr(), st())
prm, g, m, v = (torch.zeros(90128, device=dev) for _ in range(4))
tests['adam 90k'] = lambda: _lib.call(
    'bsig_adam_step', prm.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), 90128, 1, 1e-4, 0.9,
    0.999, 1e-8, 1.0, st())
tiny = torch.zeros(32, device=dev)
tests['torch tiny add_ (launch floor)'] = lambda: tiny.add_(1.0)
print('BSIG_SMALL_GEMM_MAXSPLIT =', os.environ.get('BSIG_SMALL_GEMM_MAXSPLIT', '8 (default)'))
for name, fn in tests.items():
    print('%-36s %7.2f us/launch' % (name, graph_time(fn)))
