"""Stage-by-stage cost of corr_fwd_kernel at the ShadowHand shape (BSIG_CORR_DBG bit mask:
1 no x math, 2 no W split, 4 no MMA, 8 no W copy); results are wrong by construction."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from bayes_sim_ig_b200 import _lib  # noqa: E402

dev = 'cuda:0'
s, q, m, n_out = 1050, 100, 100, 128
f = s * q + 2
ldf = (s + q + 2 + 3) // 4 * 4
fac = torch.randn(1000, ldf, device=dev)
rows = torch.randint(0, 1000, (m,), device=dev)
w = torch.randn(n_out, f, device=dev) / 300
b = torch.zeros(n_out, device=dev)
y = torch.empty(m, n_out, device=dev)
lib = _lib.load()
ws = torch.empty(lib.bsig_corr_linear_ws_bytes(m, n_out, s, q) + 256, dtype=torch.uint8, device=dev)
flush = torch.zeros(64 * 1024 * 1024, device=dev)
for dbg in (0,):
    os.environ['BSIG_CORR_DBG'] = str(dbg)
    ts = []
    for rep in range(6):
        flush.add_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call('bsig_corr_linear_fwd', fac.data_ptr(), ldf, rows.data_ptr(), s, q, w.data_ptr(),
                  b.data_ptr(), y.data_ptr(), m, n_out, 1, ws.data_ptr(), ws.numel(),
                  _lib.stream_ptr(dev))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    os.environ['BSIG_CORR_PROF'] = '1'
    _lib.call('bsig_corr_linear_fwd', fac.data_ptr(), ldf, rows.data_ptr(), s, q, w.data_ptr(),
              b.data_ptr(), y.data_ptr(), m, n_out, 1, ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
    del os.environ['BSIG_CORR_PROF']
    print('dbg=%2d  fwd+reduce %.1f us (median of 6, L2 flushed)' % (dbg, sorted(ts)[3]))

# ---- weight gradient + Adam epilogue at the same shape
dy = torch.randn(m, n_out, device=dev)
ea, es = torch.zeros_like(w), torch.zeros_like(w)
for dbg in (0,):
    ts = []
    for rep in range(7):
        flush.add_(1.0)
        if rep == 6:
            os.environ['BSIG_CORR_PROF'] = '1'
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call('bsig_corr_linear_wgrad', dy.data_ptr(), fac.data_ptr(), ldf, rows.data_ptr(), s, q,
                  m, n_out, None, w.data_ptr(), ea.data_ptr(), es.data_ptr(), rep + 1, 1e-4, 0.9,
                  0.999, 1e-8, 1.0, _lib.stream_ptr(dev))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print('dbg=%2d wgrad+adam %.1f us (median of 6, L2 flushed)' % (dbg, sorted(ts)[3]))
