"""Where the RFF projection (65536 x 680 -> 2 x 100, tcgen05) spends its time: BSIG_TC_DBG bit mask
1 no B split, 2 no A conversion loads, 4 no sincos; engines 1 (TF32) and 2 (TF32x3)."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from bayes_sim_ig_b200 import _lib  # noqa: E402

dev = 'cuda:0'
m, d, nf = 65536, 680, 100
x = torch.randn(m, d, device=dev)
coeff = torch.randn(nf, d, device=dev) / 26
out = torch.empty(m, 2 * nf, device=dev)
lib = _lib.load()
ws = torch.empty(lib.bsig_linear_ws_bytes(m, nf, d) + 256, dtype=torch.uint8, device=dev)
flush = torch.zeros(64 * 1024 * 1024, device=dev)
for eng in (1, 2):
    for dbg in (0, 1, 2, 3, 4, 7):
        os.environ['BSIG_TC_DBG'] = str(dbg)
        ts = []
        for rep in range(8):
            flush.add_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.call('bsig_rff_features', x.data_ptr(), d, None, coeff.data_ptr(), out.data_ptr(), m, d,
                      nf, 0.1, eng, ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        print('engine %d dbg=%d  %.1f us' % (eng, dbg, sorted(ts)[4]))
