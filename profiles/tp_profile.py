"""Per-phase cycle breakdown of the persistent training kernel at the bench shape
(BSIG_TP_PROF=1: thread 0 of CTA 0 accumulates clock64 deltas per phase)."""
import os
import sys
import time

import numpy as np
import torch

os.environ['BSIG_TP_PROF'] = '1'
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
import contextlib, io
with contextlib.redirect_stdout(io.StringIO()):
    model.run_training(x, y, 100, b, 0.2)
    plan = list(model._plans.values())[0]
    plan.tp_prof.zero_()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        model.run_training(x, y, 100, b, 0.2)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
prof = plan.tp_prof.cpu().numpy().astype(np.float64) / (reps * 100)
names = {31: 'loop overhead', 0: 'wait x (F0)', 1: 'F0 compute', 2: 'F0 exchange+load', 3: 'F1 compute',
         4: 'F1 exchange+load', 5: 'head fwd compute', 6: 'head exchange', 7: 'NLL loads', 8: 'NLL compute+store',
         9: 'NLL exchange + own dz', 10: 'dgrad head', 11: 'wait/none', 12: 'wgrad head + Adam',
         13: 'reload h0 + wait + collect', 14: 'dgrad L1', 15: 'none', 16: 'wgrad L1 + Adam',
         17: 'issue x + wait + collect', 18: '-', 19: 'wait x (wgrad0)', 20: 'wgrad L0 + Adam', 21: '-'}
tot = prof.sum()
for i in sorted(names, key=lambda i: (i == 31, i)):
    if prof[i] > 0:
        print('%-32s %9.0f cycles/update  %5.1f %%' % (names[i], prof[i], 100 * prof[i] / tot))
print('total %.0f cycles/update = %.2f us at 1.965 GHz; run_training wall %.3f ms' % (tot, tot / 1965.0, dt * 1e3))
