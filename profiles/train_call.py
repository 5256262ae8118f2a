"""A few BayesSim.run_training calls at the bench shape (for ncu launch lists)."""
import contextlib
import io
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import bench  # noqa: E402
from bayes_sim_ig.bayes_sim import BayesSim  # noqa: E402

dev = torch.device('cuda', 0)
n_calls = int(sys.argv[1]) if len(sys.argv) > 1 else 2
states_h, actions_h, params_h, lows, highs = bench.synth(1000, 1000, bench.TASK)
s, a, p = states_h.to(dev), actions_h.to(dev), params_h.to(dev)
cfg = {'modelClass': 'MDNN', 'summarizerFxn': bench.SUMMARIZER, 'trainTrajLen': 20,
       'components': 10, 'hiddenLayers': [128, 128], 'lr': 1e-4}
bsim = BayesSim(cfg, 4, 1, 13, lows, highs, prior=None, proposal=None, device='cuda:0')
with contextlib.redirect_stdout(io.StringIO()):
    for _ in range(n_calls):
        logs = bsim.run_training(p, s, a)
torch.cuda.synchronize()
print(logs)
