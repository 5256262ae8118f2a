"""One ShadowHand-shaped run_training call (F = 105 002, 13.5 M parameters) for ncu launch lists:
    python profiles/shadowhand_step.py [n_updates]"""
import contextlib
import io
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import bench  # noqa: E402
from bayes_sim_ig.bayes_sim import BayesSim  # noqa: E402

n_updates = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device('cuda', 0)
task = dict(name='shadowhand', D=211, A=20, T1=51, P=32, K=10)
n = 1000
states, actions, params, lows, highs = bench.synth(11, n, task)
states, actions, params = states.to(dev), actions.to(dev), params.to(dev)
cfg = {'modelClass': 'MDNN', 'summarizerFxn': 'summary_corrdiff', 'trainTrajLen': 50,
       'components': 10, 'hiddenLayers': [128, 128], 'lr': 1e-4}
with contextlib.redirect_stdout(io.StringIO()):
    bsim = BayesSim(cfg, task['D'], task['A'], task['P'], lows, highs, prior=None, proposal=None,
                    device=str(dev))
    feats = bsim._training_summaries(states, actions)     # factored unless BSIG_FUSED_CORR=0
    for _ in range(2):
        logs = bsim.model.run_training(feats, params, n_updates, 100, 0.2)
torch.cuda.synchronize()
print('done', logs['test_loss'][-1])
