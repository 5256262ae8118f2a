import os, sys, io, contextlib, traceback
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import torch, numpy as np
import bench
from bayes_sim_ig.bayes_sim import BayesSim
from bayes_sim_ig_b200.models.train_engine import run_training_captured
dev = torch.device('cuda', 0)
task = dict(name='shadowhand', D=211, A=20, T1=51, P=32, K=10)
n = 1000
states, actions, params, lows, highs = bench.synth(11, n, task)
states, actions, params = states.to(dev), actions.to(dev), params.to(dev)
cfg = {'modelClass': 'MDNN', 'summarizerFxn': 'summary_corrdiff', 'trainTrajLen': 50,
       'components': 10, 'hiddenLayers': [128, 128], 'lr': 1e-4}
bsim = BayesSim(cfg, task['D'], task['A'], task['P'], lows, highs, prior=None, proposal=None, device=str(dev))
feats = bsim.summarizer_fxn(states, actions)
print('feats', feats.shape)
for ug in (False, True):
    for pdl in (1, 0):
        from bayes_sim_ig_b200 import _lib
        _lib.load().bsig_set_pdl(pdl)
        bsim.model._plans = {}
        try:
            logs = run_training_captured(bsim.model, feats, params, 100, 100, 0.2, use_graph=ug)
            torch.cuda.synchronize()
            print('use_graph', ug, 'pdl', pdl, 'OK', logs['test_loss'][-1])
        except Exception as e:
            print('use_graph', ug, 'pdl', pdl, 'FAILED', repr(e)[:400])
            traceback.print_exc(limit=6)
            break
