"""Wall-clock breakdown of one bench step (host vs device time)."""
import contextlib
import io
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import bench  # noqa: E402
from bayes_sim_ig.bayes_sim import BayesSim  # noqa: E402
from bayes_sim_ig_b200.models import train_engine  # noqa: E402

dev = torch.device('cuda', 0)
states_h, actions_h, params_h, lows, highs = bench.synth(1000, bench.N_TRAJ, bench.TASK, pin=True)
s, a, p = states_h.to(dev), actions_h.to(dev), params_h.to(dev)
cfg = {'modelClass': 'MDNN', 'summarizerFxn': bench.SUMMARIZER, 'trainTrajLen': 20,
       'components': 10, 'hiddenLayers': [128, 128], 'lr': 1e-4}
bsim = BayesSim(cfg, 4, 1, 13, lows, highs, prior=None, proposal=None, device='cuda:0')
sink = io.StringIO()


def sync():
    torch.cuda.synchronize()


def t(fn, n=5):
    fn(); sync()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    sync()
    return (time.perf_counter() - t0) / n * 1e3


with contextlib.redirect_stdout(sink):
    feats = bsim.summarizer_fxn(s[:1000], a[:1000])
    print_ms = {}
    print_ms['summarizer(1000)'] = t(lambda: bsim.summarizer_fxn(s[:1000], a[:1000]))
    print_ms['model.run_training(1000)'] = t(lambda: bsim.model.run_training(feats, p[:1000], 100, 100, 0.2))
    plan = list(bsim.model._plans.values())[0]
    print_ms['graph.replay only'] = t(lambda: plan.graph.replay())
    print_ms['randint x100'] = t(lambda: np.stack([np.random.randint(0, 800, 100) for _ in range(100)]))
    print_ms['noise uniform_'] = t(lambda: (plan.noise_train.uniform_(0, 1), plan.noise_test.uniform_(0, 1)))
    print_ms['stage_inputs'] = t(lambda: train_engine._stage_inputs(plan, bsim.model, feats, p[:1000]))
    print_ms['predict R=1'] = t(lambda: bsim.predict(s[:1], a[:1]))
    post = bsim.predict(s[:1], a[:1])
    print_ms['gen 10000'] = t(lambda: post.gen(10000))
    print_ms['eval 10000'] = t(lambda: post.eval(np.zeros((10000, 13))))
    print_ms['bsim.run_training(1000)'] = t(lambda: bsim.run_training(p[:1000], s[:1000], a[:1000]))
for k, v in print_ms.items():
    print('%-32s %8.3f ms' % (k, v))
print('launches per replay', plan.launches_per_replay)
