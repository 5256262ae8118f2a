"""2-rank smoke test of the data-parallel path with stage prints (debug aid)."""
import contextlib
import io
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import bench  # noqa: E402


def say(*a):
    print('[rank %s %.1fs]' % (os.environ.get('RANK'), time.time() - T0), *a, flush=True)


T0 = time.time()
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
local = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
say('init done')
t = torch.ones(4, device=dev)
dist.all_reduce(t)
torch.cuda.synchronize()
say('allreduce ok', t.tolist())
from bayes_sim_ig.bayes_sim import BayesSim  # noqa: E402
from bayes_sim_ig_b200 import data_parallel  # noqa: E402
states, actions, params, lows, highs = bench.synth(1000 + rank, 1000, bench.TASK)
cfg = {'modelClass': 'MDNN', 'summarizerFxn': 'summary_corrdiff', 'trainTrajLen': 20,
       'components': 10, 'hiddenLayers': [128, 128], 'lr': 1e-4}
torch.manual_seed(rank)
bsim = BayesSim(cfg, 4, 1, 13, lows, highs, prior=None, proposal=None, device=str(dev))
data_parallel.enable(bsim.model)
torch.cuda.synchronize()
say('enable ok; param checksum', float(bsim.model.flat_params.sum()))
with contextlib.redirect_stdout(io.StringIO()):
    for i in range(3):
        t1 = time.time()
        logs = bsim.run_training(params.to(dev), states.to(dev), actions.to(dev))
        torch.cuda.synchronize()
        sys.stderr.write('[rank %d] run_training %d: %.1f ms, last losses %s\n'
                         % (rank, i, 1e3 * (time.time() - t1), logs['train_loss'][-1]))
plan = list(bsim.model._plans.values())[0]
if plan.p2p is not None:
    st = plan.p2p.last_launch_stamps()
    say('exchange kernel (last launch): publish %.2f us, wait peers %.2f us, reduce+adam %.2f us'
        % ((st[1] - st[0]) / 1e3, (st[2] - st[1]) / 1e3, (st[3] - st[2]) / 1e3))
say('param checksum after', float(bsim.model.flat_params.double().sum()))
dist.barrier()
say('done')
dist.destroy_process_group()
