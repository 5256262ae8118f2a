"""configs[3] data parallel (ShadowHand-shaped, 13.5 M parameters): ms per Adam update for the
NCCL exchange variants.  torchrun --nproc-per-node N profiles/shadowhand_dp.py"""
import contextlib
import io
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import bench  # noqa: E402
from bayes_sim_ig.bayes_sim import BayesSim  # noqa: E402
from bayes_sim_ig_b200 import data_parallel  # noqa: E402

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
torch.cuda.set_device(dev)
dist.init_process_group('nccl', device_id=dev)
task = dict(name='shadowhand', D=211, A=20, T1=51, P=32, K=10)
states, actions, params, lows, highs = bench.synth(2000 + rank, 1000, task)
states, actions, params = states.to(dev), actions.to(dev), params.to(dev)
cfg = {'modelClass': 'MDNN', 'summarizerFxn': 'summary_corrdiff', 'trainTrajLen': 50,
       'components': 10, 'hiddenLayers': [128, 128], 'lr': 1e-4}
for sharded, fused in (('1', '1'), ('0', '1'), ('1', '0')):
    os.environ['BSIG_DP_SHARDED'] = sharded
    os.environ['BSIG_FUSED_CORR'] = fused
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        bsim = BayesSim(cfg, task['D'], task['A'], task['P'], lows, highs, prior=None, proposal=None,
                        device=str(dev))
        data_parallel.enable(bsim.model)
        for _ in range(2):
            bsim.run_training(params, states, actions)
        ts = []
        for _ in range(3):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            logs = bsim.run_training(params, states, actions)
            e1.record()
            e1.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(float(t.item()))
    if rank == 0:
        print('world %d sharded=%s fused_corr=%s: %.3f ms per update (max over ranks, median of 3), '
              'final test loss %.4f' % (world, sharded, fused, sorted(ts)[1] / 100, logs['test_loss'][-1]),
              flush=True)
    del bsim
    torch.cuda.empty_cache()
dist.destroy_process_group()
