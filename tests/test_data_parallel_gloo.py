"""World-size-2 gloo checks of the data-parallel host logic (CPU only): row
sharding, and that averaging per-rank gradients of equal shards reproduces the
gradient of the concatenated batch (the exchange step of SURVEY 8.e)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


def test_shard_rows_partition():
    from bayes_sim_ig_b200.data_parallel import shard_rows
    for n, world in ((4096, 8), (1000, 3), (7, 8), (1 << 20, 4)):
        spans = [shard_rows(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (a0, a1), (b0, b1) in zip(spans[:-1], spans[1:]):
            assert a1 == b0 and a1 - a0 >= b1 - b0 >= 0
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from bayes_sim_ig_b200.data_parallel import shard_rows
    from oracle import mdn_np
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'mdn.npz'))
    case = 'diag'
    p, k = 3, 4
    params = {key[len(case + '.init.'):]: g[key].astype(np.float64)
              for key in g.files if key.startswith(case + '.init.')}
    x, y = g[case + '.x'][:8], g[case + '.y'][:8]
    noise = np.full((8, p, k), 0.5)   # constant noise: the eps term is shard-independent
    lo, hi = shard_rows(8, rank, world)
    # local gradient on this rank's shard; the eps mean is local (documented)
    _, grads = mdn_np.mdnn_loss_and_grads(params, x[lo:hi], y[lo:hi], noise[lo:hi], p, k)
    flat = torch.from_numpy(np.concatenate([grads[key].ravel() for key in sorted(grads)]))
    dist.all_reduce(flat)             # the one exchange step
    flat /= world
    _, full = mdn_np.mdnn_loss_and_grads(params, x, y, noise, p, k)
    ref = np.concatenate([full[key].ravel() for key in sorted(full)])
    err = np.abs(flat.numpy() - ref).max() / np.abs(ref).max()
    torch.save(torch.tensor(err), os.path.join(tmpdir, 'err%d.pt' % rank))
    dist.destroy_process_group()


def test_two_rank_gradient_average_equals_full_batch(tmp_path):
    world = 2
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        err = float(torch.load(os.path.join(str(tmp_path), 'err%d.pt' % r)))
        # the only difference is the local-vs-global eps mean: O(1e-5) relative
        assert err < 5e-5, err


def test_fused_exchange_is_only_chosen_for_small_models_on_one_node(monkeypatch):
    """Gating of the peer-memory exchange (no GPU needed: the communicator itself is
    never built here): single rank, more than 8 ranks, models above P2P_MAX_PARAMS and
    BSIG_DP_EXCHANGE=nccl all go through NCCL."""
    from bayes_sim_ig_b200 import data_parallel

    class FakeModel(object):
        def __init__(self, world, n_params):
            self._dp_world = world
            self.flat_params = torch.zeros(n_params)

    built = []
    monkeypatch.setattr(data_parallel, 'P2PComm',
                        lambda n, group=None: built.append(n) or type('C', (), {'n_floats': n})())
    monkeypatch.delenv('BSIG_DP_EXCHANGE', raising=False)
    assert data_parallel.p2p_comm_for(FakeModel(1, 1000)) is None
    assert data_parallel.p2p_comm_for(FakeModel(16, 1000)) is None
    assert data_parallel.p2p_comm_for(FakeModel(8, data_parallel.P2P_MAX_PARAMS + 1)) is None
    m = FakeModel(8, 90126)
    comm = data_parallel.p2p_comm_for(m)
    assert comm is not None and built == [90126]
    assert data_parallel.p2p_comm_for(m) is comm and built == [90126]      # cached per model
    monkeypatch.setenv('BSIG_DP_EXCHANGE', 'nccl')
    assert data_parallel.p2p_comm_for(FakeModel(8, 90126)) is None


def _simulate_exchange(world, calls, n_updates, seed):
    """Discrete-event model of csrc/p2p.cu + train_engine: per update a rank (1) writes its
    gradient buffer [update parity], (2) publishes its epoch to every peer, (3) waits until
    all peers' epochs have arrived, (4) reads every rank's buffer [parity].  Ranks advance in
    random order.  Returns the number of protocol violations: a buffer written while a peer
    still has to read its previous content, or read with the wrong content."""
    import random
    rnd = random.Random(seed)
    epoch = [0] * world                                   # per-rank device epoch counter
    flags = [[0] * world for _ in range(world)]           # flags[r][q]: epoch published by q at r
    content = [[None, None] for _ in range(world)]        # content[r][parity] = (call, update)
    pending_reads = [[set(), set()] for _ in range(world)]  # readers that still need content
    state = [dict(call=0, upd=0, phase=0) for _ in range(world)]
    violations = 0
    while any(s['call'] < calls for s in state):
        r = rnd.choice([q for q in range(world) if state[q]['call'] < calls])
        s = state[r]
        par = s['upd'] & 1
        if s['phase'] == 0:                               # backward: write own buffer
            if pending_reads[r][par]:
                violations += 1                           # a peer has not read the old data yet
            content[r][par] = (s['call'], s['upd'])
            pending_reads[r][par] = set(range(world))
            s['phase'] = 1
        elif s['phase'] == 1:                             # publish
            epoch[r] += 1
            for q in range(world):
                flags[q][r] = epoch[r]
            s['phase'] = 2
        elif s['phase'] == 2:                             # wait for the peers, then read
            if all(flags[r][q] >= epoch[r] for q in range(world)):
                for q in range(world):
                    if content[q][par] != (s['call'], s['upd']):
                        violations += 1                   # read a buffer of another update
                    pending_reads[q][par].discard(r)
                s['phase'] = 3
        else:                                             # next update / next call
            s['upd'] += 1
            s['phase'] = 0
            if s['upd'] == n_updates:                     # next call starts at update 0 again
                s['upd'] = 0
                s['call'] += 1
    return violations


def test_double_buffered_exchange_protocol_has_no_hazard_for_even_update_counts():
    """The fused exchange needs no host barrier between calls when n_updates is even (the
    reference's 100 and 500): exhaustive-ish random interleavings of the protocol model."""
    for world in (2, 4, 8):
        for seed in range(25):
            assert _simulate_exchange(world, calls=3, n_updates=4, seed=seed) == 0


def test_exchange_protocol_model_detects_the_odd_count_hazard():
    """With an odd number of updates the next call starts on the buffer the last exchange may
    still be reading: the model finds violations without the barrier (which is why
    train_engine inserts one in that case)."""
    bad = sum(_simulate_exchange(4, calls=3, n_updates=3, seed=s)
              for s in range(40))
    assert bad > 0
