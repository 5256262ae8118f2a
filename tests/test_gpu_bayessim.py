"""GPU parity: BayesSim.run_training / predict end to end on a slice of the
reference's own Pendulum fixture, replaying every random draw it consumed."""
import numpy as np
import pytest
import torch

from helpers import injected_rand_like, load_state, rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _split_noise(g, name, n_updates, logs):
    n = int(g[name + '.train.n_noise'])
    draws = [g[name + '.train.noise%02d' % i] for i in range(n)]
    train, test = [], []
    it = iter(draws)
    for e in range(n_updates):
        train.append(next(it))
        if e in logs:
            test.append(next(it))
    return np.stack(train), np.stack(test)


@pytest.mark.parametrize('name,model_class,summ', [('mdnn_start', 'MDNN', 'summary_start'),
                                                   ('mdrff_corrdiff', 'MDRFF', 'summary_corrdiff')])
def test_bayessim_training_and_predict_match_reference(golden, name, model_class, summ):
    from bayes_sim_ig.bayes_sim import BayesSim
    from bayes_sim_ig_b200.models.train_engine import log_steps, run_training_captured
    g = golden('bayessim')
    cfg = {'modelClass': model_class, 'summarizerFxn': summ, 'trainTrajLen': 10,
           'components': 10, 'hiddenLayers': (24, 24), 'lr': 5e-4}
    lows, highs = np.array([0.01] * 2), np.array([2.0] * 2)
    bsim = BayesSim(model_cfg=cfg, obs_dim=3, act_dim=1, params_dim=2, params_lows=lows,
                    params_highs=highs, prior=None, proposal=None, device=DEV)
    load_state(bsim.model, g.sub(name + '.init.'))
    if model_class == 'MDRFF':
        bsim.model.rff.freqs = torch.from_numpy(g[name + '.rff.freqs']).to(DEV)
        bsim.model.rff.sigma = torch.from_numpy(g[name + '.rff.sigma']).to(DEV)
    data = torch.from_numpy(g['pendulum.data']).reshape(-1, 10, 4)
    st, ac = data[:, :, :3].to(DEV), data[:, :, 3:].to(DEV)
    prm = torch.from_numpy(g['pendulum.params']).to(DEV)
    n_updates, batch = 20, 32
    logs = log_steps(n_updates)
    noise_train, noise_test = _split_noise(g, name, n_updates, logs)
    summaries = bsim.summarizer_fxn(st, ac)
    out = run_training_captured(bsim.model, summaries, prm, n_updates, batch, BayesSim.TEST_FRACTION,
                                injected=dict(idx=g[name + '.train.idx'], noise_train=noise_train,
                                              noise_test=noise_test))
    # 20 Adam steps at lr 5e-4: losses to 1e-4 relative, weights to 2e-5 absolute
    np.testing.assert_allclose(out['train_loss'], g[name + '.train.train_loss'], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out['test_loss'], g[name + '.train.test_loss'], rtol=1e-4, atol=1e-5)
    for key, ref in g.sub(name + '.after.').items():
        got = bsim.model.state_dict()[key].cpu().numpy()
        assert np.abs(got - ref).max() <= 2e-5, key
    # predict (R = 1): posterior parameters and the NLL of the true parameters
    load_state(bsim.model, g.sub(name + '.after.'))
    tdata = torch.from_numpy(g['pendulum.true_data']).reshape(1, 10, 4)
    with injected_rand_like([g[name + '.predict.noise']], DEV):
        post = bsim.predict(tdata[:, :, :3].to(DEV), tdata[:, :, 3:].to(DEV))
    assert rel_err(post.a, g[name + '.predict.a']) < 1e-5
    assert rel_err(np.stack([c.m for c in post.xs]), g[name + '.predict.m']) < 1e-5
    assert rel_err(np.stack([c.S for c in post.xs]), g[name + '.predict.S']) < 5e-5
    nll = -post.eval(g['pendulum.true_params'])
    np.testing.assert_allclose(nll, g[name + '.predict.nll_true'], rtol=1e-4, atol=1e-4)


def test_public_run_training_and_multi_trajectory_predict(golden, capsys):
    """The public API with its own random draws: losses decrease on the
    reference's Pendulum data and predict() with R = 2 refits one mixture."""
    from bayes_sim_ig.bayes_sim import BayesSim
    g = golden('bayessim')
    torch.manual_seed(2)
    np.random.seed(2)
    cfg = {'modelClass': 'MDNN', 'summarizerFxn': 'summary_start', 'trainTrajLen': 10,
           'components': 4, 'hiddenLayers': (24, 24), 'lr': 5e-4, 'fullCovariance': True}
    lows, highs = np.array([0.01] * 2), np.array([2.0] * 2)
    bsim = BayesSim(model_cfg=cfg, obs_dim=3, act_dim=1, params_dim=2, params_lows=lows,
                    params_highs=highs, prior=None, proposal=None, device=DEV)
    data = torch.from_numpy(g['pendulum.data']).reshape(-1, 10, 4)
    st, ac = data[:, :, :3], data[:, :, 3:]          # host tensors are accepted (moved once)
    prm = torch.from_numpy(g['pendulum.params'])
    first = bsim.run_training(prm, st, ac)
    for _ in range(4):
        last = bsim.run_training(prm, st, ac)
    assert len(first['train_loss']) == 6 and len(first['test_loss']) == 6
    assert last['test_loss'][-1] < first['test_loss'][0]
    assert 'loss: train' in capsys.readouterr().out
    tdata = torch.from_numpy(g['pendulum.true_data']).reshape(1, 10, 4).repeat(2, 1, 1)
    post = bsim.predict(tdata[:, :, :3].to(DEV), tdata[:, :, 3:].to(DEV))
    assert post.n_components == 4 and post.ndim == 2
    assert np.isfinite(post.eval(g['pendulum.true_params'])).all()
    assert post.gen(5).shape == (5, 2)


def test_multi_trajectory_predict_matches_reference(golden, monkeypatch):
    """BayesSim.predict with R = 2 real trajectories (reference bayes_sim.py:148-179):
    resample the two per-trajectory mixtures, refit ONE unconditional MDNN with 500 Adam
    updates, return its mixture.  Replayed draw for draw against the live reference
    (tests/golden/predict_multi.npz, make_golden.py::golden_predict_multi): numpy is seeded
    identically, the 508 torch.rand_like draws come from the same per-call generator, the
    refit network starts from the recorded initial weights.

    Tolerances: the pooled samples are a float64 affine map cast to float32 (1e-6); the
    refit is 500 chained fp32 updates on two different machines, so its losses are held to
    1e-3 and the fitted mixture to 2e-2 of each quantity's scale (measured: see the assert
    messages; agreement degrades gracefully with the number of updates, not by a jump)."""
    import bayes_sim_ig_b200.bayes_sim as bs_mod
    import bayes_sim_ig_b200.models.train_engine as te
    from bayes_sim_ig.bayes_sim import BayesSim
    from helpers import seeded_uniforms
    g = golden('predict_multi')
    gb = golden('bayessim')
    lows, highs = np.array([0.01] * 2), np.array([2.0] * 2)
    cfg = {'modelClass': 'MDNN', 'summarizerFxn': 'summary_start', 'trainTrajLen': 10,
           'components': 10, 'hiddenLayers': (24, 24), 'lr': 5e-4}
    bsim = BayesSim(model_cfg=cfg, obs_dim=3, act_dim=1, params_dim=2, params_lows=lows,
                    params_highs=highs, prior=None, proposal=None, device=DEV)
    load_state(bsim.model, gb.sub('mdnn_start.after.'))
    n_updates = int(g['n_updates'])
    n_calls = int(g['n_rand_like_calls'])
    logs = te.log_steps(n_updates)
    assert n_calls == 2 + n_updates + len(logs)

    # the refit network starts from the reference's recorded initial weights
    captured = {}

    class ReplayMDNN(bs_mod.MDNN):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            load_state(self, g.sub('refit.init.'))

        def run_training(self, x_data, y_data, n_updates, batch_size, test_frac=0.2):
            captured['pool'] = y_data.detach().cpu().numpy().copy()
            captured['pool_device'] = y_data.device.type
            out = super().run_training(x_data, y_data, n_updates, batch_size, test_frac)
            captured['logs'] = out
            return out
    monkeypatch.setattr(bs_mod, 'MDNN', ReplayMDNN)

    # reference call order of torch.rand_like: 0 = predict_MoGs of the R rows; then per
    # update one training draw and, on logging steps, one held-out draw; last = the refit
    # network's predict_MoGs
    def hook(plan):
        p, k = plan.p, plan.k
        tr = torch.empty((plan.n_updates, plan.batch, p, k))
        te_ = torch.empty((len(plan.logs), max(plan.n_test, 1), p, k))
        call = 1
        for e in range(plan.n_updates):
            tr[e] = seeded_uniforms(call, (plan.batch, p, k))
            call += 1
            if e in plan.logs:
                te_[plan.logs.index(e)] = seeded_uniforms(call, (plan.n_test, p, k))
                call += 1
        assert call == n_calls - 1
        return tr, te_
    monkeypatch.setattr(te, 'NOISE_HOOK', hook)
    forward_calls = iter([0, n_calls - 1])

    def fake_rand_like(t, *a, **k):
        return seeded_uniforms(next(forward_calls), t.shape).to(device=t.device, dtype=t.dtype)
    monkeypatch.setattr(torch, 'rand_like', fake_rand_like)

    np.random.seed(77)
    states = torch.from_numpy(g['states']).to(DEV)
    actions = torch.from_numpy(g['actions']).to(DEV)
    post = bsim.predict(states, actions)

    pool = captured['pool']
    assert captured['pool_device'] == 'cuda'          # the pool never left the device
    assert list(pool.shape) == list(g['pool.shape'])
    np.testing.assert_allclose(pool[:16], g['pool.head'], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(pool[-16:], g['pool.tail'], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(pool.astype(np.float64).mean(axis=0), g['pool.mean'], rtol=1e-5)
    np.testing.assert_allclose(np.cov(pool.astype(np.float64).T), g['pool.cov'], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(captured['logs']['train_loss'], g['refit.train_loss'], rtol=1e-3)
    np.testing.assert_allclose(captured['logs']['test_loss'], g['refit.test_loss'], rtol=1e-3)
    assert post.n_components == 10 and post.ndim == 2
    got_m = np.stack([c.m for c in post.xs])
    got_s = np.stack([c.S for c in post.xs])
    assert np.abs(post.a - g['post.a']).max() <= 2e-2 * np.abs(g['post.a']).max(), \
        np.abs(post.a - g['post.a']).max()
    assert rel_err(got_m, g['post.m']) < 2e-2, rel_err(got_m, g['post.m'])
    assert rel_err(got_s, g['post.S']) < 2e-2, rel_err(got_s, g['post.S'])
    logp = post.eval(g['post.eval_x'].astype(np.float32))
    np.testing.assert_allclose(logp, g['post.eval_logp'], rtol=2e-2, atol=2e-2)
