"""GPU parity of the fused cross-correlation -> first-layer kernels (SURVEY 8.f rank 1,
csrc/corr_layer.cu) through the C ABI, against the CPU oracle of the reference's summarizer
(oracle/summarizers_np.py <- utils/summarizers.py:90-130) followed by the float64 layer
arithmetic of models/mdnn.py:108 and its autograd / torch.optim.Adam (oracle/mdn_np.adam_step).

Tolerances (max-norm relative): the summary factors and the two statistics are bit-exact /
2e-6 like the materialised summary; the TF32x3 tensor-core products carry the tolerance of
tests/test_gpu_gemm.py (2e-5 + 4e-8*K: fp32-grade operands, truncating fp32 accumulator)."""
import numpy as np
import pytest
import torch

from helpers import synth_rollouts

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'

# (name, D, A, T1): window / widths follow the reference's rule (5 steps when D > 50)
TASKS = {
    'cartpole': (4, 1, 21),          # s=30   q=10   F=302
    'ant': (60, 8, 51),              # s=295  q=40   F=11802
    'halfcheetah': (17, 6, 31),      # s=160  q=60   F=9602
    'shadowhand': (211, 20, 51),     # s=1050 q=100  F=105002
}


def _lib():
    from bayes_sim_ig_b200 import _lib
    return _lib


def _factors(task, n, seed=0, use_diff=True):
    from bayes_sim_ig_b200.utils import summarizers as bs
    d, a, t1 = TASKS[task]
    states, actions = synth_rollouts(seed, n, t1, d, a)
    cf = bs.corr_factors(states.to(DEV), actions.to(DEV), use_state_diff=use_diff)
    return states, actions, cf


def _oracle_x(states, actions, use_diff=True):
    from oracle import summarizers_np as osum
    return np.asarray(osum.cross_correlation(states.numpy(), actions.numpy(), use_diff), np.float32)


def _rel(got, ref):
    ref = np.asarray(ref, np.float64)
    return float(np.abs(np.asarray(got, np.float64) - ref).max() / (np.abs(ref).max() + 1e-30))


@pytest.mark.parametrize('task', ['cartpole', 'ant', 'halfcheetah', 'shadowhand'])
@pytest.mark.parametrize('use_diff', [True, False])
def test_factors_reproduce_the_materialised_summary(task, use_diff):
    from bayes_sim_ig_b200.utils import summarizers as bs
    n = 37 if task != 'shadowhand' else 9
    states, actions, cf = _factors(task, n, seed=3, use_diff=use_diff)
    x_kernel = bs.cross_correlation(states.to(DEV), actions.to(DEV), use_state_diff=use_diff)
    assert cf.shape == tuple(x_kernel.shape)
    # the factored form expands to exactly what the summarizer kernel stores
    assert torch.equal(cf.materialize(), x_kernel)
    x_ref = _oracle_x(states, actions, use_diff)
    got = cf.materialize().cpu().numpy()
    assert np.array_equal(got[:, :-2], x_ref[:, :-2])          # product block bit-exact
    assert _rel(got[:, -2:], x_ref[:, -2:]) < 2e-6
    # time-major ingestion gives the same factors
    cf_tm = bs.corr_factors(states.to(DEV).transpose(0, 1).contiguous(),
                            actions.to(DEV).transpose(0, 1).contiguous(),
                            use_state_diff=use_diff, time_major=True)
    assert torch.equal(cf_tm.fac, cf.fac)


def test_factors_flag_non_finite_inputs_and_overflow():
    from bayes_sim_ig_b200.utils import summarizers as bs
    d, a, t1 = TASKS['cartpole']
    states, actions = synth_rollouts(0, 8, t1, d, a)
    bad = states.clone()
    bad[3, 2, 1] = float('nan')
    with pytest.raises(AssertionError):
        bs.corr_factors(bad.to(DEV), actions.to(DEV), use_state_diff=True)
    big_s, big_a = states.clone(), actions.clone()
    big_s[5, 1, 0] = 3e30
    big_a[5, 4, 0] = 1e10                 # finite factors, the product overflows fp32
    with pytest.raises(AssertionError):
        bs.corr_factors(big_s.to(DEV), big_a.to(DEV), use_state_diff=False)


FWD_CASES = [('cartpole', 100, 128, True), ('cartpole', 7, 20, False), ('ant', 100, 128, True),
             ('halfcheetah', 200, 64, False), ('shadowhand', 100, 128, True),
             ('shadowhand', 200, 128, False), ('ant', 300, 31, True)]


@pytest.mark.parametrize('task,m,n_out,gather', FWD_CASES)
def test_fused_forward_matches_oracle_summary_times_weight(task, m, n_out, gather):
    lib = _lib()
    n = m + 50 if gather else m
    states, actions, cf = _factors(task, n, seed=m + n_out)
    f = cf.shape[1]
    g = torch.Generator('cpu').manual_seed(f + m)
    w = (torch.randn(n_out, f, generator=g) / np.sqrt(f)).to(DEV)
    b = torch.randn(n_out, generator=g).to(DEV)
    rows = torch.randint(0, n, (m,), generator=g) if gather else None
    rows_dev = None if rows is None else rows.to(DEV)
    assert lib.load().bsig_corr_linear_applicable(min(m, 128), m, n_out, cf.s, cf.q) == 1
    ws = torch.empty(lib.load().bsig_corr_linear_ws_bytes(m, n_out, cf.s, cf.q) + 256,
                     dtype=torch.uint8, device=DEV)
    x_ref = _oracle_x(states, actions).astype(np.float64)
    if rows is not None:
        x_ref = x_ref[rows.numpy()]
    pre = x_ref @ w.double().cpu().numpy().T + b.double().cpu().numpy()
    y = torch.empty(m, n_out, device=DEV)
    for act in (0, 1):
        y.fill_(float('nan'))
        lib.call('bsig_corr_linear_fwd', cf.fac.data_ptr(), cf.fac.shape[1],
                 None if rows_dev is None else rows_dev.data_ptr(), cf.s, cf.q, w.data_ptr(),
                 b.data_ptr(), y.data_ptr(), m, n_out, act, ws.data_ptr(), ws.numel(),
                 lib.stream_ptr(DEV))
        ref = np.tanh(pre) if act else pre
        err = _rel(y.cpu().numpy(), ref)
        assert err < 2e-5 + 4e-8 * f, (task, m, n_out, act, err)


WG_CASES = [('cartpole', 100, 128, True), ('cartpole', 5, 20, False), ('ant', 100, 128, True),
            ('halfcheetah', 128, 64, False), ('shadowhand', 100, 128, True), ('ant', 33, 31, True)]


@pytest.mark.parametrize('task,m,n_out,gather', WG_CASES)
def test_fused_weight_gradient_matches_oracle(task, m, n_out, gather):
    lib = _lib()
    n = m + 20 if gather else m
    states, actions, cf = _factors(task, n, seed=2 * m + n_out)
    f = cf.shape[1]
    g = torch.Generator('cpu').manual_seed(f + 3 * m)
    dy = torch.randn(m, n_out, generator=g).to(DEV)
    rows = torch.randint(0, n, (m,), generator=g) if gather else None
    rows_dev = None if rows is None else rows.to(DEV)
    dw = torch.full((n_out, f), float('nan'), device=DEV)
    lib.call('bsig_corr_linear_wgrad', dy.data_ptr(), cf.fac.data_ptr(), cf.fac.shape[1],
             None if rows_dev is None else rows_dev.data_ptr(), cf.s, cf.q, m, n_out,
             dw.data_ptr(), None, None, None, 1, 0.0, 0.9, 0.999, 1e-8, 1.0, lib.stream_ptr(DEV))
    x_ref = _oracle_x(states, actions).astype(np.float64)
    if rows is not None:
        x_ref = x_ref[rows.numpy()]
    ref = dy.double().cpu().numpy().T @ x_ref
    err = _rel(dw.cpu().numpy(), ref)
    assert err < 2e-5 + 4e-8 * m, (task, m, n_out, err)


@pytest.mark.parametrize('task,m,n_out', [('cartpole', 100, 128), ('ant', 100, 128),
                                          ('shadowhand', 100, 128)])
def test_fused_adam_epilogue_matches_oracle_adam(task, m, n_out):
    """Three updates with the optimiser applied in the weight-gradient epilogue against the
    oracle's torch.optim.Adam restatement fed with float64 gradients."""
    from oracle import mdn_np
    lib = _lib()
    states, actions, cf = _factors(task, m + 10, seed=11)
    f = cf.shape[1]
    g = torch.Generator('cpu').manual_seed(f)
    w0 = (torch.randn(n_out, f, generator=g) / np.sqrt(f))
    w = w0.clone().to(DEV)
    ea, es = torch.zeros_like(w), torch.zeros_like(w)
    dw = torch.empty_like(w)
    x_all = _oracle_x(states, actions).astype(np.float64)
    params = {'w': w0.double().numpy().copy()}
    m_ref = {'w': np.zeros_like(params['w'])}
    v_ref = {'w': np.zeros_like(params['w'])}
    lr = 1e-3
    for step in range(1, 4):
        dy = torch.randn(m, n_out, generator=g)
        rows = torch.randint(0, m + 10, (m,), generator=g)
        dy_dev, rows_dev = dy.to(DEV), rows.to(DEV)
        lib.call('bsig_corr_linear_wgrad', dy_dev.data_ptr(), cf.fac.data_ptr(), cf.fac.shape[1],
                 rows_dev.data_ptr(), cf.s, cf.q, m, n_out, dw.data_ptr(), w.data_ptr(),
                 ea.data_ptr(), es.data_ptr(), step, lr, 0.9, 0.999, 1e-8, 1.0,
                 lib.stream_ptr(DEV))
        # (1) the gradient the epilogue consumed is the oracle's gradient ...
        g_ref = dy.double().numpy().T @ x_all[rows.numpy()]
        g_got = dw.cpu().numpy()
        assert _rel(g_got, g_ref) < 2e-5 + 4e-8 * m
        # (2) ... and the optimiser arithmetic on that gradient is torch.optim.Adam's.  (Feeding
        # the oracle its OWN float64 gradient instead would test something else: Adam divides
        # by sqrt(v), so a parameter whose gradient is far below the fp32 rounding of the sum
        # moves by +-lr with the sign of the rounding error -- in 13.4 M entries some do.)
        mdn_np.adam_step(params, {'w': g_got.astype(np.float64)}, m_ref, v_ref, step, lr)
    assert np.abs(w.cpu().numpy() - params['w']).max() < 1e-6
    assert _rel(ea.cpu().numpy(), m_ref['w']) < 1e-5
    # (1 - beta2 is formed in fp32 from the fp32 beta2 of the C ABI, as in bsig_adam_step:
    # 1.0f - 0.999f = 0.00100005, 4.7e-5 above torch's double 1 - 0.999; its effect on a
    # parameter is < 3e-8 per step)
    assert _rel(es.cpu().numpy(), v_ref['w']) < 1e-4


def test_shapes_outside_the_envelope_are_refused():
    lib = _lib()
    assert lib.load().bsig_corr_linear_applicable(100, 200, 128, 1050, 100) == 1
    assert lib.load().bsig_corr_linear_applicable(256, 256, 128, 1050, 100) == 0   # batch > 128
    assert lib.load().bsig_corr_linear_applicable(100, 100, 256, 1050, 100) == 0   # n_out > 128
    assert lib.load().bsig_corr_linear_applicable(100, 100, 128, 4096, 512) == 0   # F >= 2^20
    # Humanoid: 5*107 x 5*21 + 2 is odd -> rows of W only 4-byte aligned -> materialised path
    assert lib.load().bsig_corr_linear_applicable(100, 100, 128, 535, 105) == 0


def _oracle_training(sd0, x, y_norm, idx, noise_train, noise_test, n_train, p, k, lr):
    """models/mdnn.py:217-241 in float64 on the materialised oracle summary."""
    from oracle import mdn_np
    params = {name: np.asarray(v, np.float64).copy() for name, v in sd0.items()}
    m_ref = {name: np.zeros_like(v) for name, v in params.items()}
    v_ref = {name: np.zeros_like(v) for name, v in params.items()}
    x_tr, y_tr, x_te, y_te = x[:n_train], y_norm[:n_train], x[n_train:], y_norm[n_train:]
    train, test = [], []
    for step in range(idx.shape[0]):
        rows = idx[step]
        loss, grads = mdn_np.mdnn_loss_and_grads(params, x_tr[rows], y_tr[rows], noise_train[step],
                                                 p, k)
        mdn_np.adam_step(params, grads, m_ref, v_ref, step + 1, lr)
        w, mu, ld, low, _ = mdn_np.mdnn_forward(params, x_te, noise_test[step], p, k)
        train.append(float(loss))
        test.append(float(mdn_np.mdn_loss(w, mu, ld, low, y_te)))
    return params, train, test


@pytest.mark.parametrize('task,hidden', [('ant', (128, 64)), ('shadowhand', (128, 128)),
                                         ('halfcheetah', (40,))])
def test_training_on_factored_summaries_matches_oracle(task, hidden):
    """MDNN.run_training fed with CorrFactors (fused first layer: generated operand tiles,
    Adam in the weight-gradient epilogue) against the float64 oracle of the reference loop on
    the materialised summary: same minibatch rows and eps-noise, three updates, every loss.
    Parameters: Adam moves an entry whose gradient is below the rounding of the fp32 sum by
    +-lr with the sign of that rounding, so the first-layer weight is compared by the share of
    entries off by more than 2e-5 (< 0.1 %) and by the hard bound 2*lr per update; all other
    tensors by the 2e-5 max-norm bound of tests/test_gpu_mdn.py."""
    from bayes_sim_ig.models.mdnn import MDNN
    from bayes_sim_ig_b200.models.train_engine import run_training_captured
    n, b, n_up, p, k, lr = 140, 100, 3, 5, 4, 1e-3
    states, actions, cf = _factors(task, n, seed=21)
    f = cf.shape[1]
    rs = np.random.RandomState(5)
    lows, highs = np.full(p, -1.0), np.full(p, 3.0)
    y_raw = (lows + (highs - lows) * rs.rand(n, p)).astype(np.float32)
    torch.manual_seed(3)
    model = MDNN(f, p, lows, highs, k, False, hidden, torch.nn.Tanh, lr, device=DEV)
    sd0 = {name: v.detach().cpu().numpy().copy() for name, v in model.state_dict().items()}
    n_train = int(n * 0.8)
    n_test = n - n_train
    idx = rs.randint(0, n_train, (n_up, b))
    noise_train = rs.rand(n_up, b, p, k).astype(np.float32)
    noise_test = rs.rand(n_up, n_test, p, k).astype(np.float32)
    inj = dict(idx=idx, noise_train=noise_train, noise_test=noise_test)
    logs = run_training_captured(model, cf, torch.from_numpy(y_raw).to(DEV), n_up, b, 0.2,
                                 use_graph=True, injected=inj)
    plan = list(model._plans.values())[-1]
    assert plan.corr is not None and plan.corr_adam          # the fused path is what ran
    x = _oracle_x(states, actions).astype(np.float64)
    y_norm = (y_raw.astype(np.float64) - lows) / (highs - lows)
    params, train, test = _oracle_training(sd0, x, y_norm, idx, noise_train, noise_test, n_train,
                                           p, k, lr)
    np.testing.assert_allclose(logs['train_loss'][0], train[0], rtol=1e-5)
    np.testing.assert_allclose(logs['train_loss'], train, rtol=1e-4)
    np.testing.assert_allclose(logs['test_loss'], test, rtol=1e-4)
    for name, ref in params.items():
        diff = np.abs(model.state_dict()[name].cpu().numpy() - ref)
        if name == 'net.fcon0.weight':
            assert diff.max() <= 2 * lr * n_up + 1e-6
            assert (diff > 2e-5).mean() < 1e-3, (name, (diff > 2e-5).mean())
        else:
            assert diff.max() <= 2e-5, (name, diff.max())


def test_bayessim_uses_the_fused_first_layer_and_agrees_with_the_materialised_path(monkeypatch):
    """BayesSim.run_training (reference bayes_sim.py:91-114) on Ant-shaped rollouts with
    summary_corrdiff: the factored path (default for wide summaries) and the materialised path
    (BSIG_FUSED_CORR=0) see the same random draws and must produce the same training curve."""
    import contextlib
    import io
    from bayes_sim_ig.bayes_sim import BayesSim
    d, a, t1 = TASKS['ant']
    n, p = 400, 6
    states, actions = synth_rollouts(8, n, t1, d, a, device=DEV)
    rs = np.random.RandomState(1)
    lows, highs = np.zeros(p), np.full(p, 2.0)
    params = torch.from_numpy((2.0 * rs.rand(n, p)).astype(np.float32)).to(DEV)
    cfg = {'modelClass': 'MDNN', 'summarizerFxn': 'summary_corrdiff', 'trainTrajLen': t1 - 1,
           'components': 5, 'hiddenLayers': [128, 128], 'lr': 1e-4}
    curves = {}
    for fused in ('1', '0'):
        monkeypatch.setenv('BSIG_FUSED_CORR', fused)
        torch.manual_seed(0)
        np.random.seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            bsim = BayesSim(cfg, d, a, p, lows, highs, prior=None, proposal=None, device=DEV)
            logs = bsim.run_training(params, states, actions)
        plan = list(bsim.model._plans.values())[-1]
        assert (plan.corr is not None) == (fused == '1')
        curves[fused] = logs
    np.testing.assert_allclose(curves['1']['train_loss'], curves['0']['train_loss'], rtol=2e-3)
    np.testing.assert_allclose(curves['1']['test_loss'], curves['0']['test_loss'], rtol=2e-3)


@pytest.mark.parametrize('task,m,nf_half,gather', [('ant', 100, 100, True), ('cartpole', 64, 50, False),
                                                   ('shadowhand', 100, 100, True)])
def test_fused_rff_projection_matches_oracle_features(task, m, nf_half, gather):
    """models/rff.py:128-132 on the never-materialised summary against the oracle's
    rff_features(oracle summary): a * [cos(x (freqs/sigma)^T) | sin(.)].  The angle carries the
    TF32x3 bound of the projection (2e-5 + 4e-8*F relative to its largest magnitude); cos / sin
    turn an angle error d into at most a*d."""
    from oracle import mdn_np
    lib = _lib()
    n = m + 30 if gather else m
    states, actions, cf = _factors(task, n, seed=m + nf_half)
    f = cf.shape[1]
    g = torch.Generator('cpu').manual_seed(f + nf_half)
    freqs = torch.randn(nf_half, f, generator=g)
    x_ref = _oracle_x(states, actions).astype(np.float64)
    sigma = float(4.0 * np.sqrt(f)) * float(np.abs(x_ref).mean() + 1e-3)   # angles of order one
    coeff = (freqs / sigma).to(DEV)
    rows = torch.randint(0, n, (m,), generator=g) if gather else None
    rows_dev = None if rows is None else rows.to(DEV)
    if rows is not None:
        x_ref = x_ref[rows.numpy()]
    ws = torch.empty(lib.load().bsig_corr_linear_ws_bytes(m, nf_half, cf.s, cf.q) + 256,
                     dtype=torch.uint8, device=DEV)
    out = torch.full((m, 2 * nf_half), float('nan'), device=DEV)
    scale = float(np.sqrt(1.0 / nf_half))
    lib.call('bsig_corr_rff_features', cf.fac.data_ptr(), cf.fac.shape[1],
             None if rows_dev is None else rows_dev.data_ptr(), cf.s, cf.q, coeff.data_ptr(),
             out.data_ptr(), m, nf_half, scale, ws.data_ptr(), ws.numel(), lib.stream_ptr(DEV))
    ref = mdn_np.rff_features(x_ref, freqs.double().numpy(), np.full((1, f), sigma))
    angle_max = float(np.abs(x_ref @ (freqs.double().numpy() / sigma).T).max())
    tol = scale * (2e-5 + 4e-8 * f) * max(angle_max, 1.0) + 2e-7
    assert np.abs(out.cpu().numpy() - ref).max() < tol, (np.abs(out.cpu().numpy() - ref).max(), tol)


def test_bayessim_mdrff_on_factored_summaries_agrees_with_the_materialised_path(monkeypatch):
    """BayesSim + MDRFF + summary_corrdiff on Ant-shaped rollouts (F = 11 802): the projection of
    the factored summary (default) and of the materialised one (BSIG_FUSED_CORR=0) see the same
    frequencies and random draws and must give the same training curve."""
    import contextlib
    import io
    from bayes_sim_ig.bayes_sim import BayesSim
    d, a, t1 = TASKS['ant']
    n, p = 300, 4
    states, actions = synth_rollouts(9, n, t1, d, a, device=DEV)
    states = 0.05 * states                      # keep the RBF angles moderate
    rs = np.random.RandomState(2)
    lows, highs = np.zeros(p), np.full(p, 2.0)
    params = torch.from_numpy((2.0 * rs.rand(n, p)).astype(np.float32)).to(DEV)
    cfg = {'modelClass': 'MDRFF', 'summarizerFxn': 'summary_corrdiff', 'trainTrajLen': t1 - 1,
           'components': 4, 'hiddenLayers': [], 'lr': 1e-3}
    curves = {}
    for fused in ('1', '0'):
        monkeypatch.setenv('BSIG_FUSED_CORR', fused)
        torch.manual_seed(0)
        np.random.seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            bsim = BayesSim(cfg, d, a, p, lows, highs, prior=None, proposal=None, device=DEV)
            logs = bsim.run_training(params, states, actions)
        plan = list(bsim.model._plans.values())[-1]
        assert (plan.corr is not None) == (fused == '1')
        curves[fused] = logs
    np.testing.assert_allclose(curves['1']['train_loss'], curves['0']['train_loss'], rtol=2e-3)
    np.testing.assert_allclose(curves['1']['test_loss'], curves['0']['test_loss'], rtol=2e-3)


def test_factored_path_edge_shapes():
    """Short trajectories (T < window), a single-row minibatch, an odd first-layer width and an
    empty held-out split: the fused path must agree with the materialised summary everywhere."""
    from bayes_sim_ig.models.mdnn import MDNN
    from bayes_sim_ig_b200.models.train_engine import run_training_captured
    from bayes_sim_ig_b200.utils import summarizers as bs
    lib = _lib()
    # T = 3 < 10: window = 3 steps (reference summarizers.py:97-100 only chops longer ones)
    states, actions = synth_rollouts(4, 12, 3, 6, 2)
    cf = bs.corr_factors(states.to(DEV), actions.to(DEV), use_state_diff=True)
    x = bs.summary_corrdiff(states.to(DEV), actions.to(DEV))
    assert cf.s == 3 * 5 and cf.q == 3 * 2 and torch.equal(cf.materialize(), x)
    # m = 1, n_out = 7 (odd), no gather
    f = cf.shape[1]
    g = torch.Generator('cpu').manual_seed(1)
    w = torch.randn(7, f, generator=g).to(DEV)
    b = torch.randn(7, generator=g).to(DEV)
    y = torch.empty(1, 7, device=DEV)
    ws = torch.empty(lib.load().bsig_corr_linear_ws_bytes(1, 7, cf.s, cf.q) + 256, dtype=torch.uint8,
                     device=DEV)
    lib.call('bsig_corr_linear_fwd', cf.fac.data_ptr(), cf.fac.shape[1], None, cf.s, cf.q, w.data_ptr(),
             b.data_ptr(), y.data_ptr(), 1, 7, 0, ws.data_ptr(), ws.numel(), lib.stream_ptr(DEV))
    ref = x[:1].double() @ w.double().T + b.double()
    assert _rel(y.cpu().numpy(), ref.cpu().numpy()) < 2e-5
    dy = torch.randn(1, 7, generator=g).to(DEV)
    dw = torch.empty(7, f, device=DEV)
    lib.call('bsig_corr_linear_wgrad', dy.data_ptr(), cf.fac.data_ptr(), cf.fac.shape[1], None, cf.s,
             cf.q, 1, 7, dw.data_ptr(), None, None, None, 1, 0.0, 0.9, 0.999, 1e-8, 1.0,
             lib.stream_ptr(DEV))
    assert _rel(dw.cpu().numpy(), (dy.double().T @ x[:1].double()).cpu().numpy()) < 2e-5
    # empty held-out split (test_frac = 0): nan test losses, same training losses as materialised
    lows, highs = np.zeros(3), np.ones(3)
    yv = torch.rand(12, 3, generator=g).to(DEV)
    losses = []
    for data in (cf, x):
        torch.manual_seed(5)
        model = MDNN(f, 3, lows, highs, 2, False, (16,), torch.nn.Tanh, 1e-3, device=DEV)
        rs = np.random.RandomState(0)
        inj = dict(idx=rs.randint(0, 12, (2, 5)), noise_train=rs.rand(2, 5, 3, 2).astype(np.float32),
                   noise_test=None)
        logs = run_training_captured(model, data, yv, 2, 5, test_frac=0.0, use_graph=True, injected=inj)
        assert all(np.isnan(v) for v in logs['test_loss'])
        losses.append(logs['train_loss'])
    np.testing.assert_allclose(losses[0], losses[1], rtol=1e-4)


def test_shapes_outside_the_envelope_fall_back_to_the_materialised_summary():
    """A minibatch of more than 128 rows (or a first layer wider than 128) is outside the fused
    kernels' envelope: MDNN.run_training must materialise the CorrFactors object and train on the
    generic path, with the same losses as when it is handed the summary tensor."""
    from bayes_sim_ig.models.mdnn import MDNN
    from bayes_sim_ig_b200.models.train_engine import run_training_captured
    from bayes_sim_ig_b200.utils import summarizers as bs
    states, actions = synth_rollouts(6, 200, 21, 4, 1)
    cf = bs.corr_factors(states.to(DEV), actions.to(DEV), use_state_diff=True)
    x = cf.materialize()
    lows, highs = np.zeros(2), np.ones(2)
    g = torch.Generator('cpu').manual_seed(2)
    yv = torch.rand(200, 2, generator=g).to(DEV)
    rs = np.random.RandomState(3)
    inj = dict(idx=rs.randint(0, 160, (2, 130)), noise_train=rs.rand(2, 130, 2, 3).astype(np.float32),
               noise_test=rs.rand(2, 40, 2, 3).astype(np.float32))
    curves = []
    for data, hidden in ((cf, (16,)), (x, (16,)), (cf, (130,)), (x, (130,))):
        torch.manual_seed(7)
        batch = 130 if hidden == (16,) else 64
        inj_b = dict(idx=inj['idx'][:, :batch], noise_train=inj['noise_train'][:, :batch],
                     noise_test=inj['noise_test'])
        model = MDNN(x.shape[1], 2, lows, highs, 3, False, hidden, torch.nn.Tanh, 1e-3, device=DEV)
        logs = run_training_captured(model, data, yv, 2, batch, 0.2, use_graph=True, injected=inj_b)
        plan = list(model._plans.values())[-1]
        assert plan.corr is None
        curves.append(logs['train_loss'] + logs['test_loss'])
    np.testing.assert_allclose(curves[0], curves[1], rtol=1e-6)
    np.testing.assert_allclose(curves[2], curves[3], rtol=1e-6)


def test_humanoid_shaped_summaries_with_an_odd_width_take_the_materialised_path():
    """Humanoid (D = 108, A = 21): s*q + 2 = 535*105 + 2 is odd, the rows of the first-layer weight
    are only 4-byte aligned, outside the fused kernels' envelope.  BayesSim.run_training must fall
    back to the materialised summary (bit-identical to the summarizer's tensor) and train."""
    import contextlib
    import io
    from bayes_sim_ig.bayes_sim import BayesSim
    from bayes_sim_ig_b200.utils import summarizers as bs
    d, a, t1, n, p = 108, 21, 11, 120, 3
    states, actions = synth_rollouts(12, n, t1, d, a, device=DEV)
    cf = bs.corr_factors(states, actions, use_state_diff=True)
    assert cf.shape[1] % 2 == 1
    assert torch.equal(cf.materialize(), bs.summary_corrdiff(states, actions))
    rs = np.random.RandomState(4)
    params = torch.from_numpy(rs.rand(n, p).astype(np.float32)).to(DEV)
    cfg = {'modelClass': 'MDNN', 'summarizerFxn': 'summary_corrdiff', 'trainTrajLen': t1 - 1,
           'components': 3, 'hiddenLayers': [32], 'lr': 1e-4}
    with contextlib.redirect_stdout(io.StringIO()):
        bsim = BayesSim(cfg, d, a, p, np.zeros(p), np.ones(p), prior=None, proposal=None, device=DEV)
        logs = bsim.run_training(params, states, actions)
    plan = list(bsim.model._plans.values())[-1]
    assert plan.corr is None
    assert np.isfinite(logs['train_loss']).all() and np.isfinite(logs['test_loss']).all()


@pytest.mark.parametrize('seed', list(range(24)))
def test_fused_kernels_on_random_shapes(seed):
    """Randomised sweep over (D, A, T, minibatch, first-layer width, gather): forward, random
    Fourier projection, weight gradient and Adam epilogue against float64 torch arithmetic on the
    expanded factors (whose equality with the oracle summary is pinned above).  Shapes outside
    the envelope (odd width, > 128 rows / outputs) must be refused by the applicability query and
    by the kernels themselves."""
    from bayes_sim_ig_b200.utils import summarizers as bs
    lib = _lib()
    rs = np.random.RandomState(1000 + seed)
    d = int(rs.choice([2, 3, 4, 7, 12, 23, 52, 61]))
    a = int(rs.choice([1, 2, 3, 6, 8]))
    t1 = int(rs.choice([2, 3, 6, 11, 15]))
    m = int(rs.choice([1, 2, 7, 31, 64, 100, 128]))
    n_out = int(rs.choice([1, 2, 5, 16, 33, 64, 127, 128]))
    gather = bool(rs.randint(2))
    n = m + int(rs.randint(0, 9))
    states, actions = synth_rollouts(seed, n, t1, d, a, device=DEV)
    cf = bs.corr_factors(states, actions, use_state_diff=bool(rs.randint(2)))
    f = cf.shape[1]
    ok = lib.load().bsig_corr_linear_applicable(m, m, n_out, cf.s, cf.q)
    g = torch.Generator('cpu').manual_seed(seed)
    w = (torch.randn(n_out, f, generator=g) / np.sqrt(f)).to(DEV)
    b = torch.randn(n_out, generator=g).to(DEV)
    dy = torch.randn(m, n_out, generator=g).to(DEV)
    rows = torch.randint(0, n, (m,), generator=g).to(DEV) if gather else None
    rows_p = None if rows is None else rows.data_ptr()
    y = torch.empty(m, n_out, device=DEV)
    ws = torch.empty(max(int(lib.load().bsig_corr_linear_ws_bytes(m, n_out, cf.s, cf.q)), 0) + 256,
                     dtype=torch.uint8, device=DEV)
    args_fwd = (cf.fac.data_ptr(), cf.fac.shape[1], rows_p, cf.s, cf.q, w.data_ptr(), b.data_ptr(),
                y.data_ptr(), m, n_out, 1, ws.data_ptr(), ws.numel(), lib.stream_ptr(DEV))
    if not ok:
        assert f % 2 == 1                          # the only way these sizes leave the envelope
        with pytest.raises(lib.BsigError):
            lib.call('bsig_corr_linear_fwd', *args_fwd)
        return
    x = cf.materialize().double()
    x = x[rows] if rows is not None else x[:m]
    tol = 2e-5 + 4e-8 * f
    lib.call('bsig_corr_linear_fwd', *args_fwd)
    ref = torch.tanh(x @ w.double().T + b.double())
    assert _rel(y.cpu().numpy(), ref.cpu().numpy()) < tol, ('fwd', d, a, t1, m, n_out)
    if n_out <= 128 and 2 * n_out <= 256:
        out = torch.empty(m, 2 * n_out, device=DEV)
        lib.call('bsig_corr_rff_features', cf.fac.data_ptr(), cf.fac.shape[1], rows_p, cf.s, cf.q,
                 w.data_ptr(), out.data_ptr(), m, n_out, 0.5, ws.data_ptr(), ws.numel(),
                 lib.stream_ptr(DEV))
        ang = x @ w.double().T
        ref2 = 0.5 * torch.cat([torch.cos(ang), torch.sin(ang)], dim=1)
        assert float((out.double() - ref2).abs().max()) < 0.5 * tol * max(float(ang.abs().max()), 1.0) + 2e-7
    wv, ea, es = w.clone(), torch.rand_like(w) * 0.01, torch.rand_like(w) * 1e-4
    w0, ea0, es0 = wv.double().clone(), ea.double().clone(), es.double().clone()
    dw = torch.empty_like(w)
    lib.call('bsig_corr_linear_wgrad', dy.data_ptr(), cf.fac.data_ptr(), cf.fac.shape[1], rows_p, cf.s,
             cf.q, m, n_out, dw.data_ptr(), wv.data_ptr(), ea.data_ptr(), es.data_ptr(), 3, 1e-3, 0.9,
             0.999, 1e-8, 1.0, lib.stream_ptr(DEV))
    g_ref = dy.double().T @ x
    assert _rel(dw.cpu().numpy(), g_ref.cpu().numpy()) < 2e-5 + 4e-8 * m, ('wgrad', d, a, t1, m, n_out)
    gk = dw.double()                               # Adam arithmetic on the kernel's own gradient
    m_ref = 0.9 * ea0 + 0.1 * gk
    v_ref = 0.999 * es0 + 0.001 * gk * gk
    denom = v_ref.sqrt() / np.sqrt(1 - 0.999 ** 3) + 1e-8
    p_ref = w0 - (1e-3 / (1 - 0.9 ** 3)) * m_ref / denom
    assert float((wv.double() - p_ref).abs().max()) < 1e-6
    assert _rel(ea.cpu().numpy(), m_ref.cpu().numpy()) < 1e-5
    assert _rel(es.cpu().numpy(), v_ref.cpu().numpy()) < 1e-4
