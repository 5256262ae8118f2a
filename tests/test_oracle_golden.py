"""The oracle (oracle/*.py) against the golden vectors recorded from the live
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import mdn_np, pdf_np, signature_np, summarizers_np as osum

SUMM_CASES = ['pendulum', 'cartpole', 'ant', 'humanoid', 'exact10', 'short6', 'single_pad']


@pytest.mark.parametrize('case', SUMM_CASES)
def test_summarizers_match_reference(golden, case):
    g = golden('summarizers')
    s, a = g[case + '.states'], g[case + '.actions']
    if case + '.summary_start' in g:
        np.testing.assert_array_equal(osum.summary_start(s, a), g[case + '.summary_start'])
        np.testing.assert_array_equal(osum.summary_waypts(s, a), g[case + '.summary_waypts'])
    for fn in ('summary_corr', 'summary_corrdiff'):
        ref = g[case + '.' + fn]
        got = getattr(osum, fn)(s, a)
        assert got.shape == ref.shape and got.dtype == ref.dtype
        # the outer-product block is one fp32 multiply per entry: bit-exact
        np.testing.assert_array_equal(got[:, :-2], ref[:, :-2])
        np.testing.assert_allclose(got[:, -2:], ref[:, -2:], rtol=2e-6, atol=1e-7)


def test_pad_states_actions_raises_for_multi_traj_padding(golden):
    g = golden('summarizers')
    s, a = g['short6.states'], g['short6.actions']
    with pytest.raises(RuntimeError):
        osum.summary_start(s, a)


def test_signature_depth(golden):
    g = golden('summarizers')
    got = [osum.signature_depth(int(c)) for c in g['signature_depth.in']]
    assert got == list(g['signature_depth.out'])


def test_signature_two_formulations_agree():
    rs = np.random.RandomState(0)
    for c, length in ((2, 3), (5, 21), (6, 11), (9, 7)):
        path = rs.randn(4, length, c)
        a = signature_np.signature(path, 3)
        b = signature_np.signature_iterated_sums(path, 3)
        np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-10)


def test_signature_identities():
    rs = np.random.RandomState(1)
    c = 4
    path = rs.randn(3, 9, c)
    sig = signature_np.signature(path, 3)
    s1 = sig[:, :c]
    s2 = sig[:, c:c + c * c].reshape(-1, c, c)
    np.testing.assert_allclose(s1, path[:, -1] - path[:, 0], atol=1e-12)
    # shuffle identity: S^i S^j = S^{ij} + S^{ji}
    np.testing.assert_allclose(s1[:, :, None] * s1[:, None, :],
                               s2 + np.transpose(s2, (0, 2, 1)), atol=1e-10)
    # invariance to a repeated point
    rep = np.concatenate([path[:, :4], path[:, 3:4], path[:, 4:]], axis=1)
    np.testing.assert_allclose(signature_np.signature(rep, 3), sig, atol=1e-10)
    # straight line == tensor exponential of the increment
    line = np.stack([np.zeros((3, c)), rs.randn(3, c)], axis=1)
    d = line[:, 1]
    exp3 = np.einsum('ni,nj,nk->nijk', d, d, d).reshape(3, -1) / 6.0
    np.testing.assert_allclose(signature_np.signature(line, 3)[:, c + c * c:], exp3, atol=1e-12)
    # Chen: Sig(a*b) level 2 = S2(a) + S2(b) + S1(a) (x) S1(b)
    pa, pb = path[:, :5], path[:, 4:]
    sa, sb = signature_np.signature(pa, 2), signature_np.signature(pb, 2)
    chen2 = sa[:, c:] + sb[:, c:] + (sa[:, :c, None] * sb[:, None, :c]).reshape(3, -1)
    np.testing.assert_allclose(sig[:, c:c + c * c], chen2, atol=1e-10)


MDN_CASES = ['diag', 'full', 'full_big', 'p1', 'rff', 'rff_full']


def _mdn_case(g, case):
    meta = g[case + '.meta']
    din, p, k, full, b = [int(v) for v in meta[:5]]
    rff = None
    if case + '.rff.freqs' in g:
        rff = (g[case + '.rff.freqs'], g[case + '.rff.sigma'])
    return din, p, k, bool(full), b, rff


@pytest.mark.parametrize('case', MDN_CASES)
def test_mdn_forward_loss_grads_adam(golden, case):
    g = golden('mdn')
    din, p, k, full, b, rff = _mdn_case(g, case)
    x, y = g[case + '.x'], g[case + '.y']
    params = {key: val.astype(np.float64) for key, val in g.sub(case + '.init.').items()}
    m = {key: np.zeros_like(val) for key, val in params.items()}
    v = {key: np.zeros_like(val) for key, val in params.items()}
    if rff is not None:
        np.testing.assert_allclose(mdn_np.rff_features(x, *rff), g[case + '.rff.features'],
                                   rtol=2e-5, atol=2e-6)
    for step in range(3):
        tag = '%s.step%d.' % (case, step)
        noise = g[tag + 'noise']
        w, mu, ld, low, _ = mdn_np.mdnn_forward(params, x, noise, p, k, rff=rff)
        np.testing.assert_allclose(w, g[tag + 'weights'], rtol=2e-5, atol=1e-7)
        np.testing.assert_allclose(mu, g[tag + 'mu'], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(ld, g[tag + 'L_d'], rtol=2e-5, atol=1e-7)
        if full and p > 1:
            np.testing.assert_allclose(low, g[tag + 'L'], rtol=2e-5, atol=2e-6)
        else:
            assert low is None
        loss, grads = mdn_np.mdnn_loss_and_grads(params, x, y, noise, p, k, rff=rff)
        np.testing.assert_allclose(loss, g[tag + 'loss'], rtol=2e-5)
        for key, ref in g.sub(tag + 'grad.').items():
            scale = np.abs(ref).max() + 1e-12
            assert np.abs(grads[key] - ref).max() <= 5e-5 * scale + 1e-7, key
        params, m, v = mdn_np.adam_step(params, grads, m, v, step + 1, 1e-3)
        for key, ref in g.sub(tag + 'after.').items():
            # Adam's first steps are sign-like (|update| ~ lr): compare to lr scale
            assert np.abs(params[key] - ref).max() <= 2e-5, key


@pytest.mark.parametrize('case', MDN_CASES)
def test_predict_mog_params(golden, case):
    g = golden('mdn')
    din, p, k, full, b, rff = _mdn_case(g, case)
    params = {key: val.astype(np.float64) for key, val in g.sub(case + '.step2.after.').items()}
    xs = g[case + '.predict.xs']
    w, mu, ld, low, _ = mdn_np.mdnn_forward(params, xs, g[case + '.predict.noise'], p, k, rff=rff)
    a, means, packed = mdn_np.predict_mog_params(w, mu, ld, low, g[case + '.lows'], g[case + '.highs'])
    np.testing.assert_allclose(a, g[case + '.predict.a'], rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(means, g[case + '.predict.m'], rtol=3e-5, atol=3e-6)
    for r in range(xs.shape[0]):
        for c in range(k):
            gs = pdf_np.gaussian_from_packed_factor(means[r, c], packed[r, c])
            np.testing.assert_allclose(gs['S'], g[case + '.predict.S'][r, c], rtol=1e-4, atol=1e-6)
            np.testing.assert_allclose(gs['logdetP'], g[case + '.predict.logdetP'][r, c],
                                       rtol=1e-4, atol=1e-5)


PDF_CASES = ['f32', 'f64', 'p1', 'p13']


@pytest.mark.parametrize('case', PDF_CASES)
def test_pdf_ctor_gen_eval(golden, case):
    g = golden('pdf')
    a, ms, ls = g[case + '.a'], g[case + '.ms'], g[case + '.Ls']
    gs = [pdf_np.gaussian_from_packed_factor(m, l) for m, l in zip(ms, ls)]
    for fld in ('C', 'S', 'P', 'Pm'):
        got = np.stack([x[fld] for x in gs])
        assert got.dtype == g[case + '.' + fld].dtype
        np.testing.assert_array_equal(got, g[case + '.' + fld])
    np.testing.assert_array_equal(np.array([x['logdetP'] for x in gs]), g[case + '.logdetP'])
    # index work: bit exact given the same uniforms
    idx = pdf_np.discrete_sample_from_u(a, g[case + '.discrete.u'])
    np.testing.assert_array_equal(idx, g[case + '.discrete.idx'])
    smp, _, counts = pdf_np.mog_gen_from_draws(
        a, ms, [x['C'] for x in gs], g[case + '.gen.u'], g[case + '.gen.z'])
    assert sum(counts) == smp.shape[0]
    np.testing.assert_array_equal(smp, g[case + '.gen.samples'])
    # posterior consumer (ParamsGenerator.sample per environment): bit exact
    env, _ = pdf_np.params_samples_per_env(a, ms, [x['C'] for x in gs], g[case + '.envs.u'],
                                           g[case + '.envs.z'], g[case + '.envs.lows'],
                                           g[case + '.envs.highs'])
    np.testing.assert_array_equal(env, g[case + '.envs.samples'])
    x64 = g[case + '.eval.x64']
    precs = [x['P'] for x in gs]
    lds = [x['logdetP'] for x in gs]
    np.testing.assert_allclose(pdf_np.mog_logpdf(x64, a, ms, precs, lds),
                               g[case + '.eval.log64'], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(pdf_np.mog_logpdf(x64, a, ms, precs, lds, log=False),
                               g[case + '.eval.lin64'], rtol=1e-10, atol=1e-300)


@pytest.mark.parametrize('case', ['diag', 'full_big', 'rff'])
def test_torch_port_matches_reference(golden, case):
    """oracle/torch_port.py (the CPU timing baseline) issues the reference's torch
    ops: with the recorded noise it must reproduce losses, grads and Adam steps."""
    import torch
    from helpers import injected_rand_like
    from oracle import torch_port
    torch.set_num_threads(1)
    g = golden('mdn')
    din, p, k, full, b, rff = _mdn_case(g, case)
    hidden = tuple(int(v) for v in g[case + '.meta'][5:])
    rff_t = None if rff is None else (torch.from_numpy(rff[0]), torch.from_numpy(rff[1]))
    model = torch_port.PortModel(din, p, g[case + '.lows'], g[case + '.highs'], k, full, hidden,
                                 1e-3, rff=rff_t)
    model.load_state_dict({key: torch.from_numpy(v) for key, v in g.sub(case + '.init.').items()})
    x, y = torch.from_numpy(g[case + '.x']), torch.from_numpy(g[case + '.y'])
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    for step in range(3):
        tag = '%s.step%d.' % (case, step)
        with injected_rand_like([g[tag + 'noise']], 'cpu'):
            opt.zero_grad()
            loss = model.loss(*model(x), y)
            loss.backward()
        np.testing.assert_allclose(loss.item(), g[tag + 'loss'], rtol=1e-6)
        for name, prm in model.named_parameters():
            np.testing.assert_allclose(prm.grad.numpy(), g[tag + 'grad.' + name], rtol=1e-4, atol=1e-7)
        opt.step()
    for name, ref in g.sub(case + '.step2.after.').items():
        np.testing.assert_allclose(model.state_dict()[name].numpy(), ref, rtol=1e-5, atol=1e-6)


def test_torch_port_summarizers_match_reference(golden):
    import torch
    from oracle import torch_port
    g = golden('summarizers')
    for case in ('pendulum', 'cartpole', 'ant'):
        s, a = torch.from_numpy(g[case + '.states']), torch.from_numpy(g[case + '.actions'])
        assert np.array_equal(torch_port.summary_start(s, a).numpy(), g[case + '.summary_start'])
        assert np.array_equal(torch_port.summary_corrdiff(s, a).numpy(), g[case + '.summary_corrdiff'])
        assert np.array_equal(torch_port.summary_corr(s, a).numpy(), g[case + '.summary_corr'])


@pytest.mark.parametrize('case', ['bench_diag', 'bench_full'])
def test_oracle_full_update_at_bench_shape(golden, case):
    """The float64 oracle replays the reference's recorded updates at BASELINE configs[1]'s
    dimensions (minibatch 100, F = 302, 128 x 128, P = 13, K = 10; tests/golden/mdn_bench.npz,
    reference loop body mdnn.py:221-234): per-update losses and the final parameters."""
    from oracle import mdn_np
    g = golden('mdn_bench')
    meta = g[case + '.meta']
    p, k = int(meta[1]), int(meta[2])
    idx, noise = g[case + '.idx'], g[case + '.noise']
    prm = {key: val.astype(np.float64) for key, val in g.sub(case + '.init.').items()}
    m = {key: np.zeros_like(val) for key, val in prm.items()}
    v = {key: np.zeros_like(val) for key, val in prm.items()}
    x = g[case + '.x'].astype(np.float64)
    y = mdn_np.normalize_samples(g[case + '.y_raw'].astype(np.float64), g[case + '.lows'],
                                 g[case + '.highs'])
    for step in range(idx.shape[0]):
        rows = idx[step]
        loss, grads = mdn_np.mdnn_loss_and_grads(prm, x[rows], y[rows], noise[step], p, k)
        assert abs(loss - float(g[case + '.loss'][step])) <= 1e-5 * abs(loss)
        prm, m, v = mdn_np.adam_step(prm, grads, m, v, step + 1, 1e-3)
    for key, ref in g.sub(case + '.after.').items():
        assert np.abs(prm[key] - ref).max() <= 2e-5, key
