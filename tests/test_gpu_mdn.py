"""GPU parity: MDNN / MDRFF forward, loss, gradients, Adam and predict_MoGs
against the golden vectors recorded from the live reference, through both the
autograd (API) form and the fused CUDA-graph training form."""
import numpy as np
import pytest
import torch

from helpers import injected_rand_like, load_state, mdn_meta, rel_err
from oracle import mdn_np

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
CASES = ['diag', 'full', 'full_big', 'p1', 'rff', 'rff_full']
# fp32 tolerance of the north star: 1e-5 relative; gradients are compared
# against the reference's fp32 autograd, both sides carry ~1e-6 rounding noise
# which shows on near-zero entries, hence max-norm relative error.
RTOL = 1e-5
GRAD_RTOL = 5e-5


def build(g, case):
    from bayes_sim_ig.models.mdnn import MDNN
    from bayes_sim_ig.models.mdrff import MDRFF
    din, p, k, full, b, hidden = mdn_meta(g, case)
    kw = dict(input_dim=din, output_dim=p, output_lows=g[case + '.lows'],
              output_highs=g[case + '.highs'], n_gaussians=k, full_covariance=full,
              hidden_layers=hidden, activation=torch.nn.Tanh, lr=1e-3, device=DEV)
    if case.startswith('rff'):
        model = MDRFF(n_feat=20, sigma=4.0, kernel='RBF', **kw)
        model.rff.freqs = torch.from_numpy(g[case + '.rff.freqs']).to(DEV)
        model.rff.sigma = torch.from_numpy(g[case + '.rff.sigma']).to(DEV)
    else:
        model = MDNN(**kw)
    load_state(model, g.sub(case + '.init.'))
    return model, (din, p, k, full, b)


@pytest.mark.parametrize('case', CASES)
def test_forward_loss_grads_adam_api_path(golden, case):
    g = golden('mdn')
    model, (din, p, k, full, b) = build(g, case)
    assert sorted(model.state_dict().keys()) == sorted(g.sub(case + '.init.').keys())
    x = torch.from_numpy(g[case + '.x']).to(DEV)
    y = torch.from_numpy(g[case + '.y']).to(DEV)
    if case.startswith('rff'):
        feats = model.rff.to_features(x).cpu().numpy()
        assert rel_err(feats, g[case + '.rff.features']) < RTOL
    opt = torch.optim.Adam(model.parameters(), lr=model.lr)
    for step in range(3):
        tag = '%s.step%d.' % (case, step)
        with injected_rand_like([g[tag + 'noise']], DEV):
            opt.zero_grad()
            w, mu, ld, low = model(x)
            loss = model.mdn_loss_fn(w, mu, ld, low, y)
            loss.backward()
        assert rel_err(w.detach().cpu(), g[tag + 'weights']) < RTOL
        assert rel_err(mu.detach().cpu(), g[tag + 'mu']) < RTOL
        assert rel_err(ld.detach().cpu(), g[tag + 'L_d']) < RTOL
        if full and p > 1:
            assert rel_err(low.detach().cpu(), g[tag + 'L']) < RTOL
        else:
            assert low is None
        assert abs(loss.item() - float(g[tag + 'loss'])) <= RTOL * abs(float(g[tag + 'loss'])) + 1e-6
        for name, prm in model.named_parameters():
            assert rel_err(prm.grad.cpu(), g[tag + 'grad.' + name]) < GRAD_RTOL, (step, name)
        opt.step()
        for name, ref in g.sub(tag + 'after.').items():
            got = model.state_dict()[name].cpu().numpy()
            assert np.abs(got - ref).max() <= 2e-5, (step, name)   # lr = 1e-3 scale


@pytest.mark.parametrize('case', CASES)
def test_fused_training_form_matches_reference(golden, case):
    """run_training's captured form (gather + fused head/NLL + wgrad/dgrad + Adam)
    replayed with the reference's recorded noise: same losses and parameters."""
    from bayes_sim_ig_b200.models.train_engine import run_training_captured
    g = golden('mdn')
    for use_graph in (False, True):
        model, (din, p, k, full, b) = build(g, case)
        x = torch.from_numpy(g[case + '.x']).to(DEV)
        y_raw = torch.from_numpy(g[case + '.y_raw']).to(DEV)
        noise = np.stack([g['%s.step%d.noise' % (case, s)] for s in range(3)])
        inj = dict(idx=np.tile(np.arange(b), (3, 1)), noise_train=noise, noise_test=None)
        logs = run_training_captured(model, x, y_raw, 3, b, test_frac=0.0,
                                     use_graph=use_graph, injected=inj)
        ref_losses = [float(g['%s.step%d.loss' % (case, s)]) for s in range(3)]
        np.testing.assert_allclose(logs['train_loss'], ref_losses, rtol=2e-5, atol=1e-6)
        assert all(np.isnan(v) for v in logs['test_loss'])       # empty test split -> nan
        for name, ref in g.sub(case + '.step2.after.').items():
            got = model.state_dict()[name].cpu().numpy()
            assert np.abs(got - ref).max() <= 3e-5, (use_graph, name)


@pytest.mark.parametrize('case', CASES)
def test_predict_mogs(golden, case):
    g = golden('mdn')
    model, (din, p, k, full, b) = build(g, case)
    load_state(model, g.sub(case + '.step2.after.'))
    xs = torch.from_numpy(g[case + '.predict.xs']).to(DEV)
    with injected_rand_like([g[case + '.predict.noise']], DEV):
        mogs = model.predict_MoGs(xs)
    assert len(mogs) == xs.shape[0]
    for r, mog in enumerate(mogs):
        assert mog.a.dtype == np.float32 and mog.xs[0].m.dtype == np.float32
        assert rel_err(mog.a, g[case + '.predict.a'][r]) < RTOL
        for c, comp in enumerate(mog.xs):
            assert rel_err(comp.m, g[case + '.predict.m'][r, c]) < RTOL
            assert rel_err(comp.C, g[case + '.predict.C'][r, c]) < 2e-5
            assert rel_err(comp.S, g[case + '.predict.S'][r, c]) < 5e-5
            assert abs(comp.logdetP - g[case + '.predict.logdetP'][r, c]) < 1e-3


def test_full_cov_multi_row_predict_uses_each_rows_factor(golden):
    """the reference raises for R > 1 with full covariance; here row r
    uses L[r] -- checked against the oracle."""
    g = golden('mdn')
    case = 'full'
    model, (din, p, k, full, b) = build(g, case)
    x = torch.from_numpy(g[case + '.x'][:3]).to(DEV)
    noise = np.random.RandomState(0).rand(3, p, k).astype(np.float32)
    with injected_rand_like([noise], DEV):
        mogs = model.predict_MoGs(x)
    params = {key: val.astype(np.float64) for key, val in g.sub(case + '.init.').items()}
    w, mu, ld, low, _ = mdn_np.mdnn_forward(params, g[case + '.x'][:3], noise, p, k)
    a, means, packed = mdn_np.predict_mog_params(w, mu, ld, low, g[case + '.lows'], g[case + '.highs'])
    for r in range(3):
        for c in range(k):
            assert rel_err(mogs[r].xs[c].m, means[r, c]) < RTOL
            lower = np.zeros((p, p))
            lower[np.arange(p), np.arange(p)] = packed[r, c, :p]
            rows, cols = np.tril_indices(p, -1)
            lower[rows, cols] = packed[r, c, p:]
            assert rel_err(mogs[r].xs[c].C, lower.T) < 2e-5


def test_nll_edge_values_and_nonfinite_assert(golden):
    g = golden('mdn')
    model, (din, p, k, full, b) = build(g, 'diag')
    x = torch.from_numpy(g['diag.x']).to(DEV)
    y = torch.from_numpy(g['diag.y']).to(DEV)
    w, mu, ld, low = model(x)
    bad = mu.clone()
    bad[0, 0, 0] = float('nan')
    with pytest.raises(AssertionError):
        model.mdn_loss_fn(w, bad, ld, low, y)
    # far-away targets hit the +-1e5 clamp: loss stays finite and equals the oracle
    far = y + 1e4
    tiny = ld * 1e-3
    loss = model.mdn_loss_fn(w, mu, tiny, low, far).item()
    ref = mdn_np.mdn_loss(w.detach().cpu().numpy().astype(np.float64), mu.detach().cpu().numpy().astype(np.float64),
                          tiny.detach().cpu().numpy().astype(np.float64), None, far.cpu().numpy())
    assert np.isfinite(loss) and abs(loss - ref) <= 1e-5 * abs(ref)


@pytest.mark.parametrize('k', [1, 3, 10, 17, 40, 70])
def test_nll_component_counts_vs_oracle(k):
    """Sub-warp group widths 1..32 and multi-component lanes (K > 32)."""
    from bayes_sim_ig.models.mdnn import MDNN
    rs = np.random.RandomState(k)
    p, b = 3, 37
    model = MDNN(5, p, None, None, k, True, (8,), torch.nn.Tanh, 1e-3, device=DEV)
    w = rs.rand(b, k) + 0.01
    w = (w / w.sum(1, keepdims=True)).astype(np.float32)
    mu = rs.randn(b, p, k).astype(np.float32)
    ld = (0.2 + rs.rand(b, p, k)).astype(np.float32)
    low = (0.3 * rs.randn(b, 3, k)).astype(np.float32)
    y = rs.randn(b, p).astype(np.float32)
    t = [torch.from_numpy(v).to(DEV).requires_grad_(True) for v in (w, mu, ld, low)]
    loss = model.mdn_loss_fn(t[0], t[1], t[2], t[3], torch.from_numpy(y).to(DEV))
    loss.backward()
    ref_loss, d_w, d_mu, d_ld, d_low = mdn_np.mdn_loss_backward(
        w.astype(np.float64), mu.astype(np.float64), ld.astype(np.float64), low.astype(np.float64), y)
    assert abs(loss.item() - ref_loss) <= 1e-5 * abs(ref_loss)
    for got, ref in zip(t, (d_w, d_mu, d_ld, d_low)):
        assert rel_err(got.grad.cpu(), ref) < GRAD_RTOL


@pytest.mark.parametrize('cfg', [(3000, 13, 10, False), (700, 2, 4, False), (5000, 37, 10, False),
                                 (3000, 4, 3, True), (300, 13, 10, False), (260, 5, 32, False),
                                 (30001, 13, 10, False), (5003, 13, 10, True), (4000, 3, 40, False),
                                 (9001, 2, 70, True), (2050, 37, 10, True)])
def test_fused_head_nll_kernels_at_every_batch_regime(cfg):
    """bsig_mdn_nll_fused picks a register-resident cluster kernel (minibatch), the
    persistent bulk-copy staged streaming kernel (large batches; many tiles per CTA,
    ragged last tile, K > 32) or the plain group-per-sample kernel (rows too wide
    for shared memory): all against the float64 oracle, loss and d loss / d z."""
    from bayes_sim_ig_b200 import _lib
    b, p, k, full = cfg
    rs = np.random.RandomState(b + p)
    lsz = p * (p - 1) // 2 if full else 0
    nh = k * (1 + 2 * p + lsz)
    z = (0.4 * rs.randn(b, nh)).astype(np.float32)
    noise = rs.rand(b, p, k).astype(np.float32)
    y = rs.rand(b + 10, p).astype(np.float32)
    rows = rs.randint(0, b + 10, b).astype(np.int64)
    zt, nt, yt, rt = (torch.from_numpy(v).to(DEV) for v in (z, noise, y, rows))
    loss = torch.zeros(1, device=DEV)
    dz = torch.empty_like(zt)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    ws = torch.zeros(_lib.load().bsig_mdn_ws_bytes(b), dtype=torch.uint8, device=DEV)
    for _ in range(2):      # twice: the workspace counters must self-reset
        _lib.call('bsig_mdn_nll_fused', zt.data_ptr(), nt.data_ptr(), yt.data_ptr(), rt.data_ptr(),
                  loss.data_ptr(), dz.data_ptr(), b, p, k, 1 if full else 0, ws.data_ptr(),
                  ws.numel(), flag.data_ptr(), _lib.stream_ptr(DEV))
    z64 = z.astype(np.float64)
    pk = p * k
    w, mu, ld, low, cache = mdn_np.head_epilogue(
        z64[:, :k], z64[:, k:k + pk], z64[:, k + pk:k + 2 * pk],
        z64[:, k + 2 * pk:] if full else None, noise, p, k)
    ref_loss, d_w, d_mu, d_ld, d_low = mdn_np.mdn_loss_backward(w, mu, ld, low, y[rows].astype(np.float64))
    d_zpi, d_zmu, d_zd, d_zl = mdn_np.head_epilogue_backward(cache, w, d_w, d_mu, d_ld, d_low)
    ref_dz = np.concatenate([d_zpi, d_zmu, d_zd] + ([d_zl] if full else []), axis=1)
    assert int(flag.item()) == 0
    assert abs(loss.item() - ref_loss) <= 1e-5 * abs(ref_loss)
    assert rel_err(dz.cpu(), ref_dz) < GRAD_RTOL
    # forward-only form gives the same loss
    loss2 = torch.zeros(1, device=DEV)
    _lib.call('bsig_mdn_nll_fused', zt.data_ptr(), nt.data_ptr(), yt.data_ptr(), rt.data_ptr(),
              loss2.data_ptr(), None, b, p, k, 1 if full else 0, ws.data_ptr(), ws.numel(),
              flag.data_ptr(), _lib.stream_ptr(DEV))
    assert abs(loss2.item() - ref_loss) <= 1e-5 * abs(ref_loss)


def test_streaming_nll_without_gather_and_with_misaligned_rows():
    """y given in batch order (no row gather), and a z / dz pair that is only 4-byte
    aligned (bulk copies impossible -> plain kernel): identical results."""
    from bayes_sim_ig_b200 import _lib
    b, p, k = 6000, 13, 10
    rs = np.random.RandomState(3)
    nh = k * (1 + 2 * p)
    z = (0.4 * rs.randn(b, nh)).astype(np.float32)
    noise = rs.rand(b, p, k).astype(np.float32)
    y = rs.rand(b, p).astype(np.float32)
    nt, yt = torch.from_numpy(noise).to(DEV), torch.from_numpy(y).to(DEV)
    ws = torch.zeros(_lib.load().bsig_mdn_ws_bytes(b), dtype=torch.uint8, device=DEV)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    outs = []
    for shift in (0, 1):
        zbuf = torch.zeros(b * nh + 4, device=DEV)
        dzbuf = torch.zeros(b * nh + 4, device=DEV)
        zt = zbuf[shift:shift + b * nh].view(b, nh)
        zt.copy_(torch.from_numpy(z))
        dz = dzbuf[shift:shift + b * nh].view(b, nh)
        loss = torch.zeros(1, device=DEV)
        _lib.call('bsig_mdn_nll_fused', zt.data_ptr(), nt.data_ptr(), yt.data_ptr(), None,
                  loss.data_ptr(), dz.data_ptr(), b, p, k, 0, ws.data_ptr(), ws.numel(),
                  flag.data_ptr(), _lib.stream_ptr(DEV))
        outs.append((loss.item(), dz.clone()))
    z64 = z.astype(np.float64)
    pk = p * k
    w, mu, ld, low, cache = mdn_np.head_epilogue(z64[:, :k], z64[:, k:k + pk], z64[:, k + pk:], None,
                                                 noise, p, k)
    ref_loss, d_w, d_mu, d_ld, d_low = mdn_np.mdn_loss_backward(w, mu, ld, low, y.astype(np.float64))
    d_zpi, d_zmu, d_zd, _ = mdn_np.head_epilogue_backward(cache, w, d_w, d_mu, d_ld, d_low)
    ref_dz = np.concatenate([d_zpi, d_zmu, d_zd], axis=1)
    for loss, dz in outs:
        assert abs(loss - ref_loss) <= 1e-5 * abs(ref_loss)
        assert rel_err(dz.cpu(), ref_dz) < GRAD_RTOL
    # same per-sample formulas in both kernels (SFU vs full-precision transcendentals)
    assert rel_err(outs[0][1].cpu(), outs[1][1].cpu()) < 1e-5


def test_graph_replay_follows_an_edited_rff_bandwidth(golden):
    """The captured training graph points at the cached `freqs / sigma` tensor of the RFF.  Editing
    sigma (or freqs) replaces that tensor; the next run_training call must notice and record the
    call again instead of replaying against the old (freed) coefficients."""
    from bayes_sim_ig_b200.models.train_engine import run_training_captured
    g = golden('mdn')
    case = 'rff'
    x = torch.from_numpy(g[case + '.x']).to(DEV)
    y_raw = torch.from_numpy(g[case + '.y_raw']).to(DEV)
    model, (din, p, k, full, b) = build(g, case)
    noise = np.stack([g['%s.step%d.noise' % (case, s)] for s in range(3)])
    inj = dict(idx=np.tile(np.arange(b), (3, 1)), noise_train=noise, noise_test=None)
    init = {name: v.clone() for name, v in model.state_dict().items()}
    first = run_training_captured(model, x, y_raw, 3, b, test_frac=0.0, use_graph=True, injected=inj)
    model.load_state_dict(init)
    model.rff.sigma.mul_(2.0)                      # in place: bumps the tensor version
    edited = run_training_captured(model, x, y_raw, 3, b, test_frac=0.0, use_graph=True, injected=inj)
    # reference for the edited bandwidth: a fresh model, no graph
    ref_model, _ = build(g, case)
    ref_model.rff.sigma = ref_model.rff.sigma * 2.0
    ref = run_training_captured(ref_model, x, y_raw, 3, b, test_frac=0.0, use_graph=False, injected=inj)
    assert abs(edited['train_loss'][0] - first['train_loss'][0]) > 1e-4      # it did change
    np.testing.assert_allclose(edited['train_loss'], ref['train_loss'], rtol=1e-6)
