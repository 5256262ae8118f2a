"""GPU parity: summarizer kernels vs the golden vectors and the oracle."""
import numpy as np
import pytest
import torch

from helpers import rel_err, synth_rollouts
from oracle import summarizers_np as osum

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


def _check_cross(got, ref):
    got = got.cpu().numpy()
    assert got.shape == ref.shape and got.dtype == ref.dtype
    # one fp32 multiply per entry: bit exact
    np.testing.assert_array_equal(got[:, :-2], ref[:, :-2])
    # mean / unbiased std: tolerance 2e-6 relative (reduction order differs)
    np.testing.assert_allclose(got[:, -2:], ref[:, -2:], rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize('case', ['pendulum', 'cartpole', 'ant', 'humanoid', 'exact10'])
def test_start_waypts_bit_exact_vs_reference(golden, case):
    from bayes_sim_ig.utils import summarizers as S
    g = golden('summarizers')
    s, a = _dev(g[case + '.states']), _dev(g[case + '.actions'])
    np.testing.assert_array_equal(S.summary_start(s, a).cpu().numpy(), g[case + '.summary_start'])
    np.testing.assert_array_equal(S.summary_waypts(s, a).cpu().numpy(), g[case + '.summary_waypts'])


@pytest.mark.parametrize('case', ['pendulum', 'cartpole', 'ant', 'humanoid', 'exact10', 'short6',
                                  'single_pad'])
def test_crosscorr_vs_reference(golden, case, capsys):
    from bayes_sim_ig.utils import summarizers as S
    g = golden('summarizers')
    s, a = _dev(g[case + '.states']), _dev(g[case + '.actions'])
    _check_cross(S.summary_corr(s, a), g[case + '.summary_corr'])
    _check_cross(S.summary_corrdiff(s, a), g[case + '.summary_corrdiff'])
    assert 'cross_corr feats' in capsys.readouterr().out     # the reference prints too


def test_padding_behaviour_matches_reference(golden):
    from bayes_sim_ig.utils import summarizers as S
    g = golden('summarizers')
    s, a = _dev(g['single_pad.states']), _dev(g['single_pad.actions'])
    np.testing.assert_array_equal(S.summary_start(s, a).cpu().numpy(), g['single_pad.summary_start'])
    s, a = _dev(g['short6.states']), _dev(g['short6.actions'])
    with pytest.raises(RuntimeError):         # N > 1 padding is undefined in the reference (Q3)
        S.summary_start(s, a)
    ps, pa = S.pad_states_actions(s, a, 4)
    assert ps.shape[1] == 4 and pa.shape[1] == 4


def test_crosscorr_nonfinite_asserts():
    from bayes_sim_ig.utils import summarizers as S
    s, a = synth_rollouts(3, 6, 12, 4, 2, DEV)
    s[2, 1, 1] = float('inf')
    with pytest.raises(AssertionError):
        S.summary_corrdiff(s, a)


@pytest.mark.parametrize('shape', [(1, 10, 2, 1), (3, 11, 5, 3), (257, 21, 4, 1), (33, 51, 60, 8),
                                   (5, 11, 211, 20), (1000, 21, 3, 1), (2, 7, 6, 5), (9, 12, 7, 3)])
def test_summarizers_vs_oracle_random_shapes(shape):
    from bayes_sim_ig.utils import summarizers as S
    n, t1, d, a = shape
    s, ac = synth_rollouts(100 + n, n, t1, d, a)
    sd, ad = s.to(DEV), ac.to(DEV)
    if t1 >= 10:
        np.testing.assert_array_equal(S.summary_start(sd, ad).cpu().numpy(),
                                      osum.summary_start(s.numpy(), ac.numpy()))
    _check_cross(S.summary_corrdiff(sd, ad), osum.summary_corrdiff(s.numpy(), ac.numpy()))
    _check_cross(S.summary_corr(sd, ad), osum.summary_corr(s.numpy(), ac.numpy()))


def test_crosscorr_is_rank_one_per_trajectory():
    """Size-independent property at a BASELINE-sized batch: the outer-product
    block of every row has rank one and its row/column ratios are constant."""
    from bayes_sim_ig.utils import summarizers as S
    n, t1, d, a = 4096, 21, 4, 1
    s, ac = synth_rollouts(0, n, t1, d, a, DEV)
    f = S.summary_corrdiff(s, ac)
    assert f.shape == (n, 302)
    blk = f[:, :-2].reshape(n, 30, 10)
    sf = (s[:, :10, 1:] - s[:, :10, :-1]).reshape(n, 30)
    af = ac[:, :10, :].reshape(n, 10)
    assert torch.equal(blk, sf[:, :, None] * af[:, None, :])
    torch.testing.assert_close(f[:, -2], sf.mean(1), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(f[:, -1], sf.std(1), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('shape', [(7, 21, 3, 1), (5, 21, 4, 1), (3, 11, 15, 6), (2, 51, 60, 8),
                                   (2, 11, 108, 21), (40, 5, 2, 2), (3, 2, 4, 1)])
def test_signature_vs_float64_oracle(shape):
    """Signature kernel (fp32) vs the float64 restatement (PARITY UNPINNED against
    signatory itself, see oracle/signature_np.py).  Tolerance: 2e-5 of the level's
    largest magnitude (fp32 accumulation over L-1 Chen steps)."""
    from bayes_sim_ig.utils import summarizers as S
    n, t1, d, a = shape
    s, ac = synth_rollouts(7 + n, n, t1, d, a)
    s = s * 0.3
    ref = osum.summary_signatory(s.numpy(), ac.numpy()).astype(np.float64)
    got = S.summary_signatory(s.to(DEV), ac.to(DEV)).cpu().numpy()
    assert got.shape == ref.shape
    c = 1 + d + a
    depth = osum.signature_depth(c)
    off = 0
    for lvl in range(1, depth + 1):
        w = c ** lvl
        assert rel_err(got[:, off:off + w], ref[:, off:off + w]) < 2e-5, lvl
        off += w


@pytest.mark.parametrize('shape', [(70001, 21, 4, 1), (50003, 21, 3, 1), (20000, 8, 5, 2),
                                   (33333, 6, 1, 0)])
def test_signature_persistent_bulk_tiles(shape):
    """Many tiles per persistent CTA (cp.async.bulk double buffer, mbarrier phase
    wrap-around, ragged last tile): same tolerance against the float64 oracle, and
    rows must not depend on the tiling (bit-identical to small ragged launches)."""
    from bayes_sim_ig.utils import summarizers as S
    n, t1, d, a = shape
    s, ac = synth_rollouts(11, n, t1, d, max(a, 1))
    s = s * 0.3
    if a == 0:
        ac = ac[:, :, :0]
    sd, ad = s.to(DEV), ac.to(DEV)
    got_t = S.summary_signatory(sd, ad)
    got = got_t.cpu().numpy()
    ref = osum.summary_signatory(s.numpy(), ac.numpy()).astype(np.float64)
    assert got.shape == ref.shape
    c = 1 + d + a
    off = 0
    for lvl in range(1, osum.signature_depth(c) + 1):
        w = c ** lvl
        assert rel_err(got[:, off:off + w], ref[:, off:off + w]) < 2e-5, lvl
        off += w
    for lo, hi in ((0, 7), (n // 2 + 1, n // 2 + 4), (n - 5, n)):
        part = S.summary_signatory(sd[lo:hi], ad[lo:hi])
        assert torch.equal(part, got_t[lo:hi]), (lo, hi)
    # actions stored with more steps than the path uses (strided rows -> plain loads)
    if a > 0:
        longer = torch.cat([ad, ad[:, :3]], 1)[:4099].contiguous()
        part = S.summary_signatory(sd[:4099], longer)
        assert torch.equal(part, got_t[:4099])


def test_signature_chen_identity_on_device():
    """Concatenating two paths multiplies their signatures (level 2 check) --
    a property test that needs no oracle."""
    from bayes_sim_ig.utils import summarizers as S
    n, t1, d, a = 64, 21, 4, 1
    s, ac = synth_rollouts(5, n, t1, d, a, DEV)
    s = s * 0.2
    c = 1 + d + a
    full = S.summary_signatory(s, ac).double()
    s1 = full[:, :c]
    s2 = full[:, c:c + c * c].reshape(n, c, c)
    path0 = torch.cat([torch.ones(n, 1, device=DEV), s[:, 0], ac[:, 0]], 1).double()
    pathl = torch.cat([torch.full((n, 1), float(t1), device=DEV), s[:, -1], ac[:, -1]], 1).double()
    torch.testing.assert_close(s1, pathl - path0, rtol=1e-5, atol=1e-5)
    sym = s2 + s2.transpose(1, 2)
    torch.testing.assert_close(sym, s1[:, :, None] * s1[:, None, :], rtol=1e-4, atol=2e-4)


def _torch_signature(path, depth):
    """float64 torch restatement (Chen recursion) with autograd, the gradient oracle."""
    n, length, c = path.shape
    d = path[:, 1:] - path[:, :-1]
    s1 = torch.zeros(n, c, dtype=path.dtype, device=path.device)
    s2 = torch.zeros(n, c, c, dtype=path.dtype, device=path.device)
    s3 = torch.zeros(n, c, c, c, dtype=path.dtype, device=path.device)
    for t in range(length - 1):
        dt = d[:, t]
        if depth >= 3:
            t2 = s2 + (s1 + dt / 3)[:, :, None] * dt[:, None, :] / 2
            s3 = s3 + t2[:, :, :, None] * dt[:, None, None, :]
        if depth >= 2:
            s2 = s2 + (s1 + dt / 2)[:, :, None] * dt[:, None, :]
        s1 = s1 + dt
    parts = [s1, s2.reshape(n, -1), s3.reshape(n, -1)][:depth]
    return torch.cat(parts, dim=1)


@pytest.mark.parametrize('shape', [(7, 21, 3, 1), (40, 21, 4, 1), (33, 6, 5, 2), (3, 11, 15, 6),
                                   (2, 11, 108, 21), (5, 2, 2, 1), (9, 21, 6, 2), (4, 21, 15, 5),
                                   (6, 8, 10, 3)])
def test_signature_backward_vs_float64_autograd(shape):
    """The differentiable summarizer: gradients of a random linear functional of the
    signature wrt states/actions vs float64 autograd of the Chen recursion."""
    from bayes_sim_ig.utils import summarizers as S
    n, t1, d, a = shape
    s, ac = synth_rollouts(21 + n, n, t1, d, a)
    s = (s * 0.3).to(DEV).requires_grad_(True)
    ac = ac.to(DEV).requires_grad_(True)
    out = S.summary_signatory(s, ac)
    w = torch.randn(out.shape, generator=torch.Generator('cpu').manual_seed(3)).to(DEV)
    (out * w).sum().backward()          # depth 3 for every width that gets it (C <= 22)
    s64 = s.detach().double().requires_grad_(True)
    a64 = ac.detach().double().requires_grad_(True)
    tcol = torch.arange(1, t1 + 1, device=DEV, dtype=torch.float64).view(1, -1, 1).repeat(n, 1, 1)
    depth = osum.signature_depth(1 + d + a)
    ref = _torch_signature(torch.cat([tcol, s64, a64], dim=-1), depth)
    assert rel_err(out.detach().cpu(), ref.detach().cpu()) < 2e-5
    (ref * w.double()).sum().backward()
    assert rel_err(s.grad.cpu(), s64.grad.cpu()) < 5e-5
    assert rel_err(ac.grad.cpu(), a64.grad.cpu()) < 5e-5


@pytest.mark.parametrize('shape', [(37, 21, 4, 1), (5, 11, 108, 21), (1000, 51, 60, 8), (3, 7, 3, 2)])
def test_time_major_ingestion_is_bit_identical(shape, capsys):
    """[T, N, dim] buffers (a vectorised simulator's layout) give
    exactly the summaries of the transposed [N, T, dim] rollouts."""
    from bayes_sim_ig.utils import summarizers as S
    n, t1, d, a = shape
    s, ac = synth_rollouts(3, n, t1, d, a, DEV)
    s_tm, a_tm = s.transpose(0, 1).contiguous(), ac.transpose(0, 1).contiguous()
    for fxn in (S.summary_start, S.summary_waypts, S.summary_corr, S.summary_corrdiff):
        if fxn in (S.summary_start, S.summary_waypts) and t1 < 10:
            continue
        assert torch.equal(fxn(s_tm, a_tm, time_major=True), fxn(s, ac)), fxn.__name__
    # actions stored with extra steps
    longer = torch.cat([a_tm, a_tm[:2]], 0).contiguous()
    assert torch.equal(S.summary_corrdiff(s_tm, longer, time_major=True), S.summary_corrdiff(s, ac))


@pytest.mark.parametrize('shape', [(37, 21, 4, 1), (5, 11, 108, 21), (64, 51, 60, 8), (3, 7, 3, 2)])
def test_time_major_ingestion_matches_the_oracle(shape, capsys):
    """Time-major ingestion against the CPU oracle (oracle/summarizers_np.py, pinned to the
    live reference) evaluated on the trajectory-major rollouts: copies and the corr product
    block bit-exact, mean / std to 2e-6, signature to 2e-5 of the largest term."""
    from bayes_sim_ig.utils import summarizers as S
    n, t1, d, a = shape
    s, ac = synth_rollouts(4, n, t1, d, a, DEV)
    s = s * 0.3                                      # keeps the depth-3 signature terms O(1)
    s_np, a_np = s.cpu().numpy(), ac.cpu().numpy()
    s_tm, a_tm = s.transpose(0, 1).contiguous(), ac.transpose(0, 1).contiguous()
    if t1 >= 10:
        for name in ('summary_start', 'summary_waypts'):
            got = getattr(S, name)(s_tm, a_tm, time_major=True).cpu().numpy()
            assert np.array_equal(got, getattr(osum, name)(s_np, a_np)), name
    for name in ('summary_corr', 'summary_corrdiff'):
        got = getattr(S, name)(s_tm, a_tm, time_major=True).cpu().numpy()
        ref = getattr(osum, name)(s_np, a_np)
        assert np.array_equal(got[:, :-2], ref[:, :-2]), name
        np.testing.assert_allclose(got[:, -2:], ref[:, -2:], rtol=2e-6, atol=1e-6)
    if S.signature_depth(1 + d + a) > 0:
        got = S.summary_signatory(s_tm, a_tm, time_major=True).cpu().numpy()
        ref = osum.summary_signatory(s_np, a_np)
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= 2e-5 * max(np.abs(ref).max(), 1.0)
