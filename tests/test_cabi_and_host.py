"""CPU-side checks: the C ABI library loads and exports every symbol declared in
include/bsig.h, and the host-side logic (no kernel calls) matches the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import summarizers_np as osum

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'bsig.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(bsig_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from bayes_sim_ig_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), 'run python -m bayes_sim_ig_b200.build'
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), name
    # and the ctypes table covers exactly the header
    assert sorted(_lib.SIGNATURES) == declared
    assert _lib.load().bsig_version() == 100


def test_no_cpu_fallback():
    import torch
    from bayes_sim_ig_b200 import _lib
    from bayes_sim_ig_b200.utils import summarizers
    with pytest.raises(_lib.BsigError):
        summarizers.summary_start(torch.zeros(2, 12, 3), torch.zeros(2, 12, 1))
    with pytest.raises(_lib.BsigError):
        _lib.require_cuda('cpu')


def test_drop_in_import_paths():
    import bayes_sim_ig.bayes_sim as bs
    import bayes_sim_ig.models.mdnn as mdnn
    import bayes_sim_ig.models.mdrff as mdrff
    import bayes_sim_ig.models.rff as rff
    import bayes_sim_ig.utils.pdf as pdf
    import bayes_sim_ig.utils.summarizers as summ
    assert bs.BayesSim.NUM_GRAD_UPDATES == 100 and bs.BayesSim.MINIBATCH_SIZE == 100
    assert bs.BayesSim.TEST_FRACTION == 0.2 and bs.BayesSim.NUM_TRAIN_TRAJ_PER_BATCH == 1000
    assert mdnn.MDNN.LL_LIMIT == 1e5 and mdnn.MDNN.MIN_WEIGHT == 1e-5 and mdnn.MDNN.EPS_NOISE == 1e-5
    assert issubclass(mdrff.MDRFF, mdnn.MDNN)
    for name in ('pad_states_actions', 'summary_start', 'summary_waypts', 'cross_correlation',
                 'summary_corr', 'summary_corrdiff', 'signature_depth', 'summary_signatory'):
        assert callable(getattr(summ, name)) and callable(getattr(bs, name))
    for name in ('MoG', 'Gaussian', 'Uniform', 'discrete_sample', 'fit_mog'):
        assert hasattr(pdf, name)
    for name in ('RFF', 'RFFKernelRBF', 'RFFKernelMatern12', 'RFFKernelMatern32', 'RFFKernelMatern52'):
        assert hasattr(rff, name)


def test_summary_width_matches_oracle(golden):
    from bayes_sim_ig_b200.utils.summarizers import signature_depth, summary_width
    g = golden('summarizers')
    for case in ('pendulum', 'cartpole', 'ant', 'humanoid', 'exact10'):
        s, a = g[case + '.states'], g[case + '.actions']
        t, d, ad = s.shape[1], s.shape[2], a.shape[2]
        assert summary_width('summary_start', t, d, ad) == g[case + '.summary_start'].shape[1]
        assert summary_width('summary_corrdiff', t, d, ad) == g[case + '.summary_corrdiff'].shape[1]
        assert summary_width('summary_signatory', t, d, ad) == osum.summary_signatory(s[:1], a[:1]).shape[1]
    assert [signature_depth(int(c)) for c in g['signature_depth.in']] == list(g['signature_depth.out'])


def test_log_steps_match_reference_rule():
    from bayes_sim_ig_b200.models.train_engine import log_steps
    assert log_steps(100) == [0, 20, 40, 60, 80, 99]
    assert log_steps(20) == [0, 4, 8, 12, 16, 19]
    assert log_steps(500) == [0, 100, 200, 300, 400, 499]
    assert log_steps(3) == [0, 1, 2]
    assert log_steps(1) == [0]


def test_minibatch_index_stream_matches_reference_pattern():
    # one randint(0, n, B) per update == the reference's generator (mdnn.py:221)
    # the engine draws all updates with ONE call; the reference draws one call per
    # update: same values, same generator state afterwards
    for n in (800, 76, 5, 1):
        np.random.seed(3)
        a = np.stack([np.random.randint(0, n, 100) for _ in range(7)])
        tail_a = np.random.rand(3)
        np.random.seed(3)
        b = np.random.randint(0, n, (7, 100))
        tail_b = np.random.rand(3)
        np.testing.assert_array_equal(a, b)
        np.testing.assert_array_equal(tail_a, tail_b)


PDF_CASES = ['f32', 'f64', 'p1', 'p13']


@pytest.mark.parametrize('case', PDF_CASES)
def test_pdf_host_algebra_matches_reference(golden, case):
    """Gaussian(m, L=) / MoG construction is host numpy: bit-exact vs the reference."""
    from bayes_sim_ig_b200.utils import pdf
    g = golden('pdf')
    mog = pdf.MoG(a=g[case + '.a'], ms=list(g[case + '.ms']), Ls=list(g[case + '.Ls']))
    assert mog.n_components == len(g[case + '.a']) and mog.ndim == g[case + '.ms'].shape[1]
    for fld in ('C', 'S', 'P', 'Pm'):
        got = np.stack([getattr(x, fld) for x in mog.xs])
        assert got.dtype == g[case + '.' + fld].dtype
        np.testing.assert_array_equal(got, g[case + '.' + fld])
    np.testing.assert_array_equal(np.array([x.logdetP for x in mog.xs]), g[case + '.logdetP'])


def test_pdf_prune_and_algebra(golden):
    from bayes_sim_ig_b200.utils import pdf
    g = golden('pdf')
    mog = pdf.MoG(a=g['prune.a_in'], ms=[np.zeros(2)] * 4, Ls=[np.ones(2)] * 4)
    mog.prune_negligible_components(threshold=0.005)
    np.testing.assert_array_equal(mog.a, g['prune.a_out'])
    assert mog.n_components == 2 and len(mog.xs) == 2
    # product / quotient round trip (the reference's py2-only __div__, fixed here)
    ga = pdf.Gaussian(m=np.array([0.3, -0.2]), S=np.array([[0.5, 0.1], [0.1, 0.4]]))
    gb = pdf.Gaussian(m=np.array([0.1, 0.4]), S=np.array([[2.0, 0.0], [0.0, 3.0]]))
    back = (ga * gb) / gb
    np.testing.assert_allclose(back.m, ga.m, atol=1e-12)
    np.testing.assert_allclose(back.S, ga.S, atol=1e-12)
    assert ga.kl(ga) == pytest.approx(0.0, abs=1e-12)
    with pytest.raises(ValueError):
        pdf.Gaussian(P=np.eye(2))
    with pytest.raises(ValueError):
        pdf.MoG(a=[1.0])
    u = pdf.Uniform(np.array([0.0, 1.0]), np.array([1.0, 3.0]))
    assert np.allclose(u.eval(np.array([[0.5, 2.0]]), log=False), 0.5)
    with pytest.raises(ValueError):
        u.eval(np.array([[5.0, 5.0]]))


def test_halton_points_in_open_unit_cube():
    from bayes_sim_ig_b200.utils.halton import halton_points
    pts = halton_points(100, 50)
    assert pts.shape == (100, 50) and (pts > 0).all() and (pts < 1).all()
    assert abs(pts.mean() - 0.5) < 0.05


def test_next_row_entry_points_have_no_cpu_fallback_either(golden):
    """Time-major ingestion and the batched per-environment sampler are
    device paths like everything else: CPU tensors / a machine without CUDA raise."""
    import torch
    from bayes_sim_ig_b200 import _lib
    from bayes_sim_ig.sim.params_generator import ParamsSampler
    from bayes_sim_ig.utils import pdf, summarizers
    with pytest.raises(_lib.BsigError):
        summarizers.summary_start(torch.zeros(12, 2, 3), torch.zeros(12, 2, 1), time_major=True)
    with pytest.raises(_lib.BsigError):
        summarizers.summary_corrdiff(torch.zeros(12, 2, 3), torch.zeros(12, 2, 1), time_major=True)
    if not torch.cuda.is_available():
        g = golden('pdf')
        mog = pdf.MoG(a=g['f32.a'], ms=list(g['f32.ms']), Ls=list(g['f32.Ls']))
        sampler = ParamsSampler(g['f32.envs.lows'], g['f32.envs.highs'], mog)
        with pytest.raises(_lib.BsigError):
            sampler.sample_batch(4)


def test_params_sampler_host_semantics():
    """ParamsSampler.sample is the reference's three lines (params_generator.py:113-117);
    sample_batch falls back to the distribution's own gen when it has no device sampler."""
    from bayes_sim_ig.sim.params_generator import ParamsSampler

    class Stub(object):
        def __init__(self):
            self.calls = []

        def gen(self, n_samples=1):
            self.calls.append(n_samples)
            return np.tile(np.array([[-3.0, 0.25, 7.0]]), (n_samples, 1))

    stub = Stub()
    sampler = ParamsSampler([-1.0, 0.0, 0.0], [1.0, 1.0, 2.0])
    sampler.set_distr(stub)
    np.testing.assert_array_equal(sampler.sample(), [-1.0, 0.25, 2.0])
    out = sampler.sample_batch(5)
    assert out.shape == (5, 3) and stub.calls == [1, 5]
    np.testing.assert_array_equal(out[3], [-1.0, 0.25, 2.0])
    np.testing.assert_array_equal(sampler.lows, [-1.0, 0.0, 0.0])


def test_capture_guard_restores_gc_state():
    import gc
    from bayes_sim_ig_b200.models.train_engine import _quiet_gc
    assert gc.isenabled()
    with _quiet_gc():
        assert not gc.isenabled()
    assert gc.isenabled()
    gc.disable()
    try:
        with _quiet_gc():
            assert not gc.isenabled()
        assert not gc.isenabled()
    finally:
        gc.enable()



def test_host_pdf_surface_matches_live_reference(golden):
    """Uniform, Gaussian algebra / KL, MoG x Gaussian product and quotient, fit_mog: host
    (numpy) code of the drop-in surface against vectors recorded from the live
    reference (tests/golden/pdf_host.npz)."""
    import warnings
    from bayes_sim_ig.utils import pdf
    g = golden('pdf_host')
    tol = dict(rtol=1e-10, atol=1e-12)
    # ---- Uniform: same numpy stream, same (scrambled, Q8) layout, same densities
    uni = pdf.Uniform(g['uni.lb'], g['uni.ub'])
    np.random.seed(41)
    np.testing.assert_array_equal(uni.gen(n_samples=7), g['uni.gen'])
    np.testing.assert_allclose(uni.eval(g['uni.x'], log=False), g['uni.eval_lin'], **tol)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        np.testing.assert_allclose(uni.eval(g['uni.x'], log=True), g['uni.eval_log'], **tol)
    np.testing.assert_allclose(uni.eval(g['uni.x'][:, [0, 2]], ii=[0, 2], log=False),
                               g['uni.eval_marg'], **tol)
    # ---- Gaussian algebra
    g1 = pdf.Gaussian(m=g['g.m1'], S=g['g.S1'])
    g2 = pdf.Gaussian(m=g['g.m2'], S=g['g.S2'])
    for name, res in (('mul', g1 * g2), ('div', g1 / g2), ('pow', g1 ** 2.5)):
        np.testing.assert_allclose(res.m, g['g.%s.m' % name], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(res.S, g['g.%s.S' % name], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(res.logdetP, g['g.%s.logdetP' % name], rtol=1e-10)
    np.testing.assert_allclose(g1.kl(g2), g['g.kl'], rtol=1e-10)
    # ---- MoG x Gaussian, MoG / Gaussian (the reference's py2 __div__)
    mog = pdf.MoG(a=g['mog.a'], ms=list(g['mog.ms']), Ss=list(g['mog.Ss']))
    for name, res in (('mul', mog * g2), ('div', mog / g2)):
        np.testing.assert_allclose(res.a, g['mog.%s.a' % name], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(np.stack([x.m for x in res.xs]), g['mog.%s.ms' % name],
                                   rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(np.stack([x.S for x in res.xs]), g['mog.%s.Ss' % name],
                                   rtol=1e-8, atol=1e-10)
    # calc_mean_and_cov / project_to_gaussian raise in the reference (pdf.py:553 reads a
    # non-existent attribute); here they work: check the moment identities instead
    mean, cov = mog.calc_mean_and_cov()
    np.testing.assert_allclose(mean, np.dot(g['mog.a'], g['mog.ms']), **tol)
    second = sum(a * (s + np.outer(m, m)) for a, m, s in zip(g['mog.a'], g['mog.ms'], g['mog.Ss']))
    np.testing.assert_allclose(cov, second - np.outer(mean, mean), rtol=1e-9, atol=1e-11)
    pg = mog.project_to_gaussian()
    np.testing.assert_allclose(pg.m, mean, **tol)
    np.testing.assert_allclose(pg.S, cov, rtol=1e-9, atol=1e-11)
    # ---- fit_mog: EM from the same seeded start on the same data
    np.random.seed(43)
    fit = pdf.fit_mog(g['fit.x'], n_components=2, tol=1e-7, maxiter=200)
    np.testing.assert_allclose(fit.a, g['fit.a'], rtol=1e-7)
    np.testing.assert_allclose(np.stack([x.m for x in fit.xs]), g['fit.ms'], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(np.stack([x.S for x in fit.xs]), g['fit.Ss'], rtol=1e-6, atol=1e-8)


def test_rff_kernel_classes_match_live_reference(golden):
    """Row a7: RFFKernel*.sample_freqs consume numpy's global stream exactly like the
    reference, inv_cdf agrees, and RFF.draw_freqs(quasi_random=False) is the same draw."""
    from bayes_sim_ig.models import rff
    g = golden('rff_host')
    for name in ('RFFKernelRBF', 'RFFKernelMatern12', 'RFFKernelMatern32', 'RFFKernelMatern52'):
        kern = getattr(rff, name)()
        np.random.seed(51)
        np.testing.assert_array_equal(kern.sample_freqs((5, 4)), g[name + '.sample'])
        np.testing.assert_allclose(kern.inv_cdf(g['u']), g[name + '.inv_cdf'], rtol=1e-12, atol=1e-14)
        np.random.seed(52)
        np.testing.assert_array_equal(rff.RFF.draw_freqs(kern, 6, 3, False), g[name + '.draw'])
    import torch
    if torch.cuda.is_available():          # the constructor rejects non-CUDA devices first
        with pytest.raises(ValueError):    # reference rff.py:96
            rff.RFF(8, 3, 1.0, kernel='nope', device='cuda')


def test_packed_lane_group_reduction_scheme_for_every_width():
    """csrc/mdn.cu LaneGroup<0>: groups of exactly K lanes packed floor(32/K) to a warp are
    all-reduced with cyclic rotations -- window sums of width 1, 2, 4, ... combined along the
    binary digits of K, then lane 0's total broadcast.  The GPU tests exercise K = 3, 4, 10,
    32; this replays the same index arithmetic for every K in 1..32 on the host."""
    rs = np.random.RandomState(0)
    for k in range(1, 33):
        rpw = 32 // k
        vals = rs.randn(32)
        lane = np.arange(32)
        gi = lane // k
        lane_g = lane - gi * k
        active = gi < rpw
        base = lane - lane_g

        def rot(j):
            t = lane_g + j
            t = np.where(t >= k, t - k, t)
            return np.where(active, base + t, base + lane_g)

        def shfl(x, src):
            return x[src]

        w, tot, off = vals.copy(), np.zeros(32), 0
        width = 1
        while width <= k:
            if k & width:
                tot = tot + shfl(w, rot(off)) if off else w.copy()
                off += width
            if 2 * width <= k:
                w = w + shfl(w, rot(width))
            width <<= 1
        tot = shfl(tot, base)
        mx = vals.copy()
        width = 1
        while width < k:
            mx = np.maximum(mx, shfl(mx, rot(width)))
            width <<= 1
        for g in range(rpw):
            grp = vals[g * k:(g + 1) * k]
            np.testing.assert_allclose(tot[g * k:(g + 1) * k], grp.sum(), rtol=1e-12, atol=1e-12)
            assert len(set(tot[g * k:(g + 1) * k].tolist())) == 1      # identical bits per group
            np.testing.assert_array_equal(mx[g * k:(g + 1) * k], grp.max())


def test_corr_factors_object_expands_to_the_oracle_summary():
    """Host side of the fused first layer: a CorrFactors object built from
    the oracle's own factors must expand (materialize) to the oracle's cross-correlation summary
    (reference summarizers.py:106-119), report its shape, support the row slicing run_training's
    ordered 80/20 split uses, and refuse column indexing."""
    import torch
    from oracle import summarizers_np as osum
    from bayes_sim_ig_b200.utils.summarizers import CorrFactors
    rs = np.random.RandomState(0)
    n, t1, d, a = 9, 12, 5, 2
    states = rs.randn(n, t1, d).astype(np.float32)
    actions = rs.rand(n, t1, a).astype(np.float32)
    ref = osum.cross_correlation(states, actions, use_state_diff=True)
    w = 10
    sf = (states[:, :w, 1:] - states[:, :w, :-1]).reshape(n, -1)
    af = actions[:, :w, :].reshape(n, -1)
    s, q = sf.shape[1], af.shape[1]
    ldf = (s + q + 2 + 3) // 4 * 4
    fac = np.zeros((n, ldf), np.float32)
    fac[:, :s], fac[:, s:s + q] = sf, af
    fac[:, s + q:s + q + 2] = ref[:, -2:]
    cf = CorrFactors(torch.from_numpy(fac), s, q)
    assert cf.shape == ref.shape and len(cf) == n
    assert np.array_equal(cf.materialize().numpy(), ref)
    head, tail = cf[:7], cf[7:]
    assert head.shape == (7, ref.shape[1]) and tail.shape == (2, ref.shape[1])
    assert np.array_equal(tail.materialize().numpy(), ref[7:])
    with pytest.raises(TypeError):
        cf[:, :3]


def test_sharded_exchange_arithmetic_matches_allreduce_then_adam():
    """The sharded data-parallel exchange (reduce-scatter -> Adam on the owned slice -> all-gather
    of the weights, train_engine.replay_dp) restated with numpy: for any world size that divides
    the padded buffer it must give the parameters of all-reduce + full Adam (oracle adam_step)."""
    from oracle import mdn_np
    rs = np.random.RandomState(1)
    for world in (2, 4, 8):
        n = 4 * world * 13
        p0 = rs.randn(n)
        grads = [rs.randn(n) for _ in range(world)]
        full = {'p': p0.copy()}
        m, v = {'p': np.zeros(n)}, {'p': np.zeros(n)}
        mdn_np.adam_step(full, {'p': sum(grads) / world}, m, v, 1, 1e-3)
        shard = n // world
        out = np.empty(n)
        for r in range(world):                       # what rank r computes and all-gathers
            sl = slice(r * shard, (r + 1) * shard)
            g_r = sum(g[sl] for g in grads) / world  # its slice of the reduce-scatter
            pr = {'p': p0[sl].copy()}
            mr, vr = {'p': np.zeros(shard)}, {'p': np.zeros(shard)}
            mdn_np.adam_step(pr, {'p': g_r}, mr, vr, 1, 1e-3)
            out[sl] = pr['p']
        assert np.allclose(out, full['p'], rtol=0, atol=1e-15)
