"""GPU parity: mixture sampling / component selection / log-density."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
CASES = ['f32', 'f64', 'p1', 'p13']


def _mog(g, case):
    from bayes_sim_ig.utils import pdf
    return pdf.MoG(a=g[case + '.a'], ms=list(g[case + '.ms']), Ls=list(g[case + '.Ls']))


@pytest.mark.parametrize('case', CASES)
def test_component_selection_bit_exact(golden, case):
    from bayes_sim_ig.utils import pdf
    g = golden('pdf')
    np.random.seed(32)                      # the seed the golden indices were drawn with
    idx = pdf.discrete_sample(g[case + '.a'], 100)
    np.testing.assert_array_equal(idx, g[case + '.discrete.idx'])


@pytest.mark.parametrize('case', CASES)
def test_gen_matches_reference_stream(golden, case):
    g = golden('pdf')
    mog = _mog(g, case)
    np.random.seed(31)
    smp = mog.gen(n_samples=257)
    ref = g[case + '.gen.samples']
    assert smp.shape == ref.shape and smp.dtype == np.float64
    # float64 affine map: only the summation order of the dot product differs
    np.testing.assert_allclose(smp, ref, rtol=1e-12, atol=1e-12)
    assert mog.gen(n_samples=1).shape == (1, mog.ndim)
    assert mog.gen(n_samples=0).shape == (0, mog.ndim)


@pytest.mark.parametrize('case', CASES)
def test_eval_matches_reference(golden, case):
    g = golden('pdf')
    mog = _mog(g, case)
    x64 = g[case + '.eval.x64']
    np.testing.assert_allclose(mog.eval(x64, log=True), g[case + '.eval.log64'], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(mog.eval(x64, log=False), g[case + '.eval.lin64'], rtol=1e-9, atol=1e-300)
    got32 = mog.eval(x64.astype(np.float32), log=True)
    assert got32.dtype == g[case + '.eval.log32'].dtype
    np.testing.assert_allclose(got32, g[case + '.eval.log32'], rtol=2e-4, atol=2e-4)
    # single Gaussian == one-component mixture
    one = mog.xs[0].eval(x64)
    from bayes_sim_ig.utils import pdf
    solo = pdf.MoG(a=np.ones(1), xs=[mog.xs[0]])
    np.testing.assert_allclose(one, solo.eval(x64), rtol=1e-12)


def test_eval_with_zero_weight_components():
    """A component with weight exactly 0 (log a = -inf; arises when MoG.__mul__ /
    __truediv__ underflow a float32 weight) must be ignored as scipy's logsumexp does
    (pdf.py:489), wherever it sits in the mixture -- a leading one used to give NaN."""
    from bayes_sim_ig.utils import pdf
    from oracle import pdf_np
    rs = np.random.RandomState(5)
    p, k = 3, 4
    ms = [rs.randn(p) for _ in range(k)]
    ls = [np.concatenate([0.5 + rs.rand(p), 0.1 * rs.randn(p * (p - 1) // 2)]) for _ in range(k)]
    x = rs.randn(64, p)
    for a in ([0.0, 0.2, 0.3, 0.5], [0.4, 0.0, 0.6, 0.0], [0.0, 0.0, 0.0, 1.0]):
        mog = pdf.MoG(a=np.array(a), ms=ms, Ls=ls)
        precs = np.stack([g.P for g in mog.xs])
        logdets = np.array([g.logdetP for g in mog.xs])
        with np.errstate(divide='ignore'):
            ref = pdf_np.mog_logpdf(x, np.array(a), np.stack(ms), precs, logdets, log=True)
        got = mog.eval(x, log=True)
        assert np.isfinite(got).all()
        np.testing.assert_allclose(got, ref, rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(mog.eval(x, log=False), np.exp(ref), rtol=1e-9)


def test_sampling_moments_large_n(golden):
    """Size-independent property: 200k samples reproduce mixture mean/cov."""
    g = golden('pdf')
    mog = _mog(g, 'f64')
    np.random.seed(5)
    smp = mog.gen(n_samples=200000)
    mean, cov = mog.calc_mean_and_cov()
    np.testing.assert_allclose(smp.mean(0), mean, atol=0.02)
    np.testing.assert_allclose(np.cov(smp.T), cov, atol=0.05)
    dev = mog.gen(n_samples=200000, method='philox')
    assert dev.dtype == np.float32
    np.testing.assert_allclose(dev.mean(0), mean, atol=0.02)
    np.testing.assert_allclose(np.cov(dev.T), cov, atol=0.05)


def test_marginal_eval_close_to_joint_marginalisation(golden):
    g = golden('pdf')
    mog = _mog(g, 'f64')
    x = np.random.RandomState(1).randn(50, 2)
    np.random.seed(0)
    lp = mog.eval(x, ii=[0, 2], log=True)
    from scipy.stats import multivariate_normal as mvn
    ref = np.log(sum(a * mvn.pdf(x, c.m[[0, 2]], c.S[[0, 2]][:, [0, 2]]) for a, c in zip(mog.a, mog.xs)))
    np.testing.assert_allclose(lp, ref, rtol=1e-3, atol=1e-3)   # reference adds 1e-5 jitter


@pytest.mark.parametrize('case', CASES)
def test_batched_params_sampling_matches_per_env_reference(golden, case):
    """one device launch == ParamsGenerator.sample called once per
    environment in the live reference (golden: 41 environments, clipping active)."""
    from bayes_sim_ig.sim.params_generator import ParamsSampler
    from oracle import pdf_np
    g = golden('pdf')
    mog = _mog(g, case)
    lows, highs = g[case + '.envs.lows'], g[case + '.envs.highs']
    sampler = ParamsSampler(lows, highs)
    sampler.set_distr(mog)
    got = sampler.sample_batch(41, u=g[case + '.envs.u'], z=g[case + '.envs.z'])
    ref = g[case + '.envs.samples']
    assert got.shape == ref.shape and got.dtype == np.float64
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12)
    # clipped entries and the component choice are exact
    assert np.array_equal(got == highs, ref == highs) and np.array_equal(got == lows, ref == lows)
    _, comp = mog.gen_per_env(41, u=g[case + '.envs.u'], z=g[case + '.envs.z'],
                              return_components=True)
    np.testing.assert_array_equal(comp, pdf_np.discrete_sample_from_u(g[case + '.a'],
                                                                      g[case + '.envs.u']))
    # reference-semantics single draw, and the same RNG stream for the batched form
    np.random.seed(5)
    one = sampler.sample()
    assert one.shape == (mog.ndim,) and (one >= lows).all() and (one <= highs).all()
    np.random.seed(6)
    a1 = sampler.sample_batch(1000)
    np.random.seed(6)
    u, z = np.random.rand(1000, 1), np.random.randn(1000, mog.ndim)
    a2 = sampler.sample_batch(1000, u=u, z=z)
    np.testing.assert_array_equal(a1, a2)
    # device RNG: moments of the unclipped draws match the mixture
    free = ParamsSampler(np.full(mog.ndim, -1e9), np.full(mog.ndim, 1e9), mog)
    big = free.sample_batch(200000, method='philox')
    mean, cov = mog.calc_mean_and_cov()
    assert np.abs(big.mean(0) - mean).max() < 0.02
    assert np.abs(np.cov(big.T).reshape(cov.shape) - cov).max() < 0.05


def test_batched_pairwise_marginal_grids_match_reference(golden):
    """every pairwise 2-D marginal a posterior plot evaluates
    (utils/plot.py:38-44), all pairs in ONE launch, against the live reference called pair by
    pair (tests/golden/pdf_marginal.npz): same jitter draws (numpy seeded as the reference
    loop was), densities and log densities to 1e-9."""
    from bayes_sim_ig.utils import pdf
    g = golden('pdf_marginal')
    mog = pdf.MoG(a=g['a'], ms=list(g['ms']), Ls=list(g['Ls']))
    pairs, lims, nbins = [tuple(pr) for pr in g['pairs']], g['lims'], int(g['nbins'])
    np.random.seed(91)
    dens = mog.eval_marginal_grids(pairs, lims, nbins=nbins, log=False)
    assert dens.shape == g['density'].shape and dens.dtype == np.float64
    np.testing.assert_allclose(dens, g['density'], rtol=1e-9, atol=1e-300)
    np.random.seed(92)
    logd = mog.eval_marginal_grids(pairs, lims, nbins=nbins, log=True)
    np.testing.assert_allclose(logd, g['logdensity'], rtol=1e-9, atol=1e-9)
    # one limit tuple for all pairs; the pair-by-pair path of MoG.eval gives the same grid
    np.random.seed(5)
    one = mog.eval_marginal_grids(pairs[:2], (0.0, 1.5, 0.0, 1.5), nbins=8)
    np.random.seed(5)
    xi, yi = np.mgrid[0.0:1.5:8j, 0.0:1.5:8j]
    pts = np.stack([xi.ravel(), yi.ravel()], axis=1)
    for q in range(2):
        np.testing.assert_allclose(one[q].ravel(), mog.eval(pts, ii=list(pairs[q]), log=False),
                                   rtol=1e-9)
