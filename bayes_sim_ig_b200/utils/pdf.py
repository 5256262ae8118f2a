"""Uniform, Gaussian and mixture-of-Gaussians densities.

Mirror of the reference module ``bayes_sim_ig/utils/pdf.py`` (same class and
method names, argument meaning and error behaviour).  Construction and the
small per-component algebra (inverses, products, pruning) stay on the host in
numpy, in the dtype of the inputs exactly as the reference does (float32 when
the parameters come from ``MDNN.predict_MoGs``).  The per-sample work --
``MoG.gen`` / ``Gaussian.gen`` (component pick + affine map) and
``MoG.eval`` / ``Gaussian.eval`` (log-density of many points) -- runs in the
CUDA kernels of csrc/posterior.cu; there is no CPU path for it.

Random draws keep the reference's generator and order (numpy global RNG:
``rand(n,1)`` then the normals, utils/pdf.py:74,299), so seeding numpy gives
the reference's samples.  ``gen(..., method='philox')`` is an extension that
draws on the device instead (fp32, draw order) for bulk sampling.

Deliberate fixes relative to the reference (dead code there):
``__truediv__`` is defined (the reference only has the Python-2 ``__div__``),
and ``MoG.calc_mean_and_cov`` uses the component covariances (the reference
reads a non-existent ``.sigma`` attribute).
"""
import numpy as np
import numpy.random as rng
import scipy.special
import torch
from scipy.special import erfinv, logsumexp

from .. import _lib
from .halton import halton_points


# --------------------------------------------------------------------------- device glue
def _device():
    if not torch.cuda.is_available():
        raise _lib.BsigError('pdf.gen / pdf.eval need a CUDA device; there is no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def _dev64(arr, dev):
    return torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64)).to(dev)


def _stage_mixture(a, means, cmats, dev):
    """Mixture parameters as device tensors (a in its own dtype: the component choice
    compares float64 uniforms with a float32 cumsum when a is float32)."""
    a = np.ascontiguousarray(a)
    a_is_f32 = a.dtype == np.float32
    if not a_is_f32:
        a = a.astype(np.float64)
    return (torch.from_numpy(a).to(dev), a_is_f32, _dev64(means, dev), _dev64(cmats, dev))


def _mixture_sample_device(a, means, cmats, u, z, want_comp=True, staged=None, device_out=False):
    """a [K] (f32 or f64), means [K,P], cmats [K,P,P], u [n], z [n,P] (f64 draws)
    -> (samples [n,P] f64 grouped by component, comp_idx [n] int32 or None).
    ``staged``: parameters already on the device (_stage_mixture; MoG caches them across
    calls); ``device_out``: leave the samples on the device (torch float64 tensor)."""
    dev = _device()
    n, p = z.shape
    if staged is None:
        staged = _stage_mixture(a, means, cmats, dev)
    a_d, a_is_f32, m_d, c_d = staged
    k = a_d.shape[0]
    u_d, z_d = _dev64(u.reshape(-1), dev), _dev64(z, dev)
    comp = torch.empty(n, dtype=torch.int32, device=dev)
    counts = torch.zeros(k, dtype=torch.int32, device=dev)
    out = torch.empty((n, p), dtype=torch.float64, device=dev)
    _lib.call('bsig_mog_sample', a_d.data_ptr(), 1 if a_is_f32 else 0, u_d.data_ptr(),
              z_d.data_ptr(), m_d.data_ptr(), c_d.data_ptr(), comp.data_ptr(),
              counts.data_ptr(), out.data_ptr(), n, p, k, _lib.stream_ptr(dev))
    return (out if device_out else out.cpu().numpy()), (comp.cpu().numpy() if want_comp else None)


def _mixture_logpdf_device(x, a, means, precs, logdets, log):
    dev = _device()
    x = np.ascontiguousarray(x)
    x_is_f32 = x.dtype == np.float32
    if not x_is_f32:
        x = x.astype(np.float64)
    m, p = x.shape
    k = len(a)
    x_d = torch.from_numpy(x).to(dev)
    out = torch.empty(m, dtype=torch.float64, device=dev)
    # keep the staged parameter tensors alive until the result has been read back
    a = np.asarray(a)
    with np.errstate(divide='ignore'):
        log_a = np.log(a)               # in a's own dtype, like np.log(self.a) at pdf.py:486
    a_d, la_d, m_d, p_d, l_d = (_dev64(t, dev) for t in (a, log_a, means, precs, logdets))
    _lib.call('bsig_mog_logpdf', x_d.data_ptr(), 1 if x_is_f32 else 0, a_d.data_ptr(),
              la_d.data_ptr(), m_d.data_ptr(), p_d.data_ptr(), l_d.data_ptr(), out.data_ptr(), m, p, k,
              1 if log else 0, _lib.stream_ptr(dev))
    return out.cpu().numpy()


def discrete_sample(p, n_samples=1):
    """Reference pdf.py:61-76: index = #{j : u > cumsum(p[:-1])_j}, u ~ rand(n,1).
    The comparison runs on the device (bit-exact for identical uniforms)."""
    p = np.asarray(p)
    u = rng.rand(n_samples, 1)
    k = p.shape[0]
    _, comp = _mixture_sample_device(p, np.zeros((k, 1)), np.zeros((k, 1, 1)),
                                     u, np.zeros((n_samples, 1)))
    return comp.astype(int)


# --------------------------------------------------------------------------- Uniform
class Uniform:
    """Box-uniform prior (reference pdf.py:79-192), including its sampling
    quirks: 'random' concatenates per-dimension draws and reshapes
    row-major; 'halton' uses lb[0] / ub[1] for every dimension."""

    def __init__(self, lb_array=None, ub_array=None):
        assert len(lb_array) == len(ub_array)
        self.lb_array = lb_array
        self.ub_array = ub_array
        self.cur_index = 0
        self.cached_samples = None
        self.param_dim = len(lb_array)

    def __str__(self):
        return 'Uniform: \nlower bounds:\n' + str(self.lb_array) + \
               '\nupper bounds:\n' + str(self.ub_array)

    def generate_halton_samples(self, n_samples=1000):
        lo = np.full(self.param_dim, self.lb_array[0], dtype=np.float64)
        hi = np.full(self.param_dim, self.ub_array[1], dtype=np.float64)
        pts = halton_points(n_samples, self.param_dim)
        return lo + pts * (hi - lo)

    def gen(self, n_samples=1, method='random'):
        if method == 'halton':
            result = self.generate_halton_samples(n_samples=n_samples)
        elif method == 'random':
            cols = [np.random.uniform(self.lb_array[ix], self.ub_array[ix], size=n_samples)
                    for ix in range(len(self.lb_array))]
            result = np.concatenate(cols, axis=0)
        else:
            raise ValueError('Unknown gen method ' + method)
        return result.reshape(-1, len(self.lb_array))

    def eval(self, x, ii=None, log=True, debug=False):
        if ii is None:
            ii = np.arange(self.param_dim)
        npts = np.atleast_2d(x).shape[0]
        dens = np.ones((npts,)) / np.prod(self.ub_array[ii] - self.lb_array[ii])
        inside = (x > self.lb_array[ii]) & (x < self.ub_array[ii])
        dens[np.prod(inside, axis=1) == 0] = 0
        if not log:
            return dens
        if not inside.any():
            raise ValueError('log prob. not defined outside of truncation')
        return np.log(dens)


# --------------------------------------------------------------------------- Gaussian
class Gaussian:
    """Gaussian with mean m, precision P, covariance S, factor C (S = C'C),
    Pm = P m and logdetP (reference pdf.py:195-411).  Valid constructor
    combinations: (m | Pm) with one of P, U (U'U = P), S; or m with the packed
    lower factor L = [diag | strict lower] (L L' = S)."""

    def __init__(self, m=None, P=None, U=None, S=None, Pm=None, L=None):
        if m is None and Pm is None:
            raise ValueError('Mean information missing.')
        first = np.asarray(m if m is not None else Pm)
        self.ndim = first.size
        if P is not None:
            P = np.asarray(P)
            chol = np.linalg.cholesky(P)
            self.P = P
            self.C = np.linalg.inv(chol)
            self.S = np.dot(self.C.T, self.C)
            self.logdetP = 2.0 * np.sum(np.log(np.diagonal(chol)))
        elif U is not None:
            U = np.asarray(U)
            self.P = np.dot(U.T, U)
            self.C = np.linalg.inv(U.T)
            self.S = np.dot(self.C.T, self.C)
            self.logdetP = 2.0 * np.sum(np.log(np.diagonal(U)))
        elif L is not None and m is not None:
            packed = np.asarray(L)
            lower = np.diag(packed[0:self.ndim])
            if 1 < self.ndim < packed.shape[0]:          # full covariance
                rows, cols = np.tril_indices(self.ndim, -1)
                lower[rows, cols] = packed[self.ndim:]
            self.C = lower.T
            self.S = np.dot(self.C.T, self.C)
            self.P = np.linalg.inv(self.S)
            self.logdetP = -2.0 * np.sum(np.log(np.diagonal(self.C)))
        elif S is not None:
            S = np.asarray(S)
            self.P = np.linalg.inv(S)
            self.C = np.linalg.cholesky(S).T
            self.S = S
            self.logdetP = -2.0 * np.sum(np.log(np.diagonal(self.C)))
        else:
            raise ValueError('Precision information missing.')
        if m is not None:
            self.m = first
            self.Pm = np.dot(self.P, self.m)
        else:
            self.Pm = first
            self.m = np.dot(self.S, first) if (S is not None and P is None and U is None) \
                else np.linalg.solve(self.P, first)

    def _adopt(self, other):
        for name in ('m', 'P', 'C', 'S', 'Pm', 'logdetP'):
            setattr(self, name, getattr(other, name))

    def gen(self, n_samples=1, method='random'):
        """Reference pdf.py:296-309: z @ C + m with z ~ randn (or Halton)."""
        n_samples = int(n_samples)
        if method == 'random':
            z = rng.randn(n_samples, self.ndim)
        elif method == 'halton':
            z = erfinv(2 * halton_points(n_samples, self.ndim) - 1) * np.sqrt(2)
        else:
            raise ValueError('Unknown gen method ' + method)
        if n_samples == 0:
            return np.zeros((0, self.ndim))
        out, _ = _mixture_sample_device(np.ones(1), self.m[None, :], self.C[None, :, :],
                                        np.zeros(n_samples), z)
        return out

    def eval(self, x, ii=None, log=True):
        """Reference pdf.py:311-342."""
        x = np.asarray(x)
        if ii is None:
            res = _mixture_logpdf_device(np.atleast_2d(x), np.ones(1), self.m[None, :],
                                         self.P[None, :, :], np.array([self.logdetP]), True)
            res = res.astype(np.result_type(x.dtype, self.m.dtype, self.P.dtype), copy=False)
        else:
            mean, cov = self._marginal(ii)
            res = _mixture_logpdf_device(np.atleast_2d(x), np.ones(1), mean[None, :],
                                         np.linalg.inv(cov)[None, :, :],
                                         np.array([-np.linalg.slogdet(cov)[1]]), True)
        return res if log else np.exp(res)

    def _marginal(self, ii):
        """Mean / jittered covariance of the marginal over dims ii
        (pdf.py:334-337; the jitter consumes numpy's global RNG like the reference)."""
        mean = np.asarray(self.m[ii], dtype=np.float64)
        cov = np.asarray(self.S[ii][:, ii], dtype=np.float64)
        cov = cov + 1.e-5 * cov.mean() * np.diag(np.random.rand(cov.shape[0]))
        return mean, cov

    def __mul__(self, other):
        assert isinstance(other, Gaussian)
        return Gaussian(P=self.P + other.P, Pm=self.Pm + other.Pm)

    def __imul__(self, other):
        assert isinstance(other, Gaussian)
        res = self * other
        self._adopt(res)
        return res

    def __truediv__(self, other):
        """Quotient of Gaussians; may be improper (then cholesky raises)."""
        assert isinstance(other, Gaussian)
        return Gaussian(P=self.P - other.P, Pm=self.Pm - other.Pm)

    __div__ = __truediv__

    def __itruediv__(self, other):
        assert isinstance(other, Gaussian)
        res = self / other
        self._adopt(res)
        return res

    __idiv__ = __itruediv__

    def __pow__(self, power, modulo=None):
        return Gaussian(P=power * self.P, Pm=power * self.Pm)

    def __ipow__(self, power):
        res = self ** power
        self._adopt(res)
        return res

    def kl(self, other):
        """KL(self | other), reference pdf.py:400-411."""
        assert isinstance(other, Gaussian)
        assert self.ndim == other.ndim
        diff = other.m - self.m
        total = np.sum(other.P * self.S) + np.dot(diff, np.dot(other.P, diff)) \
            + self.logdetP - other.logdetP - self.ndim
        return 0.5 * total


# --------------------------------------------------------------------------- MoG
class MoG:
    """Mixture of Gaussians (reference pdf.py:414-582)."""

    def __init__(self, a, ms=None, Ps=None, Us=None, Ss=None, xs=None, Ls=None):
        if ms is not None:
            if Ps is not None:
                self.xs = [Gaussian(m=m, P=P) for m, P in zip(ms, Ps)]
            elif Us is not None:
                self.xs = [Gaussian(m=m, U=U) for m, U in zip(ms, Us)]
            elif Ss is not None:
                self.xs = [Gaussian(m=m, S=S) for m, S in zip(ms, Ss)]
            elif Ls is not None:
                self.xs = [Gaussian(m=m, L=L) for m, L in zip(ms, Ls)]
            else:
                raise ValueError('Precision information missing.')
        elif xs is not None:
            self.xs = xs
        else:
            raise ValueError('Mean information missing.')
        self.a = np.asarray(a)
        self.ndim = self.xs[0].ndim
        self.n_components = len(self.xs)
        self.ncomp = self.n_components

    @property
    def weights(self):
        return self.a

    @property
    def components(self):
        return self.xs

    def gen(self, n_samples=1, method='random', device_out=False):
        """Reference pdf.py:465-472: pick components with discrete_sample, then
        draw each component's block; the result is grouped by component.
        ``device_out=True`` (not in the reference) returns the float64 samples as a CUDA
        tensor instead of copying them to the host (BayesSim.predict's resample + refit)."""
        n_samples = int(n_samples)
        means = np.stack([g.m for g in self.xs])
        cmats = np.stack([g.C for g in self.xs])
        if method == 'philox':
            return self._gen_philox(n_samples, means, cmats)
        u = rng.rand(n_samples, 1)
        if method == 'random':
            z = rng.randn(n_samples, self.ndim)
        elif method == 'halton':
            # the reference restarts the sequence for every component block
            idx = np.sum((u > np.cumsum(self.a[:-1])[np.newaxis, :]).astype(int), axis=1)
            blocks = [erfinv(2 * halton_points(int(np.sum(idx == i)), self.ndim) - 1) * np.sqrt(2)
                      for i in range(self.n_components)]
            z = np.concatenate(blocks, axis=0).reshape(n_samples, self.ndim)
        else:
            raise ValueError('Unknown gen method ' + method)
        if n_samples == 0:
            return np.zeros((0, self.ndim))
        with _lib.nvtx_range('bsig.MoG.gen'):
            samples, _ = _mixture_sample_device(self.a, means, cmats, u, z, want_comp=False,
                                                staged=self._staged(means, cmats),
                                                device_out=device_out)
        return samples

    def _staged(self, means, cmats):
        """Device copies of (a, means, cmats), kept across gen calls while the mixture is
        unchanged (the reference's consumers call gen once per environment reset)."""
        dev = _device()
        key = (dev, self.a.dtype.str, self.a.tobytes(), means.tobytes(), cmats.tobytes())
        cache = getattr(self, '_stage_cache', None)
        if cache is None or cache[0] != key:
            cache = (key, _stage_mixture(self.a, means, cmats, dev))
            self._stage_cache = cache
        return cache[1]

    def gen_per_env(self, n_envs, lows=None, highs=None, method='random', u=None, z=None,
                    return_components=False):
        """One (optionally clipped) draw per environment in a single device call: the
        batched form of ``ParamsGenerator.sample`` (reference sim/params_generator.py:
        115-118 -- ``distr.gen(n_samples=1)[0]`` then ``np.clip``), which the reference
        runs once per environment reset.  Row e equals what ``gen(n_samples=1)[0]`` returns
        for the uniform ``u[e]`` and the normals ``z[e]`` (draw order, not grouped by
        component).  ``method='random'`` draws u = rand(n,1) then z = randn(n,P) from the
        module RNG (or takes them as arguments); ``'philox'`` uses the device RNG (fp32)."""
        n_envs = int(n_envs)
        means = np.stack([g.m for g in self.xs])
        cmats = np.stack([g.C for g in self.xs])
        k, p = means.shape
        assert (lows is None) == (highs is None)
        dev = _device()
        comp = torch.empty(max(n_envs, 1), dtype=torch.int32, device=dev)
        if method == 'philox':
            f32 = lambda arr: torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).to(dev)
            a_d, m_d, c_d = f32(self.a), f32(means), f32(cmats)
            lo_d = f32(lows) if lows is not None else None
            hi_d = f32(highs) if highs is not None else None
            out = torch.empty((n_envs, p), dtype=torch.float32, device=dev)
            seed = int(rng.randint(0, 2**31 - 1))
            _lib.call('bsig_mog_sample_envs_philox', a_d.data_ptr(), m_d.data_ptr(), c_d.data_ptr(),
                      None if lo_d is None else lo_d.data_ptr(),
                      None if hi_d is None else hi_d.data_ptr(), comp.data_ptr(), out.data_ptr(),
                      seed, n_envs, p, k, _lib.stream_ptr(dev))
        elif method == 'random':
            if u is None:
                u = rng.rand(n_envs, 1)
            if z is None:
                z = rng.randn(n_envs, self.ndim)
            a = np.ascontiguousarray(self.a)
            a_is_f32 = a.dtype == np.float32
            if not a_is_f32:
                a = a.astype(np.float64)
            a_d = torch.from_numpy(a).to(dev)
            u_d, z_d = _dev64(np.asarray(u).reshape(-1), dev), _dev64(z, dev)
            m_d, c_d = _dev64(means, dev), _dev64(cmats, dev)
            lo_d = _dev64(lows, dev) if lows is not None else None
            hi_d = _dev64(highs, dev) if highs is not None else None
            out = torch.empty((n_envs, p), dtype=torch.float64, device=dev)
            _lib.call('bsig_mog_sample_envs', a_d.data_ptr(), 1 if a_is_f32 else 0, u_d.data_ptr(),
                      z_d.data_ptr(), m_d.data_ptr(), c_d.data_ptr(),
                      None if lo_d is None else lo_d.data_ptr(),
                      None if hi_d is None else hi_d.data_ptr(), out.data_ptr(), comp.data_ptr(),
                      n_envs, p, k, _lib.stream_ptr(dev))
        else:
            raise ValueError('Unknown gen method ' + method)
        res = out.cpu().numpy()
        if return_components:
            return res, comp[:n_envs].cpu().numpy()
        return res

    def _gen_philox(self, n_samples, means, cmats):
        dev = _device()
        k, p = means.shape
        f32 = lambda arr: torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).to(dev)
        a_d, m_d, c_d = f32(self.a), f32(means), f32(cmats)
        out = torch.empty((n_samples, p), dtype=torch.float32, device=dev)
        seed = int(rng.randint(0, 2**31 - 1))
        _lib.call('bsig_mog_sample_philox', a_d.data_ptr(), m_d.data_ptr(), c_d.data_ptr(), None,
                  out.data_ptr(), seed, n_samples, p, k, _lib.stream_ptr(dev))
        return out.cpu().numpy()

    def eval(self, x, ii=None, log=True, debug=False):
        """Reference pdf.py:474-491: log sum_k a_k N_k(x) (or the density)."""
        x = np.asarray(x)
        x2 = np.atleast_2d(x)
        if ii is None:
            means = np.stack([g.m for g in self.xs])
            precs = np.stack([g.P for g in self.xs])
            logdets = np.array([g.logdetP for g in self.xs])
            res = _mixture_logpdf_device(x2, self.a, means, precs, logdets, log)
            res = res.astype(np.result_type(x.dtype, self.a.dtype, means.dtype), copy=False)
        else:
            marg = [g._marginal(ii) for g in self.xs]
            means = np.stack([mm for mm, _ in marg])
            precs = np.stack([np.linalg.inv(cc) for _, cc in marg])
            logdets = np.array([-np.linalg.slogdet(cc)[1] for _, cc in marg])
            res = _mixture_logpdf_device(x2, self.a, means, precs, logdets, log)
        if debug:
            print('weights\n', self.a, '\nres\n', res)
        return res

    def eval_marginal_grids(self, pairs, lims, nbins=100, log=False):
        """All pairwise 2-D marginals on regular grids in ONE device launch -- the batched
        form of what ``plot_posterior`` does pair by pair (reference utils/plot.py:38-44:
        ``posterior.eval(X.T, ii=dims, log=False)`` on ``np.mgrid[xmin:xmax:nbins*1j,
        ymin:ymax:nbins*1j]``).  ``pairs``: sequence of (i, j) parameter indices; ``lims``:
        one (xmin, xmax, ymin, ymax) for all pairs or one per pair.  Returns
        [len(pairs), nbins, nbins] float64, entry [q, i, j] = density at grid point (i, j).
        The covariance jitter of the marginal branch (pdf.py:336) is drawn from numpy's global
        RNG pair by pair, component by component -- the order a loop of reference calls uses."""
        pairs = [tuple(int(v) for v in pr) for pr in pairs]
        n_pairs, k = len(pairs), self.n_components
        lims = np.asarray(lims, dtype=np.float64)
        if lims.ndim == 1:
            lims = np.tile(lims, (n_pairs, 1))
        assert lims.shape == (n_pairs, 4)
        prm = np.empty((n_pairs, k, 6), dtype=np.float64)
        for q, pr in enumerate(pairs):
            for c, g in enumerate(self.xs):
                mean, cov = g._marginal(list(pr))
                prec = np.linalg.inv(cov)
                prm[q, c] = (mean[0], mean[1], prec[0, 0], 0.5 * (prec[0, 1] + prec[1, 0]),
                             prec[1, 1], -np.linalg.slogdet(cov)[1])
        if n_pairs == 0:
            return np.zeros((0, nbins, nbins))
        dev = _device()
        with np.errstate(divide='ignore'):
            log_a = np.log(np.asarray(self.a))
        a_d, la_d, p_d, l_d = (_dev64(t, dev) for t in (self.a, log_a, prm, lims))
        out = torch.empty((n_pairs, nbins, nbins), dtype=torch.float64, device=dev)
        _lib.call('bsig_mog_marginal_grid', a_d.data_ptr(), la_d.data_ptr(), p_d.data_ptr(),
                  l_d.data_ptr(), out.data_ptr(), n_pairs, k, int(nbins), 1 if log else 0,
                  _lib.stream_ptr(dev))
        return out.cpu().numpy()

    def __str__(self):
        mus = np.array([g.m.tolist() for g in self.xs])
        diag_s = np.array([np.diagonal(g.S).tolist() for g in self.xs])
        return 'MoG:\nweights:\n' + str(self.a) + '\nmeans:\n' + str(mus) + \
               '\ndiagS:\n' + str(diag_s)

    def _reweighted(self, ys, other, sign):
        """Mixture weights after multiplying (sign=+1) / dividing (sign=-1) every
        component by the Gaussian ``other`` (pdf.py:501-515, 525-539)."""
        logc = np.empty_like(self.a)
        quad = lambda g: np.dot(g.m, np.dot(g.P, g.m))
        for i, (x, y) in enumerate(zip(self.xs, ys)):
            val = x.logdetP + sign * other.logdetP - y.logdetP
            val = val - quad(x) + sign * quad(other) - quad(y)
            logc[i] = 0.5 * val
        la = np.log(self.a) + logc
        la -= logsumexp(la)
        return np.exp(la)

    def __mul__(self, other):
        assert isinstance(other, Gaussian)
        ys = [x * other for x in self.xs]
        return MoG(a=self._reweighted(ys, other, +1.0), xs=ys)

    def __imul__(self, other):
        assert isinstance(other, Gaussian)
        res = self * other
        self.a, self.xs = res.a, res.xs
        return res

    def __truediv__(self, other):
        assert isinstance(other, Gaussian)
        ys = [x / other for x in self.xs]
        return MoG(a=self._reweighted(ys, other, -1.0), xs=ys)

    __div__ = __truediv__

    def __itruediv__(self, other):
        assert isinstance(other, Gaussian)
        res = self / other
        self.a, self.xs = res.a, res.xs
        return res

    __idiv__ = __itruediv__

    def calc_mean_and_cov(self):
        """Mean and covariance of the mixture (moment matching)."""
        ms = np.stack([g.m for g in self.xs])
        mean = np.dot(self.a, ms)
        second = sum(w * (g.S + np.outer(g.m, g.m)) for w, g in zip(self.a, self.xs))
        return mean, second - np.outer(mean, mean)

    def project_to_gaussian(self):
        mean, cov = self.calc_mean_and_cov()
        return Gaussian(m=mean, S=cov)

    def prune_negligible_components(self, threshold):
        """Reference pdf.py:562-570: drop components with a < threshold and
        spread their mass equally over the survivors."""
        drop = np.nonzero((self.a < threshold).astype(int))[0]
        removed = np.sum(self.a[drop])
        self.n_components -= drop.size
        self.a = np.delete(self.a, drop)
        self.a += removed / self.n_components
        self.xs = [g for i, g in enumerate(self.xs) if i not in drop]

    def kl(self, other, n_samples=10000):
        """Monte-Carlo KL(self | other), reference pdf.py:572-581."""
        x = self.gen(n_samples)
        diff = self.eval(x, log=True) - other.eval(x, log=True)
        return np.mean(diff), np.std(diff, ddof=1) / np.sqrt(n_samples)


def fit_mog(x, n_components, w=None, tol=1.0e-9, maxiter=float('inf'), verbose=False):
    """EM fit of a mixture to (optionally weighted) data (reference
    pdf.py:584-642; unused by the hot path, kept for API completeness)."""
    from scipy.stats import multivariate_normal as mvn
    x = x[:, np.newaxis] if x.ndim == 1 else x
    n_data, n_dim = x.shape
    a = np.ones(n_components) / n_components
    ms = rng.randn(n_components, n_dim)
    covs = [np.eye(n_dim) for _ in range(n_components)]

    def joint(singular):
        lp = np.array([mvn.logpdf(x, ms[k], covs[k], allow_singular=singular)
                       for k in range(n_components)]).reshape(n_components, n_data)
        return lp + np.log(a)[:, np.newaxis]

    log_pxz = joint(False)
    log_px = logsumexp(log_pxz, axis=0)
    prev = np.mean(log_px) if w is None else np.dot(w, log_px)
    it = 0
    while True:
        resp = np.exp(log_pxz - log_px)
        if w is not None:
            resp = resp * w
        mass = np.sum(resp, axis=1)
        a = mass / n_data if w is None else mass
        ms = np.dot(resp, x) / mass[:, np.newaxis]
        for k in range(n_components):
            xm = x - ms[k]
            covs[k] = np.dot(xm.T * resp[k], xm) / mass[k]
        log_pxz = joint(True)
        log_px = logsumexp(log_pxz, axis=0)
        cur = np.mean(log_px) if w is None else np.dot(w, log_px)
        it += 1
        if verbose:
            print('Iteration = {0}, log likelihood = {1}, diff = {2}'.format(it, cur, cur - prev))
        if cur - prev < tol or it > maxiter:
            break
        prev = cur
    return MoG(a=a, ms=ms, Ss=covs)
