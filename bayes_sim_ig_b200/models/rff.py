"""Random Fourier features on the device.

Mirror of the reference module ``bayes_sim_ig/models/rff.py``: class ``RFF``
with ``to_features`` / ``draw_freqs`` and the four spectral-density kernels.
The frequencies are drawn once on the host exactly as the reference does
(quasi-random inverse-CDF points, or numpy's global RNG); the projection
``x @ (freqs/sigma)^T`` and the cos/sin epilogue run as ONE GEMM kernel with a
fused sincos epilogue (csrc/linear.cu: bsig_rff_features) instead of the
reference's divide + matmul + cos + sin + cat + scale (rff.py:128-132).
"""
import numpy as np
import torch
from scipy.special import erfinv

from .. import _lib
from ..utils.halton import halton_points


class RFF:
    """Random Fourier Features, vanilla or quasi-random (reference rff.py:44-132).
    Make sure the input space is normalised."""

    def __init__(self, n_feat, d, sigma, cos_only=False, quasi_random=True,
                 kernel='RBF', device='cuda'):
        self.n_feat = n_feat
        self.d = int(d)
        self.freqs = None
        self.offset = None
        self.a = 1.0
        self.device = device
        self.cos_only = bool(cos_only)
        self.gemm_engine = -1           # BSIG_GEMM_* selector (-1 = auto)
        dev = _lib.require_cuda(device)
        if isinstance(sigma, list):
            assert (len(sigma) == d)
            sig = np.array(sigma, dtype=np.float32)
        else:
            sig = np.ones(d, dtype=np.float32) * sigma
        self.sigma = torch.from_numpy(sig).float().reshape(1, -1).to(dev)
        if kernel == 'RBF':
            rff_kernel = RFFKernelRBF()
        elif kernel == 'Laplace' or kernel == 'Matern12':
            rff_kernel = RFFKernelMatern12()
        elif kernel == 'Matern32':
            rff_kernel = RFFKernelMatern32()
        elif kernel == 'Matern52':
            rff_kernel = RFFKernelMatern52()
        else:
            raise ValueError('Kernel {} is not recognised.'.format(kernel))
        if cos_only:
            freqs = RFF.draw_freqs(rff_kernel, n_feat, d, quasi_random)
            self.offset = torch.from_numpy(
                2.0 * np.pi * np.random.rand(1, n_feat)).float().to(dev)
            self.a = np.sqrt(1.0 / float(n_feat))
            self.to_features = self._to_cos_only_features
        else:
            assert (self.n_feat % 2 == 0)
            freqs = RFF.draw_freqs(rff_kernel, n_feat // 2, d, quasi_random)
            self.a = np.sqrt(1.0 / float(n_feat / 2))
            self.to_features = self._to_cos_sin_features
        self.freqs = torch.from_numpy(np.asarray(freqs)).float().to(dev)
        self._coeff_key = None
        self._coeff = None

    @staticmethod
    def draw_freqs(rff_kernel, m, d, quasi_random):
        if quasi_random:
            return rff_kernel.inv_cdf(halton_points(m, d))
        return rff_kernel.sample_freqs((m, d))

    def coeff(self, sigma=None):
        """freqs / sigma, cached until either tensor is replaced or edited."""
        sigma = self.sigma if sigma is None else sigma
        key = (self.freqs.data_ptr(), self.freqs._version, sigma.data_ptr(), sigma._version)
        if key != self._coeff_key:
            self._coeff = (self.freqs / sigma).contiguous()
            self._coeff_key = key
        return self._coeff

    def _project(self, x, sigma, rows=None, out=None):
        if not x.is_cuda:
            raise _lib.BsigError('RFF.to_features needs a CUDA tensor; no CPU fallback')
        x = x.float().contiguous() if (x.dtype != torch.float32 or not x.is_contiguous()) else x
        coeff = self.coeff(sigma)
        nf_half = coeff.shape[0]
        m = x.shape[0] if rows is None else rows.shape[0]
        if out is None:
            out = torch.empty((m, 2 * nf_half), dtype=torch.float32, device=x.device)
        ws_bytes = _lib.load().bsig_linear_ws_bytes(m, nf_half, self.d)
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            _lib.call('bsig_rff_features', _lib.ptr(x), x.shape[1],
                      None if rows is None else _lib.ptr(rows, torch.int64), _lib.ptr(coeff),
                      _lib.ptr(out), m, self.d, nf_half, float(self.a), int(self.gemm_engine),
                      ws.data_ptr(), ws.numel(), _lib.stream_ptr(x.device))
        return out

    def _to_cos_only_features(self, x, sigma=None):
        """Reference rff.py:122-126 (unused by MDRFF): a*cos(x coeff^T + offset).
        Built from the cos/sin kernel via cos(t+o) = cos t cos o - sin t sin o."""
        both = self._project(x, sigma)
        half = both.shape[1] // 2
        return both[:, :half] * torch.cos(self.offset) - both[:, half:] * torch.sin(self.offset)

    def _to_cos_sin_features(self, x, sigma=None):
        """Reference rff.py:128-132: a * [cos(x coeff^T) | sin(x coeff^T)]."""
        return self._project(x, sigma)


class RFFKernel:
    def sample_freqs(self, shape):
        raise NotImplementedError

    def inv_cdf(self, x):
        raise NotImplementedError


class RFFKernelRBF(RFFKernel):
    """Gaussian spectral density."""

    def sample_freqs(self, shape):
        return np.random.normal(0.0, 1.0, shape)

    def inv_cdf(self, x):
        return erfinv(2 * x - 1) * np.sqrt(2)


class _StudentT(RFFKernel):
    """Matern-(nu/2) kernels have Student-t spectral densities with nu dof."""
    nu = 1

    def sample_freqs(self, shape):
        return np.random.normal(0, 1, shape) * np.sqrt(self.nu / np.random.chisquare(self.nu, shape))


class RFFKernelMatern12(_StudentT):
    nu = 1

    def inv_cdf(self, x):
        return np.tan(np.pi * (x - 0.5))            # standard Cauchy quantile


class RFFKernelMatern32(_StudentT):
    nu = 3

    def inv_cdf(self, x):
        # closed-form t(nu=2)-style quantile used by the reference (Shaw 2006)
        return (2 * x - 1) / np.sqrt(2 * x * (1 - x))


class RFFKernelMatern52(_StudentT):
    nu = 5

    def inv_cdf(self, x):
        alpha = 4 * x * (1 - x)
        p = 4 * np.cos(np.arccos(np.sqrt(alpha)) / 3) / np.sqrt(alpha)
        return np.sign(x - 0.5) * np.sqrt(p - 4)
