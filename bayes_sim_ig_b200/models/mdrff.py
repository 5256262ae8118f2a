"""Mixture Density model on random Fourier features (reference models/mdrff.py).

``MDRFF`` is an ``MDNN`` without hidden layers whose input is first mapped by
fixed random Fourier features (``self.rff``).  In training the gather of the
minibatch rows, the projection GEMM and the cos/sin epilogue are one kernel.
"""
from .mdnn import MDNN
from .rff import RFF


class MDRFF(MDNN):
    def __init__(self, input_dim, output_dim, output_lows, output_highs,
                 n_gaussians, lr, activation, full_covariance, device='cuda',
                 n_feat=500, kernel='RBF', sigma=1.0, **kwargs):
        super().__init__(n_feat, output_dim, output_lows, output_highs,
                         n_gaussians, hidden_layers=[], lr=lr,
                         activation=activation, full_covariance=full_covariance,
                         device=device)
        self.rff = RFF(n_feat, input_dim, sigma, cos_only=False,
                       quasi_random=False if input_dim > 100 else True,
                       kernel=kernel, device=device)
        self.raw_input_dim = input_dim
        print('MDRFF n_feat', n_feat, 'sigma', sigma)
        print(self)

    def _features(self, x):
        return self.rff.to_features(x)
