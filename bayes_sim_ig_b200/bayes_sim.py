"""The BayesSim facade on the B200 kernels (reference bayes_sim_ig/bayes_sim.py).

Same constructor, class constants, ``run_training`` and ``predict`` as the
reference; the summarizer and model named in ``model_cfg`` are looked up by name
among this package's implementations.
"""
import os

import numpy as np
import torch

from . import _lib
from .models.mdnn import MDNN
from .models.mdrff import MDRFF
from .utils import pdf
from .utils import summarizers as _summarizers
from .utils.summarizers import *  # noqa: F401,F403  (same names as the reference exposes)

_MODEL_CLASSES = {'MDNN': MDNN, 'MDRFF': MDRFF}


class BayesSim(object):
    NUM_TRAIN_TRAJ_PER_BATCH = 1000  # num trajs for each training batch
    NUM_TRAIN_EPOCHS = 10            # num times to go over the batch
    MINIBATCH_SIZE = 100             # minibatch size for NN training
    NUM_GRAD_UPDATES = NUM_TRAIN_EPOCHS*NUM_TRAIN_TRAJ_PER_BATCH//MINIBATCH_SIZE
    TEST_FRACTION = 0.2              # fraction of dataset to use as test

    def __init__(self, model_cfg, obs_dim, act_dim, params_dim, params_lows, params_highs,
                 prior, proposal=None, device='cuda'):
        """Reference bayes_sim.py:27-82."""
        _lib.require_cuda(device)
        self.prior = prior
        self.proposal = proposal
        model_class = model_cfg['modelClass']
        name = model_cfg['summarizerFxn']
        if name not in _summarizers.__all__ or not name.startswith(('summary_', 'cross_')):
            raise NameError("name '%s' is not defined" % name)
        self.summarizer_fxn = getattr(_summarizers, name)
        # the reference sizes the model by running the summarizer on zeros of
        # length trainTrajLen (bayes_sim.py:57-60); the width is known in closed form
        traj_summaries_dim = _summarizers.summary_width(
            name, model_cfg['trainTrajLen'], obs_dim, act_dim)
        full_covariance = False
        if 'fullCovariance' in model_cfg:
            full_covariance = model_cfg['fullCovariance']
        kwargs = {'input_dim': traj_summaries_dim, 'output_dim': params_dim,
                  'output_lows': params_lows, 'output_highs': params_highs,
                  'n_gaussians': model_cfg['components'],
                  'hidden_layers': model_cfg['hiddenLayers'],
                  'lr': model_cfg['lr'],
                  'activation': torch.nn.Tanh,
                  'full_covariance': full_covariance,
                  'device': device}
        if model_class.startswith('MDRFF'):
            kernel = 'RBF'
            sigma = 4.0
            if '_' in model_class:
                model_params = model_class.split('_')
                model_class = model_params[0]
                kernel = model_params[1]
                if len(model_params) > 2:
                    sigma = float(model_params[2])
            kwargs.update({'n_feat': 200, 'sigma': sigma, 'kernel': kernel})
        if model_class not in _MODEL_CLASSES:
            raise NameError("name '%s' is not defined" % model_class)
        self.model = _MODEL_CLASSES[model_class](**kwargs)

    @staticmethod
    def get_n_trajs_per_batch(n_train_trajs, n_train_trajs_done):
        n_trajs_per_batch = BayesSim.NUM_TRAIN_TRAJ_PER_BATCH
        if n_train_trajs_done + n_trajs_per_batch > n_train_trajs:
            n_trajs_per_batch = n_train_trajs - n_train_trajs_done
        return n_trajs_per_batch

    def run_training(self, params, traj_states, traj_actions):
        """Reference bayes_sim.py:91-114: summarize, then NUM_GRAD_UPDATES Adam
        steps of MINIBATCH_SIZE on the summaries."""
        dev = self.model.flat_params.device
        with _lib.nvtx_range('bsig.summarize'):
            traj_summaries = self._training_summaries(traj_states.to(dev), traj_actions.to(dev))
        with _lib.nvtx_range('bsig.run_training'):
            log_dict = self.model.run_training(
                x_data=traj_summaries, y_data=params,
                n_updates=BayesSim.NUM_GRAD_UPDATES,
                batch_size=BayesSim.MINIBATCH_SIZE,
                test_frac=BayesSim.TEST_FRACTION)
        return log_dict

    # Cross-correlation summaries at least this wide are handed to the MDNN in factored form
    # (SURVEY 8.f rank 1): the [N, F] tensor is never materialised, the first layer's GEMMs
    # generate it tile by tile (csrc/corr_layer.cu).  BSIG_FUSED_CORR=0 switches it off.
    FUSED_CORR_MIN_WIDTH = 4096

    def _training_summaries(self, states, actions):
        name = self.summarizer_fxn.__name__
        width = getattr(self.model, 'raw_input_dim', self.model.input_dim)
        fused = (name in ('summary_corr', 'summary_corrdiff', 'cross_correlation') and
                 ((type(self.model) is MDNN and self.model.net is not None) or
                  type(self.model) is MDRFF) and width >= self.FUSED_CORR_MIN_WIDTH and
                 os.environ.get('BSIG_FUSED_CORR', '1') != '0')
        if not fused:
            return self.summarizer_fxn(states, actions)
        feats = _summarizers.corr_factors(states, actions, use_state_diff=(name == 'summary_corrdiff'))
        print('cross_corr feats', torch.Size(feats.shape), feats.device)   # as the summarizer prints
        return feats

    def predict(self, states, actions, threshold=0.005):
        """Reference bayes_sim.py:116-179 -> host ``pdf.MoG`` posterior."""
        dev = self.model.flat_params.device
        with _lib.nvtx_range('bsig.summarize'):
            xs = self.summarizer_fxn(states.to(dev), actions.to(dev))
        with _lib.nvtx_range('bsig.predict_MoGs'):
            mogs = self.model.predict_MoGs(xs)
        if self.proposal is not None:
            for tmp_i, mog in enumerate(mogs):
                mog.prune_negligible_components(threshold=threshold)
                if isinstance(self.prior, pdf.Uniform):
                    post = mog / self.proposal
                elif isinstance(self.prior, pdf.Gaussian):
                    post = (mog * self.prior) / self.proposal
                else:
                    raise NotImplementedError
                mogs[tmp_i] = post
        if len(mogs) == 1:
            return mogs[0]
        with _lib.nvtx_range('bsig.refit_pooled_mixture'):
            return self._refit_pooled_mixture(mogs)

    # Constants of the multi-trajectory branch (reference bayes_sim.py:150-176, where they
    # are literals): pool size, refit minibatch, passes over the pool, refit network.
    POOL_SAMPLES = int(1e4)
    REFIT_MINIBATCH = 100
    REFIT_PASSES = 5
    REFIT_HIDDEN = (128, 128)

    def _refit_pooled_mixture(self, mogs):
        """Several real trajectories -> one posterior (reference bayes_sim.py:148-179): draw
        POOL_SAMPLES // R samples from each per-trajectory mixture, fit an UNCONDITIONAL
        mixture-density net (constant input) to the pool and return its mixture.
        The pool never leaves the GPU: every mixture samples into a device tensor
        (MoG.gen(device_out=True): numpy draws the uniforms / normals exactly as the
        reference does, the component choice and the affine map run on the device) and the
        refit trains on it in place -- the reference moves the pool host -> device here."""
        model = self.model
        dev = model.flat_params.device
        n_traj = len(mogs)
        per_mog = int(self.POOL_SAMPLES / n_traj)
        refit = MDNN(input_dim=1, output_dim=model.output_dim,
                     output_lows=model.output_lows.detach().cpu().numpy(),
                     output_highs=model.output_highs.detach().cpu().numpy(),
                     n_gaussians=model.n_gaussians, full_covariance=model.L_size > 0,
                     hidden_layers=self.REFIT_HIDDEN, activation=model.activation,
                     lr=model.lr, device=model.device)
        pool = torch.cat([mog.gen(n_samples=per_mog, device_out=True) for mog in mogs], dim=0)
        pool = pool.to(dtype=torch.float32)
        print(f'Fitting posterior from {n_traj:d} mogs')
        n_updates = self.REFIT_PASSES * self.POOL_SAMPLES // self.REFIT_MINIBATCH
        const_in = torch.zeros((pool.shape[0], 1), dtype=torch.float32, device=dev)
        refit.run_training(const_in, pool, n_updates, self.REFIT_MINIBATCH)
        fitted = refit.predict_MoGs(const_in[0:1, :])
        assert (len(fitted) == 1)
        return fitted[0]
