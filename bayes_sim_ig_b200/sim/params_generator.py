"""The posterior consumer next to the hot path.

The reference's ``ParamsGenerator`` (sim/params_generator.py) is mostly Isaac Gym
property plumbing (closed simulator, out of scope); the part that touches the
posterior is three lines (``:110-118``): ``set_distr`` stores the MoG returned by
``BayesSim.predict`` and ``sample`` draws ONE parameter vector per call, clipped
to the parameter ranges -- once per environment reset, i.e. ``numEnvs`` (up to
10 000) host round trips per reset wave (sim/apply_randomizations.py:154-158).

``ParamsSampler`` keeps those two methods with the reference's semantics and adds
``sample_batch``: all environments' parameter vectors in one device launch.
"""
import numpy as np


class ParamsSampler(object):
    """``lows`` / ``highs``: per-parameter ranges (``ParamsGenerator.lows/highs``)."""

    def __init__(self, lows, highs, distr=None):
        self._lows = np.asarray(lows)
        self._highs = np.asarray(highs)
        assert self._lows.shape == self._highs.shape
        self._distr = distr

    @property
    def lows(self):
        return self._lows

    @property
    def highs(self):
        return self._highs

    def set_distr(self, distr):
        """Reference params_generator.py:110-111."""
        self._distr = distr

    def sample(self):
        """Reference params_generator.py:113-117: one clipped draw."""
        flat_smpl = self._distr.gen(n_samples=1)[0]
        flat_smpl = np.clip(flat_smpl, self._lows, self._highs)
        return flat_smpl

    def sample_batch(self, n_envs, method='random', u=None, z=None):
        """[n_envs, P] clipped draws, row e = what ``sample()`` returns for the e-th
        (uniform, normals) pair; one kernel launch instead of n_envs host calls.
        Distributions without a device sampler fall back to their own ``gen``."""
        if hasattr(self._distr, 'gen_per_env'):
            return self._distr.gen_per_env(n_envs, self._lows, self._highs, method=method, u=u, z=z)
        return np.clip(self._distr.gen(n_samples=n_envs), self._lows, self._highs)
