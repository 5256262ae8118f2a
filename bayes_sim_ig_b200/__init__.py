"""B200-native implementation of the BayesSimIG inference hot path.

Python mirror of the reference API (``bayes_sim``, ``models.mdnn``,
``models.mdrff``, ``models.rff``, ``utils.summarizers``, ``utils.pdf``) over
hand-written sm_100a kernels reached through the C ABI in ``include/bsig.h``
(``libbsig_b200.so``).  The alias package ``bayes_sim_ig`` re-exports these
modules under the reference's import paths.
"""
__version__ = '0.1.0'
