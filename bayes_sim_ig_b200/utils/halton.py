"""Low-discrepancy points for quasi-random RFF frequencies / Halton sampling.

The reference delegates to the third-party ``ghalton`` package
(models/rff.py:114-116, utils/pdf.py:121-123,302-304), which is un-pinned and
absent from this image.  If ``ghalton`` is importable it is used, so installed
systems keep the reference's exact point set; otherwise a digit-scrambled
Halton sequence is generated here.  The points are constructor-time constants
(never on the per-sample hot path), and parity tests copy the reference's
frequency matrix instead of re-deriving it.
"""
import numpy as np


def _first_primes(n):
    primes, cand = [], 2
    while len(primes) < n:
        if all(cand % q for q in primes if q * q <= cand):
            primes.append(cand)
        cand += 1
    return primes


def _scramble(base):
    """Deterministic digit permutation with perm[0] == 0 (keeps points in (0,1))."""
    rs = np.random.RandomState(base)
    perm = np.arange(base)
    if base > 2:
        perm[1:] = rs.permutation(np.arange(1, base))
    return perm


def halton_points(n_points, dim, skip=1):
    """[n_points, dim] points in (0,1), skipping the first ``skip`` elements."""
    try:
        import ghalton  # noqa: F401  (optional, un-pinned third party)
        seq = ghalton.GeneralizedHalton(ghalton.EA_PERMS[:dim])
        return np.array(seq.get(n_points + skip))[skip:]
    except ImportError:
        pass
    bases = _first_primes(dim)
    out = np.empty((n_points, dim), dtype=np.float64)
    for c, base in enumerate(bases):
        perm = _scramble(base)
        idx = np.arange(skip, skip + n_points)
        x = np.zeros(n_points)
        f = 1.0 / base
        while idx.max() > 0:
            x += f * perm[idx % base]
            idx //= base
            f /= base
        out[:, c] = x
    return out
