"""Stand-in for the third-party ``ghalton`` module (absent on this image).

Only used when importing the *reference* from /root/reference inside this
container (oracle/load_reference.py) -- the reference imports ghalton
unconditionally at models/rff.py:40 and utils/pdf.py:53.  The values produced
here are a plain (identity-permutation) Halton sequence, NOT the Ea-permuted
generalized Halton of the real package: parity tests therefore always copy the
frequency matrix out of the reference object instead of re-deriving it.
"""
import numpy as np


def _first_primes(n):
    primes, c = [], 2
    while len(primes) < n:
        if all(c % p for p in primes if p * p <= c):
            primes.append(c)
        c += 1
    return primes


_PRIMES = _first_primes(4096)
# The real module exposes a table of permutations, one per dimension; callers
# only slice it (EA_PERMS[:d]) and hand it back to GeneralizedHalton.
EA_PERMS = [list(range(p)) for p in _PRIMES]


class GeneralizedHalton(object):
    def __init__(self, perms):
        self._bases = [len(p) for p in perms]
        self._next = 1

    def get(self, n):
        out = np.empty((n, len(self._bases)), dtype=np.float64)
        for r in range(n):
            idx = self._next + r
            for c, base in enumerate(self._bases):
                f, x, i = 1.0, 0.0, idx
                while i > 0:
                    f /= base
                    x += f * (i % base)
                    i //= base
                out[r, c] = x
        self._next += n
        return out.tolist()
