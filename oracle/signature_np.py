"""float64 restatement of the truncated path signature.  TEST INFRASTRUCTURE.

PARITY UNPINNED: the reference obtains signatures from the third-party
``signatory`` package (``signatory.signature(paths, depth=depth)``, reference
bayes_sim_ig/utils/summarizers.py:158,164; library defaults: no basepoint, no
stream, no inverse, no scalar term).  signatory is absent from /root/reference
and from this image, and the reference's tests hold no signature values, so
this file follows the *published definition* only:

    Sig(x_0..x_{L-1}) = exp(d_1) (x) exp(d_2) (x) ... (x) exp(d_{L-1}),
    d_t = x_t - x_{t-1},   exp(d) = (d, d(x)d/2!, d(x)d(x)d/3!, ...),

and returns levels 1..depth, each flattened C-order, concatenated level-major
(which is signatory's documented output layout).  Two independent
formulations are provided so that they can be checked against each other:
``signature`` (Chen recursion) and ``signature_iterated_sums`` (closed-form
iterated sums of a piecewise-linear path).
"""
import numpy as np


def _tensor_exp_levels(d, depth):
    """Levels 1..depth of exp(d) for d [N, C] -> list of [N, C^k]."""
    n = d.shape[0]
    levels = [d]
    for k in range(2, depth + 1):
        nxt = (levels[-1][:, :, None] * d[:, None, :]).reshape(n, -1) / k
        levels.append(nxt)
    return levels


def _chen_product(a, b, depth):
    """(1, a_1..a_depth) (x) (1, b_1..b_depth), truncated at depth."""
    n = a[0].shape[0]
    out = []
    for lvl in range(1, depth + 1):
        acc = a[lvl - 1] + b[lvl - 1]
        for i in range(1, lvl):
            j = lvl - i
            acc = acc + (a[i - 1][:, :, None] * b[j - 1][:, None, :]).reshape(n, -1)
        out.append(acc)
    return out


def signature(path, depth):
    """path [N, L, C] (any float dtype) -> float64 [N, sum_{k=1..depth} C^k]."""
    path = np.asarray(path, dtype=np.float64)
    n, length, c = path.shape
    if depth < 1:
        return np.zeros((n, 0), dtype=np.float64)
    if length < 2:
        raise ValueError('a path needs at least two points')
    incr = path[:, 1:, :] - path[:, :-1, :]
    sig = _tensor_exp_levels(incr[:, 0, :], depth)
    for t in range(1, length - 1):
        sig = _chen_product(sig, _tensor_exp_levels(incr[:, t, :], depth), depth)
    return np.concatenate(sig, axis=1)


def signature_iterated_sums(path, depth):
    """Same quantity from the explicit iterated-sum formulas (depth <= 3)."""
    path = np.asarray(path, dtype=np.float64)
    n, length, c = path.shape
    d = path[:, 1:, :] - path[:, :-1, :]            # [N, T, C]
    before = np.cumsum(d, axis=1) - d               # sum_{s<t} d_s
    out = [d.sum(axis=1)]
    if depth >= 2:
        lvl2_before = np.einsum('nti,ntj->ntij', before, d)
        lvl2_self = 0.5 * np.einsum('nti,ntj->ntij', d, d)
        lvl2_steps = lvl2_before + lvl2_self          # contribution of step t
        out.append(lvl2_steps.sum(axis=1).reshape(n, -1))
    if depth >= 3:
        # S2 of the path strictly before step t
        s2_before = np.cumsum(lvl2_steps, axis=1) - lvl2_steps
        t1 = np.einsum('ntij,ntk->ntijk', s2_before, d)
        t2 = 0.5 * np.einsum('nti,ntj,ntk->ntijk', before, d, d)
        t3 = np.einsum('nti,ntj,ntk->ntijk', d, d, d) / 6.0
        out.append((t1 + t2 + t3).sum(axis=1).reshape(n, -1))
    if depth > 3:
        raise ValueError('iterated-sum formulation only written up to depth 3')
    return np.concatenate(out, axis=1)
