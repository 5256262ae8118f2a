"""numpy restatement of the posterior algebra on the hot path.  TEST INFRASTRUCTURE.

Follows reference bayes_sim_ig/utils/pdf.py (lines cited per function).
Pinned against the live reference by tests/golden/make_golden.py
(tests/golden/pdf_*.npz).  Random draws are explicit inputs: ``u`` is the
``rng.rand(n, 1)`` of pdf.py:74 and ``z`` the concatenation, in component
order, of the ``rng.randn(n_k, P)`` blocks of pdf.py:299.
"""
import numpy as np


def gaussian_from_packed_factor(m, packed):
    """pdf.py:241-251 (Gaussian(m=, L=)): packed = [diag | strict lower in
    np.tril_indices order].  dtype follows the inputs (float32 when the
    parameters come from predict_MoGs).  Returns dict(m, C, S, P, Pm, logdetP)."""
    m = np.asarray(m)
    packed = np.asarray(packed)
    ndim = m.size
    lm = np.diag(packed[:ndim])
    if 1 < ndim < packed.shape[0]:
        r, c = np.tril_indices(ndim, -1)
        lm[r, c] = packed[ndim:]
    cmat = lm.T
    s = np.dot(cmat.T, cmat)
    prec = np.linalg.inv(s)
    return dict(m=m, C=cmat, S=s, P=prec, Pm=np.dot(prec, m),
                logdetP=-2.0 * np.sum(np.log(np.diagonal(cmat))))


def discrete_sample_from_u(p, u):
    """pdf.py:61-76: idx = #{j : u > cumsum(p[:-1])_j}."""
    cumul = np.cumsum(np.asarray(p)[:-1])[np.newaxis, :]
    return np.sum((np.asarray(u).reshape(-1, 1) > cumul).astype(int), axis=1)


def mog_gen_from_draws(a, means, cmats, u, z):
    """pdf.py:465-472 + 296-300: component pick, then samples GROUPED BY
    COMPONENT: block k = z_block_k @ C_k + m_k."""
    idx = discrete_sample_from_u(a, u)
    k = len(means)
    counts = [int(np.sum(idx == i)) for i in range(k)]
    out, start = [], 0
    z = np.asarray(z)
    for i in range(k):
        zi = z[start:start + counts[i]]
        out.append(np.dot(zi, cmats[i]) + means[i])
        start += counts[i]
    return np.concatenate(out, axis=0), idx, counts


def params_samples_per_env(a, means, cmats, u, z, lows=None, highs=None):
    """sim/params_generator.py:115-118 applied once per environment:
    ``np.clip(distr.gen(n_samples=1)[0], lows, highs)`` with environment e consuming
    the uniform u[e] and the normals z[e] (pdf.py:465-472, 296-300 with n_samples=1)."""
    idx = discrete_sample_from_u(a, u)
    z = np.asarray(z)
    out = []
    for e, k in enumerate(idx):
        x = (np.dot(z[e:e + 1], cmats[k]) + means[k])[0]
        if lows is not None:
            x = np.clip(x, lows, highs)
        out.append(x)
    return np.stack(out), idx


def gaussian_logpdf(x, m, prec, logdet_p):
    """pdf.py:328-332 (joint)."""
    xm = x - m
    lp = -np.sum(np.dot(xm, prec) * xm, axis=1)
    lp += logdet_p - m.size * np.log(2.0 * np.pi)
    return 0.5 * lp


def mog_logpdf(x, a, means, precs, logdets, log=True):
    """pdf.py:474-491 (joint, ii=None)."""
    ps = np.array([gaussian_logpdf(x, means[i], precs[i], logdets[i])
                   for i in range(len(a))]).T
    if not log:
        return np.dot(np.exp(ps), a)
    t = ps + np.log(a)
    mx = t.max(axis=1, keepdims=True)
    return (mx + np.log(np.exp(t - mx).sum(axis=1, keepdims=True)))[:, 0]
