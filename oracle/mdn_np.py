"""numpy restatement of the MDNN / MDRFF arithmetic.  TEST INFRASTRUCTURE.

Follows reference bayes_sim_ig/models/mdnn.py, mdrff.py and rff.py (lines cited
per function).  Pinned against the live reference by
tests/golden/make_golden.py (tests/golden/mdn_*.npz): forward outputs, loss,
parameter gradients and post-Adam parameters.

Parameters travel as a dict keyed by the reference's state_dict names
(``net.fcon0.weight`` ..., ``pi.weight``, ``mu.weight``, ``Diag.0.weight``,
``Lower.weight`` + biases).  Everything is computed in the dtype given by
``dtype`` (float64 by default so that the oracle is the more accurate side of
every comparison).  The uniform noise that the reference draws with
``torch.rand_like`` (mdnn.py:116) is an explicit input ``noise_u``.
"""
import numpy as np

LL_LIMIT = 1.0e5      # mdnn.py:22
MIN_WEIGHT = 1.0e-5   # mdnn.py:23
EPS_NOISE = 1.0e-5    # mdnn.py:24
LOG_2PI = float(np.log(2.0 * np.pi))


def hidden_layer_names(params):
    names, i = [], 0
    while 'net.fcon%d.weight' % i in params:
        names.append('net.fcon%d' % i)
        i += 1
    return names


def rff_features(x, freqs, sigma, dtype=np.float64):
    """rff.py:128-132: a * [cos(x (freqs/sigma)^T) | sin(.)], a = sqrt(2/n_feat)."""
    x = np.asarray(x, dtype)
    coeff = np.asarray(freqs, dtype) / np.asarray(sigma, dtype).reshape(1, -1)
    inner = x @ coeff.T
    a = np.sqrt(1.0 / float(coeff.shape[0]))
    return (a * np.concatenate([np.cos(inner), np.sin(inner)], axis=-1)).astype(dtype)


def trunk_forward(params, x, dtype=np.float64):
    """mdnn.py:108: Linear + Tanh per hidden layer; returns all activations."""
    acts = [np.asarray(x, dtype)]
    for name in hidden_layer_names(params):
        w = np.asarray(params[name + '.weight'], dtype)
        b = np.asarray(params[name + '.bias'], dtype)
        acts.append(np.tanh(acts[-1] @ w.T + b))
    return acts


def head_logits(params, h, dtype=np.float64):
    def lin(name):
        return h @ np.asarray(params[name + '.weight'], dtype).T + \
            np.asarray(params[name + '.bias'], dtype)
    z_pi, z_mu, z_d = lin('pi'), lin('mu'), lin('Diag.0')
    z_l = lin('Lower') if 'Lower.weight' in params else None
    return z_pi, z_mu, z_d, z_l


def head_epilogue(z_pi, z_mu, z_d, z_l, noise_u, n_out, n_comp):
    """mdnn.py:109-119.  Returns (weights, mu, L_d, L, cache)."""
    b = z_pi.shape[0]
    zs = z_pi - z_pi.max(axis=1, keepdims=True)
    soft = np.exp(zs)
    soft = soft / soft.sum(axis=1, keepdims=True)
    clamped = np.clip(soft, MIN_WEIGHT, 1.0)
    csum = clamped.sum(axis=1, keepdims=True)
    weights = clamped / csum
    mu = z_mu.reshape(b, n_out, n_comp)
    e = np.exp(z_d).reshape(b, n_out, n_comp)
    eps = EPS_NOISE * e.mean()
    u = np.asarray(noise_u, e.dtype).reshape(b, n_out, n_comp)
    l_d = e + u * eps
    low = None if z_l is None else z_l.reshape(b, -1, n_comp)
    cache = dict(soft=soft, clamped=clamped, csum=csum, e=e, u=u)
    return weights, mu, l_d, low, cache


def mdnn_forward(params, x, noise_u, n_out, n_comp, dtype=np.float64,
                 rff=None):
    """mdnn.py:89-125 (and mdrff.py:28-30 when ``rff=(freqs, sigma)``)."""
    if rff is not None:
        x = rff_features(x, rff[0], rff[1], dtype)
    acts = trunk_forward(params, x, dtype)
    z = head_logits(params, acts[-1], dtype)
    weights, mu, l_d, low, cache = head_epilogue(*z, noise_u, n_out, n_comp)
    cache['acts'] = acts
    for t in (weights, mu, l_d) + (() if low is None else (low,)):
        assert np.isfinite(t).all()          # mdnn.py:120-124
    return weights, mu, l_d, low, cache


def scale_tril(l_d_k, low_k):
    """mdnn.py:152-155: diag_embed + strict lower in np.tril_indices order."""
    b, p = l_d_k.shape
    m = np.zeros((b, p, p), dtype=l_d_k.dtype)
    idx = np.arange(p)
    m[:, idx, idx] = l_d_k
    if low_k is not None:
        r, c = np.tril_indices(p, -1)
        m[:, r, c] = low_k
    return m


def component_log_density(y, mu_k, tril):
    """MultivariateNormal(loc, scale_tril).log_prob(y) as used at mdnn.py:156-158:
    -1/2 |L^-1 (y-mu)|^2 - sum log diag L - P/2 log 2pi.  Returns (g, z)."""
    b, p = y.shape
    d = y - mu_k
    z = np.zeros_like(d)
    for i in range(p):                     # forward substitution
        acc = d[:, i] - np.einsum('bj,bj->b', tril[:, i, :i], z[:, :i])
        z[:, i] = acc / tril[:, i, i]
    idx = np.arange(p)
    g = -0.5 * (z * z).sum(axis=1) - np.log(tril[:, idx, idx]).sum(axis=1) \
        - 0.5 * p * LOG_2PI
    return g, z


def mdn_loss(weights, mu, l_d, low, y, return_parts=False):
    """mdnn.py:127-178: -mean_b logsumexp_k( clamp(g_k) + log clamp(w_k) )."""
    y = np.asarray(y, mu.dtype)
    b, p, k = mu.shape
    res = np.zeros((b, k), dtype=mu.dtype)
    parts = []
    for c in range(k):
        tril = scale_tril(l_d[:, :, c], None if low is None else low[:, :, c])
        g, z = component_log_density(y, mu[:, :, c], tril)
        assert np.isfinite(g).all()        # mdnn.py:172
        gc = np.clip(g, -LL_LIMIT, LL_LIMIT)
        wc = np.clip(weights[:, c], MIN_WEIGHT, 1.0)
        res[:, c] = gc + np.log(wc)
        parts.append((tril, g, z, wc))
    mx = res.max(axis=1, keepdims=True)
    lse = (mx + np.log(np.exp(res - mx).sum(axis=1, keepdims=True)))[:, 0]
    loss = -lse.mean()
    if return_parts:
        return loss, res, lse, parts
    return loss


def mdn_loss_backward(weights, mu, l_d, low, y):
    """Analytic gradient of mdn_loss wrt (weights, mu, l_d, low)."""
    y = np.asarray(y, mu.dtype)
    b, p, k = mu.shape
    loss, res, lse, parts = mdn_loss(weights, mu, l_d, low, y, True)
    rho = np.exp(res - lse[:, None])                 # softmax over components
    d_w = np.zeros_like(weights)
    d_mu = np.zeros_like(mu)
    d_ld = np.zeros_like(l_d)
    d_low = None if low is None else np.zeros_like(low)
    rows, cols = np.tril_indices(p, -1)
    for c in range(k):
        tril, g, z, wc = parts[c]
        coef = -rho[:, c] / b                        # dloss / dres[:, c]
        in_w = (weights[:, c] >= MIN_WEIGHT) & (weights[:, c] <= 1.0)
        d_w[:, c] = coef / wc * in_w
        in_g = (g >= -LL_LIMIT) & (g <= LL_LIMIT)
        cg = coef * in_g
        v = np.zeros_like(z)                         # v = L^-T z
        for i in reversed(range(p)):
            acc = z[:, i] - np.einsum('bj,bj->b', tril[:, i + 1:, i], v[:, i + 1:])
            v[:, i] = acc / tril[:, i, i]
        d_mu[:, :, c] = cg[:, None] * v
        idx = np.arange(p)
        d_ld[:, :, c] = cg[:, None] * (v * z - 1.0 / tril[:, idx, idx])
        if low is not None:
            d_low[:, :, c] = cg[:, None] * (v[:, rows] * z[:, cols])
    return loss, d_w, d_mu, d_ld, d_low


def head_epilogue_backward(cache, weights, d_w, d_mu, d_ld, d_low):
    """Gradient wrt the head logits, including the non-detached eps term
    (mdnn.py:115-116)."""
    soft, clamped, csum = cache['soft'], cache['clamped'], cache['csum']
    d_c = (d_w - (d_w * weights).sum(axis=1, keepdims=True)) / csum
    d_soft = d_c * ((soft >= MIN_WEIGHT) & (soft <= 1.0))
    d_zpi = soft * (d_soft - (d_soft * soft).sum(axis=1, keepdims=True))
    e, u = cache['e'], cache['u']
    d_e = d_ld + EPS_NOISE * (d_ld * u).sum() / e.size
    b = e.shape[0]
    d_zd = (e * d_e).reshape(b, -1)
    d_zmu = d_mu.reshape(b, -1)
    d_zl = None if d_low is None else d_low.reshape(b, -1)
    return d_zpi, d_zmu, d_zd, d_zl


def mdnn_loss_and_grads(params, x, y, noise_u, n_out, n_comp, dtype=np.float64,
                        rff=None):
    """One forward + loss + backward; returns (loss, grads-dict)."""
    weights, mu, l_d, low, cache = mdnn_forward(
        params, x, noise_u, n_out, n_comp, dtype, rff)
    loss, d_w, d_mu, d_ld, d_low = mdn_loss_backward(weights, mu, l_d, low, y)
    d_zpi, d_zmu, d_zd, d_zl = head_epilogue_backward(
        cache, weights, d_w, d_mu, d_ld, d_low)
    acts = cache['acts']
    h = acts[-1]
    grads = {}
    d_h = np.zeros_like(h)
    for name, dz in (('pi', d_zpi), ('mu', d_zmu), ('Diag.0', d_zd), ('Lower', d_zl)):
        if dz is None:
            continue
        grads[name + '.weight'] = dz.T @ h
        grads[name + '.bias'] = dz.sum(axis=0)
        d_h = d_h + dz @ np.asarray(params[name + '.weight'], dtype)
    names = hidden_layer_names(params)
    for li in reversed(range(len(names))):
        d_pre = d_h * (1.0 - acts[li + 1] ** 2)
        grads[names[li] + '.weight'] = d_pre.T @ acts[li]
        grads[names[li] + '.bias'] = d_pre.sum(axis=0)
        d_h = d_pre @ np.asarray(params[names[li] + '.weight'], dtype)
    return loss, grads


def adam_step(params, grads, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam defaults as used at mdnn.py:203,234 (no weight decay,
    no amsgrad): p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)."""
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    for key in params:
        g = grads[key]
        m[key] = beta1 * m[key] + (1.0 - beta1) * g
        v[key] = beta2 * v[key] + (1.0 - beta2) * g * g
        denom = np.sqrt(v[key]) / np.sqrt(bc2) + eps
        params[key] = params[key] - (lr / bc1) * (m[key] / denom)
    return params, m, v


def normalize_samples(y, lows, highs):
    """mdnn.py:245-248."""
    return (y - lows) / (highs - lows)


def predict_mog_params(weights, mu, l_d, low, lows, highs):
    """mdnn.py:250-289 with the evidently intended L[pt,:,k]:
    returns a [R,K], means [R,K,P], packed factors [R,K,P(+P(P-1)/2)]
    = [diag(Rng*Lwr) | (Rng*Lwr)[tril_ids]]."""
    r, p, k = mu.shape
    rng = (highs - lows) if lows is not None else np.ones(p, mu.dtype)
    base = lows if lows is not None else np.zeros(p, mu.dtype)
    means = np.transpose(mu, (0, 2, 1)) * rng + base
    diag = np.transpose(l_d, (0, 2, 1)) * rng
    if low is None:
        return weights, means, diag
    rows, cols = np.tril_indices(p, -1)
    lower = np.transpose(low, (0, 2, 1)) * rng[rows]
    return weights, means, np.concatenate([diag, lower], axis=-1)
