"""numpy restatement of the reference trajectory summarizers.  TEST INFRASTRUCTURE.

Follows reference bayes_sim_ig/utils/summarizers.py (line numbers cited per
function).  Pinned against the live reference by tests/golden/make_golden.py
(golden vectors: tests/golden/summarizers_*.npz) except ``summary_signatory``
whose arithmetic is signatory's (parity unpinned, see oracle/__init__.py).
All functions take/return numpy arrays; float32 in -> float32 out, with the
reductions (mean / unbiased std) carried in float64 and rounded once.
"""
import numpy as np

from . import signature_np


def pad_states_actions(states, actions, tgt_actions_len=None):
    """summarizers.py:20-62.  Chop to ``tgt`` steps, or pad by repeating the
    last step.  The reference's padding branch builds ``last.repeat(1,npad,1)``
    from a 2-D tensor, which only concatenates when ntraj == 1;
    for ntraj > 1 it raises, and so does this restatement."""
    assert states.ndim == 3 and actions.ndim == 3
    tgt = states.shape[1] if tgt_actions_len is None else tgt_actions_len

    def fit(x):
        npad = tgt - x.shape[1]
        if npad <= 0:
            return x[:, :tgt, :]
        if x.shape[0] != 1:
            raise RuntimeError('reference padding is only defined for ntraj == 1')
        return np.concatenate([x, np.repeat(x[:, -1:, :], npad, axis=1)], axis=1)

    states, actions = fit(states), fit(actions)
    assert states.shape[1] == actions.shape[1]
    return states, actions


def summary_start(states, actions, max_t=10):
    """summarizers.py:65-70: first max_t steps, [s_t | a_t] per step, flattened."""
    s, a = pad_states_actions(states, actions, max_t)
    return np.concatenate([s, a], axis=-1).reshape(s.shape[0], -1)


def summary_waypts(states, actions, n_waypts=10):
    """summarizers.py:73-87.  Because the chop to n_waypts steps happens before
    the stride is computed, chunk_sz is always 1 and the result equals
    summary_start(max_t=n_waypts); output dtype is float32."""
    s, a = pad_states_actions(states, actions, n_waypts)
    ntraj, traj_len, sdim = s.shape
    assert n_waypts <= traj_len and traj_len == a.shape[1]
    chunk = int(traj_len / n_waypts)
    out = np.zeros((ntraj, n_waypts, sdim + a.shape[-1]), dtype=np.float32)
    t = 0
    for w in range(n_waypts):
        out[:, w, :sdim] = s[:, t, :]
        out[:, w, sdim:] = a[:, t, :]
        t += chunk
    return out.reshape(ntraj, -1)


def cross_correlation(states, actions, use_state_diff=False):
    """summarizers.py:90-122.  W = 10 (5 if state_dim > 50) leading steps;
    state features are differences between ADJACENT STATE DIMENSIONS (Q2) or
    the first D-1 dimensions; features = [outer(sf, af) | mean(sf) | std(sf)]."""
    states, actions = pad_states_actions(states, actions)
    ntraj, traj_len, sdim = states.shape
    assert traj_len > 1 and actions.shape[1] == traj_len
    w = 5 if sdim > 50 else 10
    if traj_len > w:
        sa = summary_waypts(states, actions, n_waypts=w).reshape(ntraj, w, -1)
        states, actions = sa[:, :, :sdim], sa[:, :, sdim:]
    if use_state_diff:
        sf = states[:, :, 1:] - states[:, :, :-1]
    else:
        sf = states[:, :, :-1]
    sf = np.ascontiguousarray(sf).reshape(ntraj, -1)
    af = np.ascontiguousarray(actions).reshape(ntraj, -1)
    cross = (sf[:, :, None] * af[:, None, :]).reshape(ntraj, -1)
    mu = sf.astype(np.float64).mean(axis=-1, keepdims=True)
    if sf.shape[1] < 2:
        std = np.zeros_like(mu)
    else:
        std = sf.astype(np.float64).std(axis=-1, ddof=1, keepdims=True)
    feats = np.concatenate(
        [cross, mu.astype(cross.dtype), std.astype(cross.dtype)], axis=-1)
    assert np.isfinite(feats).all()
    return feats


def summary_corrdiff(states, actions):
    """summarizers.py:125-126."""
    return cross_correlation(states, actions, use_state_diff=True)


def summary_corr(states, actions):
    """summarizers.py:129-130."""
    return cross_correlation(states, actions, use_state_diff=False)


def signature_depth(ndim):
    """summarizers.py:133-141: largest depth in {3,2,1,0} with ndim^depth <= 110^2."""
    for depth in (3, 2, 1, 0):
        if ndim ** depth <= 110 ** 2:
            return depth
    return 1


def signature_paths(states, actions):
    """summarizers.py:153-155: path = [t (1..L) | s_t | a_t]."""
    n, length, _ = states.shape
    t = np.arange(1, length + 1, dtype=np.float64).reshape(1, -1, 1)
    t = np.repeat(t, n, axis=0)
    return np.concatenate([t, states.astype(np.float64),
                           actions.astype(np.float64)], axis=-1)


def summary_signatory(states, actions, keep_tail_rows=True):
    """summarizers.py:144-168.  keep_tail_rows=False reproduces the reference's
    N > 10000 chunking, which silently drops the last N mod 10 rows (Q5)."""
    paths = signature_paths(states, actions)
    depth = signature_depth(paths.shape[-1])
    n = paths.shape[0]
    if n > 10000 and not keep_tail_rows:
        n = (n // 10) * 10
    out = signature_np.signature(paths[:n], depth)
    return out.astype(states.dtype)
