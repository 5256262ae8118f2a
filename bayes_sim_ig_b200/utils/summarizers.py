"""Trajectory summarizers on sm_100a kernels.

Mirror of the reference module ``bayes_sim_ig/utils/summarizers.py`` (same
function names, arguments, return shapes and assertion behaviour):

    f(states [N, T, D], actions [N, T, A]) -> feats [N, F]     (fp32, CUDA)

``summary_start`` / ``summary_waypts`` / ``summary_corr`` / ``summary_corrdiff``
run the streaming kernels in csrc/summarizers.cu; ``summary_signatory`` runs
the Chen-recursion kernel in csrc/signature.cu (no ``signatory`` dependency,
SURVEY Q4).  Inputs must live on a CUDA device: there is no CPU path.

Documented divergences from the reference:
  * ``summary_signatory`` never drops rows (the reference's N > 10000 chunking
    loses the last N mod 10 trajectories, summarizers.py:159-168);
  * non-fp32 inputs are converted to fp32 (the reference already returns fp32
    from ``summary_waypts`` whatever the input dtype, summarizers.py:81).
"""
import torch

from .. import _lib

__all__ = ['pad_states_actions', 'summary_start', 'summary_waypts', 'cross_correlation',
           'summary_corrdiff', 'summary_corr', 'signature_depth', 'summary_signatory',
           'summary_width', 'CorrFactors', 'corr_factors']


def _as_kernel_input(t):
    if not t.is_cuda:
        raise _lib.BsigError('summarizers need CUDA tensors (got %s); no CPU fallback' % t.device)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def pad_states_actions(states, actions, tgt_actions_len=None):
    """Reference summarizers.py:20-62.  Chop (a view) or pad by repeating the
    last step; like the reference, padding only concatenates for ntraj == 1."""
    assert (len(states.shape) == 3), 'Need states: ntraj x n_steps x state_dim'
    assert (len(actions.shape) == 3), 'Need actions: ntraj x n_steps x state_dim'
    if tgt_actions_len is None:
        tgt_actions_len = states.shape[1]

    def fit(x):
        npad = tgt_actions_len - x.shape[1]
        if npad > 0:
            last = x[:, -1, :].clone()
            # same construction as the reference: a [1, N*npad, dim] block,
            # which torch.cat only accepts when N == 1
            return torch.cat([x, last.repeat(1, npad, 1)], dim=1)
        return x[:, :tgt_actions_len, :]

    states, actions = fit(states), fit(actions)
    assert (states.shape[1] == actions.shape[1])
    return states, actions


def _leading_steps(states, actions, n_steps):
    """Tensors whose first n_steps steps are valid + their stored lengths."""
    if states.shape[1] < n_steps or actions.shape[1] < n_steps:
        states, actions = pad_states_actions(states, actions, n_steps)
    states, actions = _as_kernel_input(states), _as_kernel_input(actions)
    return states, actions


def summary_start(states, actions, max_t=10, time_major=False):
    """Reference summarizers.py:65-70: first max_t steps, [s_t | a_t] per step.
    ``time_major=True`` (not in the reference signature): the inputs
    are [n_steps, ntraj, dim] buffers as a vectorised simulator fills them step by step
    (what collect_trajectories.py:55-69 re-assembles per episode); same output."""
    assert (len(states.shape) == 3), 'Need states: ntraj x n_steps x state_dim'
    assert (len(actions.shape) == 3), 'Need actions: ntraj x n_steps x state_dim'
    if time_major:
        assert states.shape[1] == actions.shape[1]
        assert states.shape[0] >= max_t and actions.shape[0] >= max_t, \
            'time-major rollouts must hold at least max_t steps'
        states, actions = _as_kernel_input(states), _as_kernel_input(actions)
        ts, n, d = states.shape
        ta, a = actions.shape[0], actions.shape[2]
        entry = 'bsig_summary_start_tm'
    else:
        assert states.shape[0] == actions.shape[0]
        states, actions = _leading_steps(states, actions, max_t)
        n, ts, d = states.shape
        ta, a = actions.shape[1], actions.shape[2]
        entry = 'bsig_summary_start'
    out = torch.empty((n, max_t * (d + a)), dtype=torch.float32, device=states.device)
    with torch.cuda.device(states.device):
        _lib.call(entry, _lib.ptr(states), _lib.ptr(actions), _lib.ptr(out),
                  n, ts, ta, d, a, max_t, _lib.stream_ptr(states.device))
    return out


def summary_waypts(states, actions, n_waypts=10, time_major=False):
    """Reference summarizers.py:73-87.  The reference chops to n_waypts steps
    before computing its stride, so the stride is always 1 and the result is
    summary_start(max_t=n_waypts)."""
    return summary_start(states, actions, max_t=n_waypts, time_major=time_major)


def cross_correlation(states, actions, use_state_diff=False, time_major=False):
    """Reference summarizers.py:90-122 (``time_major``: see summary_start)."""
    assert (len(states.shape) == 3), 'Need states: ntraj x n_steps x state_dim'
    assert (len(actions.shape) == 3), 'Need actions: ntraj x n_steps x state_dim'
    if time_major:
        traj_len, ntraj, state_dim = states.shape
        assert actions.shape[0] >= min(traj_len, 10) and actions.shape[1] == ntraj
        t_states, t_actions = states.shape[0], actions.shape[0]
        entry = 'bsig_summary_crosscorr_tm'
    else:
        ntraj, traj_len, state_dim = states.shape
        if actions.shape[1] < traj_len:      # reference pads actions up to the states' length
            states, actions = pad_states_actions(states, actions)
        t_states, t_actions = states.shape[1], actions.shape[1]
        entry = 'bsig_summary_crosscorr'
    assert (traj_len > 1)  # empty episodes are problematic
    max_traj_len = 5 if state_dim > 50 else 10
    w = min(traj_len, max_traj_len)
    states, actions = _as_kernel_input(states), _as_kernel_input(actions)
    act_dim = actions.shape[2]
    width = w * (state_dim - 1) * w * act_dim + 2
    feats = torch.empty((ntraj, width), dtype=torch.float32, device=states.device)
    flag = torch.zeros(1, dtype=torch.int32, device=states.device)
    with torch.cuda.device(states.device):
        _lib.call(entry, _lib.ptr(states), _lib.ptr(actions), _lib.ptr(feats),
                  ntraj, t_states, t_actions, state_dim, act_dim, w,
                  1 if use_state_diff else 0, _lib.ptr(flag, torch.int32),
                  _lib.stream_ptr(states.device))
    assert (int(flag.item()) == 0)       # torch.isfinite(feats).all() in the reference
    print('cross_corr feats', feats.shape, feats.device)
    return feats


class CorrFactors(object):
    """Cross-correlation summaries in factored form.

    A row of ``cross_correlation`` is ``[outer(sf, af).ravel() | mean(sf) | std(sf)]``
    (reference summarizers.py:106-119): rank one.  This object holds only the factors,
    ``fac [N, ldf] = [sf (s) | af (q) | mean | std | 0-pad]`` -- 4.6 KB instead of 420 KB per
    ShadowHand trajectory -- and stands in for the ``[N, s*q+2]`` summary tensor wherever the
    consumer is the first dense layer of an MDNN (``MDNN.run_training`` generates the summary
    tiles inside the layer's tcgen05 GEMMs, csrc/corr_layer.cu).  ``materialize()`` gives the
    reference's tensor, bit for bit what ``cross_correlation`` returns."""

    def __init__(self, fac, s, q):
        self.fac, self.s, self.q = fac, int(s), int(q)

    @property
    def shape(self):
        return (self.fac.shape[0], self.s * self.q + 2)

    @property
    def device(self):
        return self.fac.device

    def __len__(self):
        return self.fac.shape[0]

    def __getitem__(self, key):
        """Row selection only (the ordered train / test split of run_training)."""
        if isinstance(key, tuple):
            raise TypeError('CorrFactors supports row indexing only; call materialize()')
        rows = self.fac[key]
        if rows.dim() == 1:
            rows = rows.unsqueeze(0)
        return CorrFactors(rows, self.s, self.q)

    def materialize(self):
        sf = self.fac[:, :self.s]
        af = self.fac[:, self.s:self.s + self.q]
        prod = (sf.unsqueeze(2) * af.unsqueeze(1)).reshape(self.fac.shape[0], -1)
        return torch.cat([prod, self.fac[:, self.s + self.q:self.s + self.q + 2]], dim=1)


def corr_factors(states, actions, use_state_diff=False, time_major=False):
    """``cross_correlation`` without the outer product: same inputs, same padding and
    window rules (reference summarizers.py:90-105), same finiteness assertion; returns
    ``CorrFactors``."""
    assert (len(states.shape) == 3), 'Need states: ntraj x n_steps x state_dim'
    assert (len(actions.shape) == 3), 'Need actions: ntraj x n_steps x state_dim'
    if time_major:
        traj_len, ntraj, state_dim = states.shape
        assert actions.shape[0] >= min(traj_len, 10) and actions.shape[1] == ntraj
        t_states, t_actions = states.shape[0], actions.shape[0]
    else:
        ntraj, traj_len, state_dim = states.shape
        if actions.shape[1] < traj_len:
            states, actions = pad_states_actions(states, actions)
        t_states, t_actions = states.shape[1], actions.shape[1]
    assert (traj_len > 1)
    max_traj_len = 5 if state_dim > 50 else 10
    w = min(traj_len, max_traj_len)
    states, actions = _as_kernel_input(states), _as_kernel_input(actions)
    act_dim = actions.shape[2]
    s, q = w * (state_dim - 1), w * act_dim
    ldf = (s + q + 2 + 3) // 4 * 4
    fac = torch.empty((ntraj, ldf), dtype=torch.float32, device=states.device)
    flag = torch.zeros(1, dtype=torch.int32, device=states.device)
    with torch.cuda.device(states.device):
        _lib.call('bsig_corr_factors', _lib.ptr(states), _lib.ptr(actions), _lib.ptr(fac), ldf,
                  ntraj, t_states, t_actions, state_dim, act_dim, w,
                  1 if use_state_diff else 0, 1 if time_major else 0,
                  _lib.ptr(flag, torch.int32), _lib.stream_ptr(states.device))
    assert (int(flag.item()) == 0)       # torch.isfinite(feats).all() in the reference
    return CorrFactors(fac, s, q)


def summary_corrdiff(states, actions, time_major=False):
    return cross_correlation(states, actions, use_state_diff=True, time_major=time_major)


def summary_corr(states, actions, time_major=False):
    return cross_correlation(states, actions, use_state_diff=False, time_major=time_major)


def signature_depth(ndim):
    """Reference summarizers.py:133-141."""
    max_output_dim = 110**2
    for depth in reversed(range(4)):
        if ndim**depth <= max_output_dim:
            return depth
    return 1


class _SignatureFn(torch.autograd.Function):
    """Truncated signature of [t | states | actions] with our forward and backward
    kernels (the reference's signatory.signature is differentiable too)."""

    @staticmethod
    def forward(ctx, states, actions, depth):
        bsz, path_len, state_dim = states.shape
        act_dim = actions.shape[2]
        c = 1 + state_dim + act_dim
        width = sum(c ** k for k in range(1, depth + 1))
        out = torch.empty((bsz, width), dtype=torch.float32, device=states.device)
        with torch.cuda.device(states.device):
            _lib.call('bsig_signature_fwd', _lib.ptr(states), _lib.ptr(actions), _lib.ptr(out),
                      bsz, path_len, states.shape[1], actions.shape[1], state_dim, act_dim, depth,
                      _lib.stream_ptr(states.device))
        ctx.save_for_backward(states, actions)
        ctx.depth = depth
        ctx.path_len = path_len
        return out

    @staticmethod
    def backward(ctx, grad_out):
        states, actions = ctx.saved_tensors
        bsz, _, state_dim = states.shape
        act_dim = actions.shape[2]
        path_len = ctx.path_len
        grad_out = grad_out.contiguous().float()
        d_states = torch.zeros_like(states)
        d_actions = torch.zeros_like(actions)
        # kernels write dense [n, len, dim] gradients
        ds = d_states if states.shape[1] == path_len else torch.zeros(
            (bsz, path_len, state_dim), dtype=torch.float32, device=states.device)
        da = d_actions if actions.shape[1] == path_len else torch.zeros(
            (bsz, path_len, act_dim), dtype=torch.float32, device=states.device)
        with torch.cuda.device(states.device):
            _lib.call('bsig_signature_bwd', _lib.ptr(states), _lib.ptr(actions),
                      _lib.ptr(grad_out), _lib.ptr(ds), _lib.ptr(da), bsz, path_len,
                      states.shape[1], actions.shape[1], state_dim, act_dim, ctx.depth,
                      _lib.stream_ptr(states.device))
        if ds is not d_states:
            d_states[:, :path_len] = ds
        if da is not d_actions:
            d_actions[:, :path_len] = da
        return d_states, d_actions, None


def summary_signatory(states, actions, time_major=False):
    """Reference summarizers.py:144-168: signature of the time-augmented path
    [t | s_t | a_t] truncated at signature_depth(1 + D + A).
    ``time_major=True``: ``states`` [T, N, D] / ``actions`` [T', N, A] as a
    vectorised simulator writes them.  The signature kernel streams whole trajectories with
    bulk copies, so time-major buffers are transposed on the device first (one extra pass
    over the rollouts; the other summarizers read time-major buffers in place)."""
    if time_major:
        states, actions = states.transpose(0, 1), actions.transpose(0, 1)
    assert (len(states.shape) == 3), 'states should be batch x time x state_dim'
    bsz, path_len, state_dim = states.shape
    assert actions.shape[0] == bsz and actions.shape[1] >= path_len
    states, actions = _as_kernel_input(states), _as_kernel_input(actions)
    act_dim = actions.shape[2]
    c = 1 + state_dim + act_dim
    depth = signature_depth(c)
    if depth == 0:
        return torch.zeros((bsz, 0), dtype=torch.float32, device=states.device)
    return _SignatureFn.apply(states, actions, depth)


def summary_width(name, traj_len, obs_dim, act_dim):
    """Feature width of summarizer ``name`` for rollouts of ``traj_len`` steps
    (what BayesSim.__init__ obtains by probing, bayes_sim.py:57-60)."""
    if name in ('summary_start', 'summary_waypts'):
        return 10 * (obs_dim + act_dim)
    if name in ('summary_corr', 'summary_corrdiff', 'cross_correlation'):
        w = min(traj_len, 5 if obs_dim > 50 else 10)
        return w * (obs_dim - 1) * w * act_dim + 2
    if name == 'summary_signatory':
        c = 1 + obs_dim + act_dim
        return sum(c ** k for k in range(1, signature_depth(c) + 1))
    raise ValueError('unknown summarizer ' + str(name))
