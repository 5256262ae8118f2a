"""Mixture Density NN for BayesSim on hand-written sm_100a kernels.

Mirror of the reference module ``bayes_sim_ig/models/mdnn.py``: the class
``MDNN(nn.Module)`` keeps the reference's constructor, attributes, parameter
names (``net.fcon{i}``, ``pi``, ``mu``, ``Diag.0``, ``Lower`` -> state_dict
compatible), ``forward``, ``mdn_loss_fn``, ``run_training``,
``normalize_samples`` and ``predict_MoGs``.

Two execution forms share the same kernels (csrc/):

* the API form -- ``forward`` / ``mdn_loss_fn`` are differentiable through
  ``torch.autograd.Function`` wrappers whose forward AND backward are our
  kernels (dense layers, head epilogue, mixture NLL); ``loss.backward()`` and
  ``torch.optim`` work exactly as with the reference;
* the training form -- ``run_training`` records the whole call (every Adam
  step: row gather + layers + fused head/NLL forward-backward + dgrad/wgrad +
  Adam over one flat parameter buffer, plus the periodic test-loss
  evaluations) into ONE CUDA graph and replays it; losses and the finiteness
  flag come back in a single device->host copy at the end, instead of the
  reference's ~45 synchronisations per step.

All parameters live in one flat fp32 buffer (``self.flat_params``); the
``nn.Parameter`` objects are views into it, so ``state_dict`` /
``load_state_dict`` / external optimisers keep working.
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from ..utils import pdf

ACT_NONE, ACT_TANH = 0, 1


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


class _LinearFn(torch.autograd.Function):
    """y = act(x @ w^T + b) with our GEMM kernels in both directions."""

    @staticmethod
    def forward(ctx, x, w, b, act, engine):
        x = x.contiguous()
        w = w.contiguous()
        b = b.contiguous()
        m, k = x.shape
        n = w.shape[0]
        y = torch.empty((m, n), dtype=torch.float32, device=x.device)
        ws = _ws(_lib.load().bsig_linear_ws_bytes(m, n, k), x.device)
        with torch.cuda.device(x.device):
            _lib.call('bsig_linear_fwd', _lib.ptr(x), k, None, _lib.ptr(w), _lib.ptr(b),
                      _lib.ptr(y), m, n, k, act, engine, ws.data_ptr(), ws.numel(),
                      _lib.stream_ptr(x.device))
        ctx.save_for_backward(x, w, y)
        ctx.act, ctx.engine = act, engine
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        m, k = x.shape
        n = w.shape[0]
        dev = x.device
        dy = dy.contiguous()
        ws = _ws(_lib.load().bsig_linear_ws_bytes(m, n, k), dev)
        st = _lib.stream_ptr(dev)
        with torch.cuda.device(dev):
            if ctx.act == ACT_TANH:
                dpre = torch.empty_like(dy)
                _lib.call('bsig_tanh_bwd', _lib.ptr(dy), _lib.ptr(y), _lib.ptr(dpre), dy.numel(), st)
            else:
                dpre = dy
            dx = None
            if ctx.needs_input_grad[0]:
                dx = torch.empty_like(x)
                _lib.call('bsig_linear_dgrad', _lib.ptr(dpre), _lib.ptr(w), None, _lib.ptr(dx),
                          m, n, k, ACT_NONE, ctx.engine, ws.data_ptr(), ws.numel(), st)
            dw = torch.empty_like(w)
            db = torch.empty(n, dtype=torch.float32, device=dev)
            _lib.call('bsig_linear_wgrad', _lib.ptr(dpre), _lib.ptr(x), k, None, _lib.ptr(dw),
                      _lib.ptr(db), m, n, k, ctx.engine, ws.data_ptr(), ws.numel(), st)
        return dx, dw, db, None, None


class _HeadEpilogueFn(torch.autograd.Function):
    """(z, noise) -> weights, mu, L_d, L  (reference mdnn.py:109-119)."""

    @staticmethod
    def forward(ctx, z, noise, p, k, full_cov):
        z = z.contiguous()
        noise = noise.contiguous()
        b, nh = z.shape
        dev = z.device
        pk = p * k
        weights = torch.empty((b, k), dtype=torch.float32, device=dev)
        l_d = torch.empty((b, p, k), dtype=torch.float32, device=dev)
        ws = torch.zeros(_lib.load().bsig_mdn_ws_bytes(b), dtype=torch.uint8, device=dev)
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.call('bsig_mdn_head_fwd', _lib.ptr(z), _lib.ptr(noise),
                      _lib.ptr(weights), _lib.ptr(l_d), b, p, k, 1 if full_cov else 0,
                      ws.data_ptr(), ws.numel(), _lib.ptr(flag, torch.int32), _lib.stream_ptr(dev))
        mu = z[:, k:k + pk].reshape(b, p, k).clone()
        low = z[:, k + 2 * pk:].reshape(b, -1, k).clone() if nh > k + 2 * pk else None
        ctx.save_for_backward(z, noise, weights)
        ctx.dims = (p, k, full_cov)
        ctx.mark_non_differentiable(flag)
        if low is None:
            return weights, mu, l_d, flag
        return weights, mu, l_d, low, flag

    @staticmethod
    def backward(ctx, d_w, d_mu, d_ld, *rest):
        z, noise, weights = ctx.saved_tensors
        p, k, full_cov = ctx.dims
        b, nh = z.shape
        dev = z.device
        d_low = rest[0] if len(rest) == 2 else None

        def dense(g, shape):
            return torch.zeros(shape, dtype=torch.float32, device=dev) if g is None \
                else g.contiguous()
        d_w = dense(d_w, (b, k))
        d_mu = dense(d_mu, (b, p, k))
        d_ld = dense(d_ld, (b, p, k))
        has_low = nh > k + 2 * p * k
        if has_low:
            d_low = dense(d_low, (b, (nh - k - 2 * p * k) // k, k))
        dz = torch.empty_like(z)
        ws = torch.zeros(_lib.load().bsig_mdn_ws_bytes(b), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.call('bsig_mdn_head_bwd', _lib.ptr(z), _lib.ptr(noise),
                      _lib.ptr(weights), _lib.ptr(d_w), _lib.ptr(d_mu), _lib.ptr(d_ld),
                      _lib.ptr(d_low) if has_low else None, _lib.ptr(dz), b, p, k,
                      1 if full_cov else 0, ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
        return dz, None, None, None, None


class _MogNllFn(torch.autograd.Function):
    """loss = -mean_b logsumexp_k(...)  (reference mdnn.py:127-178)."""

    @staticmethod
    def forward(ctx, weights, mu, l_d, low, y):
        weights, mu, l_d, y = (t.contiguous() for t in (weights, mu, l_d, y))
        low = None if low is None else low.contiguous()
        b, p, k = mu.shape
        dev = mu.device
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = torch.zeros(_lib.load().bsig_mdn_ws_bytes(b), dtype=torch.uint8, device=dev)
        lsz = 0 if low is None else low.shape[1]
        with torch.cuda.device(dev):
            _lib.call('bsig_mog_nll_fwd', _lib.ptr(weights), _lib.ptr(mu), p * k, _lib.ptr(l_d),
                      p * k, None if low is None else _lib.ptr(low), lsz * k, _lib.ptr(y), None,
                      _lib.ptr(loss), b, p, k, ws.data_ptr(), ws.numel(),
                      _lib.ptr(flag, torch.int32), _lib.stream_ptr(dev))
        ctx.save_for_backward(weights, mu, l_d, y, *(() if low is None else (low,)))
        ctx.has_low = low is not None
        ctx.mark_non_differentiable(flag)
        return loss.reshape(()), flag

    @staticmethod
    def backward(ctx, g_loss, _g_flag):
        saved = ctx.saved_tensors
        weights, mu, l_d, y = saved[:4]
        low = saved[4] if ctx.has_low else None
        b, p, k = mu.shape
        dev = mu.device
        lsz = 0 if low is None else low.shape[1]
        d_w = torch.empty_like(weights)
        d_mu = torch.empty_like(mu)
        d_ld = torch.empty_like(l_d)
        d_low = None if low is None else torch.empty_like(low)
        ws = torch.zeros(_lib.load().bsig_mdn_ws_bytes(b), dtype=torch.uint8, device=dev)
        gs = g_loss.reshape(1).contiguous().float()
        with torch.cuda.device(dev):
            _lib.call('bsig_mog_nll_bwd', _lib.ptr(weights), _lib.ptr(mu), p * k, _lib.ptr(l_d),
                      p * k, None if low is None else _lib.ptr(low), lsz * k, _lib.ptr(y), None,
                      _lib.ptr(gs), _lib.ptr(d_w), _lib.ptr(d_mu), _lib.ptr(d_ld),
                      None if low is None else _lib.ptr(d_low), b, p, k, ws.data_ptr(),
                      ws.numel(), _lib.stream_ptr(dev))
        return d_w, d_mu, d_ld, d_low, None


class MDNN(nn.Module):
    LL_LIMIT = 1.0e5     # limit log likelihood to avoid large gradients
    MIN_WEIGHT = 1.0e-5  # minimum component weights to enable updates
    EPS_NOISE = 1.e-5    # small noise e.g. for numerical stability

    def __init__(self, input_dim, output_dim, output_lows, output_highs,
                 n_gaussians, full_covariance, hidden_layers, activation, lr,
                 device='cuda', **kwargs):
        """Same parameters as the reference (mdnn.py:26-52).  ``activation`` must
        be torch.nn.Tanh (the only activation BayesSim passes, bayes_sim.py:69)."""
        super(MDNN, self).__init__()
        dev = _lib.require_cuda(device)
        if len(hidden_layers) > 0 and activation is not torch.nn.Tanh:
            raise NotImplementedError('the sm_100a layer kernels implement Tanh only')
        self.output_dim = output_dim
        self.output_lows = None
        self.output_highs = None
        if output_lows is not None:
            self.output_lows = torch.from_numpy(np.asarray(output_lows)).float().to(dev)
            self.output_highs = torch.from_numpy(np.asarray(output_highs)).float().to(dev)
        self.n_gaussians = n_gaussians
        self.activation = activation
        self.lr = lr
        self.device = device
        self.input_dim = input_dim
        self.gemm_engine = -1    # BSIG_GEMM_* selector for the dense layers (-1 = auto)
        # Modules are created on the host in the reference's order so that a
        # given torch seed yields the reference's initial weights.
        net = OrderedDict()
        last_layer_size = input_dim
        for l, layer_size in enumerate(hidden_layers):
            net['fcon%d' % l] = nn.Linear(last_layer_size, layer_size)
            net['nl%d' % l] = activation()
            last_layer_size = layer_size
        self.net = nn.Sequential(net) if len(hidden_layers) > 0 else None
        self.pi = nn.Linear(last_layer_size, n_gaussians)
        self.mu = nn.Linear(last_layer_size, output_dim * n_gaussians)
        self.Diag = nn.Sequential(nn.Linear(last_layer_size, output_dim * n_gaussians))
        self.Lower = None
        self.L_size = int(0.5 * output_dim * (output_dim - 1))
        if self.L_size > 0 and full_covariance:
            self.Lower = nn.Linear(last_layer_size, self.L_size * n_gaussians)
        self.full_covariance = self.Lower is not None
        self.head_in = last_layer_size
        self._flatten_parameters(dev)
        self._plans = {}

    # ------------------------------------------------------------------ parameters
    def _trunk_layers(self):
        return [] if self.net is None else [m for m in self.net if isinstance(m, nn.Linear)]

    def _head_layers(self):
        heads = [self.pi, self.mu, self.Diag[0]]
        if self.Lower is not None:
            heads.append(self.Lower)
        return heads

    def _flatten_parameters(self, dev, pad_multiple=4):
        """Move every parameter into one flat fp32 CUDA buffer, laid out as
        [trunk W,b ...][head weights pi|mu|Diag|Lower][head biases ...] so that the
        four heads form ONE [n_head, H] GEMM operand."""
        order = []
        for lin in self._trunk_layers():
            order += [lin.weight, lin.bias]
        order += [h.weight for h in self._head_layers()]
        order += [h.bias for h in self._head_layers()]
        total = sum(p.numel() for p in order)
        # (pad_multiple = 4 * world in data-parallel runs: equal, 16-byte aligned shards of the
        # flat buffer for the reduce-scatter / sharded Adam / all-gather exchange)
        pad_multiple = max(4, int(pad_multiple))
        self._pad_multiple = pad_multiple
        padded = (total + pad_multiple - 1) // pad_multiple * pad_multiple
        flat = torch.zeros(padded, dtype=torch.float32, device=dev)
        off = 0
        self._offsets = []
        for prm in order:
            n = prm.numel()
            view = flat[off:off + n].view(prm.shape)
            view.copy_(prm.data)
            prm.data = view
            self._offsets.append(off)
            off += n
        self._param_order = order
        self.flat_params = flat
        self.n_params = total
        self.n_head = sum(h.weight.shape[0] for h in self._head_layers())
        n_trunk = sum(lin.weight.numel() + lin.bias.numel() for lin in self._trunk_layers())
        self._head_w_off = n_trunk
        self._head_b_off = n_trunk + self.n_head * self.head_in

    def _params_are_flat(self):
        base = self.flat_params.data_ptr()
        return all(p.data_ptr() == base + 4 * off and p.is_contiguous()
                   for p, off in zip(self._param_order, self._offsets))

    def _ensure_flat(self):
        if not self._params_are_flat():     # e.g. after module.to(...) / .data reassignment
            dev = self._param_order[0].device
            self._flatten_parameters(_lib.require_cuda(dev), getattr(self, '_pad_multiple', 4))
            self._plans = {}

    def _head_weight_bias(self):
        """Concatenated head weight [n_head, H] and bias [n_head] (autograd-visible)."""
        heads = self._head_layers()
        return (torch.cat([h.weight for h in heads], dim=0),
                torch.cat([h.bias for h in heads], dim=0))

    # --------------------------------------------------------------------- forward
    def _features(self, x):
        return x          # MDRFF overrides: random Fourier features

    def forward(self, x):
        """Reference mdnn.py:89-125 -> (weights, mu, L_d, L)."""
        if not x.is_cuda:
            raise _lib.BsigError('MDNN.forward needs CUDA tensors; there is no CPU fallback')
        x = self._features(x.float())
        h = x.contiguous()
        for lin in self._trunk_layers():
            h = _LinearFn.apply(h, lin.weight, lin.bias, ACT_TANH, self.gemm_engine)
        w_heads, b_heads = self._head_weight_bias()
        z = _LinearFn.apply(h, w_heads, b_heads, ACT_NONE, self.gemm_engine)
        noise = torch.rand_like(torch.empty(
            (z.shape[0], self.output_dim, self.n_gaussians), device=z.device)).detach()
        outs = _HeadEpilogueFn.apply(z, noise, self.output_dim, self.n_gaussians,
                                     self.full_covariance)
        if self.full_covariance:
            weights, mu, L_d, L, flag = outs
        else:
            (weights, mu, L_d, flag), L = outs, None
        # one synchronisation instead of the reference's 3-4 isfinite asserts
        assert (int(flag.item()) == 0), 'non-finite MDNN output'
        return weights, mu, L_d, L

    def mdn_loss_fn(self, weights, mu, L_d, L, y):
        """Reference mdnn.py:127-178."""
        loss, flag = _MogNllFn.apply(weights, mu, L_d, L, y.float())
        assert (int(flag.item()) == 0), 'non-finite mixture log-likelihood'
        return loss

    def normalize_samples(self, params):
        rng = self.output_highs - self.output_lows
        normed_params = (params - self.output_lows) / rng
        return normed_params

    # -------------------------------------------------------------------- training
    def run_training(self, x_data, y_data, n_updates, batch_size, test_frac=0.2):
        """Reference mdnn.py:180-243: fresh Adam, ordered 80/20 split, n_updates
        minibatch steps with replacement sampling from numpy's global RNG, test
        loss every n_updates//5 steps.  Returns {'train_loss', 'test_loss'}."""
        from .train_engine import run_training_captured
        return run_training_captured(self, x_data, y_data, n_updates, batch_size, test_frac)

    # ------------------------------------------------------------------ prediction
    def predict_MoGs(self, xs):
        """Reference mdnn.py:250-289: the conditional mixture at every row of xs
        as host ``pdf.MoG`` objects (float32 parameters, as in the reference).
        Full covariance with more than one row uses L[pt]."""
        ntest, dim = xs.size()
        with torch.no_grad():
            pi, mu, L_d, L = self(xs)
        pi, mu, L_d = pi.contiguous(), mu.contiguous(), L_d.contiguous()
        L = None if L is None else L.contiguous()
        dev = mu.device
        p, k = self.output_dim, self.n_gaussians
        lsz = 0 if L is None else L.shape[1]
        a_out = torch.empty((ntest, k), dtype=torch.float32, device=dev)
        means = torch.empty((ntest, k, p), dtype=torch.float32, device=dev)
        packed = torch.empty((ntest, k, p + lsz), dtype=torch.float32, device=dev)
        normalize = self.output_lows is not None
        with torch.cuda.device(dev):
            _lib.call('bsig_mog_denorm', _lib.ptr(pi), _lib.ptr(mu), p * k, _lib.ptr(L_d), p * k,
                      None if L is None else _lib.ptr(L), lsz * k,
                      _lib.ptr(self.output_lows) if normalize else None,
                      _lib.ptr(self.output_highs) if normalize else None,
                      _lib.ptr(a_out), _lib.ptr(means), _lib.ptr(packed), ntest, p, k,
                      _lib.stream_ptr(dev))
        a_np, m_np, l_np = a_out.cpu().numpy(), means.cpu().numpy(), packed.cpu().numpy()
        return [pdf.MoG(a=a_np[pt], ms=list(m_np[pt]), Ls=list(l_np[pt]))
                for pt in range(ntest)]
