"""torch-CPU port of the reference hot path.  TEST INFRASTRUCTURE / CPU BASELINE.

The reference's own implementation of this path IS torch on the CPU
(bayes_sim_ig/bayes_sim.py, models/mdnn.py, models/mdrff.py, models/rff.py,
utils/summarizers.py).  /root/reference cannot travel to the GPU box, so
``bench.py``'s ``cpu_baseline`` leg and ``--impl reference`` arm time this port
instead: it issues the same torch operations in the same order as the
reference (per-component ``MultivariateNormal`` loop with its finiteness
asserts, autograd backward, ``torch.optim.Adam``, numpy minibatch indices,
host-side ``MoG`` sampling), so its cost on the host cores is representative.

Pinned: tests/test_oracle_golden.py::test_torch_port_* replays the recorded
draws of tests/golden/mdn.npz / bayessim.npz through this port and requires the
reference's losses, gradients and parameters (same ops -> agreement to fp32
rounding).
"""
import numpy as np
import torch
from torch.distributions.multivariate_normal import MultivariateNormal

from . import pdf_np

MIN_WEIGHT, LL_LIMIT, EPS_NOISE = 1.0e-5, 1.0e5, 1.0e-5


# ------------------------------------------------------------------ summarizers
def chop(states, actions, steps):
    """summarizers.py:20-62, chop branch only (T >= steps on the timed path)."""
    return states[:, :steps, :], actions[:, :steps, :]


def summary_start(states, actions, max_t=10):
    """summarizers.py:65-70."""
    s, a = chop(states, actions, max_t)
    return torch.cat([s, a], dim=-1).view(s.shape[0], -1)


def summary_waypts(states, actions, n_waypts=10):
    """summarizers.py:73-87 (python loop of strided copies, stride 1)."""
    s, a = chop(states, actions, n_waypts)
    n, length, sdim = s.shape
    step = int(length / n_waypts)
    out = torch.zeros(n, n_waypts, sdim + a.shape[-1]).to(s.device)
    pos = 0
    for w in range(n_waypts):
        out[:, w, :sdim] = s[:, pos, :]
        out[:, w, sdim:] = a[:, pos, :]
        pos += step
    return out.view(n, -1)


def cross_correlation(states, actions, use_state_diff):
    """summarizers.py:90-122 including the finiteness assert."""
    n, length, sdim = states.shape
    actions = actions[:, :length, :]
    w = 5 if sdim > 50 else 10
    if length > w:
        sa = summary_waypts(states, actions, n_waypts=w).view(n, w, -1)
        states, actions = sa[:, :, :sdim], sa[:, :, sdim:]
    sf = (states[:, :, 1:] - states[:, :, :-1]) if use_state_diff else states[:, :, :-1]
    sf = sf.contiguous().view(n, -1)
    af = actions.contiguous().view(n, -1)
    cross = torch.bmm(sf.unsqueeze(2), af.unsqueeze(1)).view(n, -1)
    mu = sf.mean(dim=-1, keepdim=True)
    std = sf.std(dim=-1, keepdim=True) if sf.shape[1] >= 2 else torch.zeros_like(mu)
    feats = torch.cat([cross, mu, std], dim=-1)
    assert torch.isfinite(feats).all()
    return feats


def summary_corrdiff(states, actions):
    return cross_correlation(states, actions, True)


def summary_corr(states, actions):
    return cross_correlation(states, actions, False)


SUMMARIZERS = {'summary_start': summary_start, 'summary_waypts': summary_waypts,
               'summary_corr': summary_corr, 'summary_corrdiff': summary_corrdiff}


# ------------------------------------------------------------------------ model
class PortModel(torch.nn.Module):
    """MDNN (hidden tanh layers) or MDRFF (rff=(freqs, sigma), no hidden layers)
    with the reference's parameter names."""

    def __init__(self, input_dim, output_dim, lows, highs, n_comp, full_cov, hidden, lr,
                 rff=None):
        super().__init__()
        self.p, self.k, self.lr = output_dim, n_comp, lr
        self.lows = None if lows is None else torch.as_tensor(lows).float()
        self.highs = None if highs is None else torch.as_tensor(highs).float()
        self.rff = rff
        width = input_dim if rff is None else 2 * rff[0].shape[0]
        layers = []
        from collections import OrderedDict
        seq = OrderedDict()
        for i, h in enumerate(hidden):
            seq['fcon%d' % i] = torch.nn.Linear(width, h)
            seq['nl%d' % i] = torch.nn.Tanh()
            width = h
        self.net = torch.nn.Sequential(seq) if len(hidden) else None
        self.pi = torch.nn.Linear(width, n_comp)
        self.mu = torch.nn.Linear(width, output_dim * n_comp)
        self.Diag = torch.nn.Sequential(torch.nn.Linear(width, output_dim * n_comp))
        self.l_size = output_dim * (output_dim - 1) // 2
        self.Lower = torch.nn.Linear(width, self.l_size * n_comp) \
            if (full_cov and self.l_size > 0) else None
        del layers

    def features(self, x):
        if self.rff is None:
            return x
        freqs, sigma = self.rff                       # rff.py:128-132
        inner = torch.matmul(x, (freqs / sigma).T)
        scale = np.sqrt(1.0 / float(freqs.shape[0]))
        return scale * torch.cat([torch.cos(inner), torch.sin(inner)], dim=-1)

    def forward(self, x):
        """mdnn.py:89-125."""
        h = self.features(x)
        h = self.net(h) if self.net is not None else h
        w = torch.nn.functional.softmax(self.pi(h), -1)
        w = torch.clamp(w, MIN_WEIGHT, 1.0)
        w = w / torch.sum(w, dim=1, keepdim=True)
        mu = self.mu(h).reshape(-1, self.p, self.k)
        ld = torch.exp(self.Diag(h)).reshape(-1, self.p, self.k)
        ld = ld + torch.rand_like(ld).detach() * (EPS_NOISE * ld.mean())
        low = self.Lower(h).reshape(-1, self.l_size, self.k) if self.Lower is not None else None
        for t in (w, mu, ld) + (() if low is None else (low,)):
            assert torch.isfinite(t).all()
        return w, mu, ld, low

    def loss(self, w, mu, ld, low, y):
        """mdnn.py:127-178: per-component MultivariateNormal loop + logsumexp."""
        res = torch.zeros(y.size()[0], self.k)
        rows, cols = np.tril_indices(self.p, -1)
        for c in range(self.k):
            tril = torch.diag_embed(ld[:, :, c], offset=0, dim1=-2, dim2=-1)
            if low is not None:
                tril[:, rows, cols] = low[:, :, c]
            g = MultivariateNormal(loc=mu[:, :, c], scale_tril=tril).log_prob(y)
            g = torch.clamp(g, -LL_LIMIT, LL_LIMIT)
            wc = torch.clamp(w[:, c], MIN_WEIGHT, 1.0)
            res[:, c] = g + wc.log()
            assert torch.isfinite(g).all()
            assert torch.isfinite(wc).all()
            assert torch.isfinite(res).all()
        return (-1.0 * torch.logsumexp(res, dim=1)).mean()

    def normalize(self, y):
        return (y - self.lows) / (self.highs - self.lows)

    def run_training(self, x, y, n_updates, batch, test_frac=0.2, idx=None):
        """mdnn.py:180-243; ``idx`` [n_updates, batch] overrides the numpy draws."""
        self.train()
        opt = torch.optim.Adam(self.parameters(), lr=self.lr)
        if self.lows is not None:
            y = self.normalize(y)
        n_train = max(int(x.shape[0] * (1.0 - test_frac)), 1)
        xtr, ytr, xte, yte = x[:n_train], y[:n_train], x[n_train:], y[n_train:]
        train_log, test_log = [], []
        every = max(n_updates // 5, 1)
        for e in range(n_updates):
            ids = np.random.randint(0, n_train, batch) if idx is None else idx[e]
            opt.zero_grad()
            loss = self.loss(*self(xtr[ids]), ytr[ids])
            loss.backward()
            opt.step()
            if e % every == 0 or e + 1 == n_updates:
                test_log.append(self.loss(*self(xte), yte).item())
                train_log.append(loss.item())
        return {'train_loss': train_log, 'test_loss': test_log}

    def predict_mogs(self, xs):
        """mdnn.py:250-289 -> list of (a, means, packed factors) per row."""
        w, mu, ld, low = self(xs)
        rng = self.highs - self.lows
        out = []
        rows, cols = np.tril_indices(self.p, -1)
        for r in range(xs.shape[0]):
            means, packed = [], []
            for c in range(self.k):
                means.append((mu[r, :, c] * rng + self.lows).detach().numpy())
                lwr = torch.diag_embed(ld[r, :, c])
                if low is not None:
                    lwr[rows, cols] = low[r, :, c]
                lwr = torch.matmul(torch.diag(rng), lwr)
                combo = torch.diag(lwr)
                if low is not None:
                    combo = torch.cat([combo, lwr[rows, cols]], dim=-1)
                packed.append(combo.detach().numpy())
            out.append((w[r].detach().numpy(), means, packed))
        return out


def mog_sample_host(a, means, packed, n_samples):
    """pdf.py:465-472 on the host (numpy global RNG), as the reference does."""
    comps = [pdf_np.gaussian_from_packed_factor(m, l) for m, l in zip(means, packed)]
    u = np.random.rand(n_samples, 1)
    idx = pdf_np.discrete_sample_from_u(a, u)
    blocks = []
    for i, comp in enumerate(comps):
        n_i = int(np.sum(idx == i))
        blocks.append(np.dot(np.random.randn(n_i, len(means[0])), comp['C']) + comp['m'])
    return np.concatenate(blocks, axis=0)


def fit_pipeline(states, actions, params, model, summarizer, chunk=1000, n_updates=100,
                 batch=100, test_frac=0.2, n_posterior_samples=10000):
    """The benchmarked unit of work (bayes_sim_main.py:157-167 + predict + sampling):
    for every chunk of <= 1000 trajectories summarize + run_training; then the
    posterior for one held-out trajectory and n_posterior_samples draws from it."""
    fn = SUMMARIZERS[summarizer]
    logs = None
    for lo in range(0, states.shape[0], chunk):
        feats = fn(states[lo:lo + chunk], actions[lo:lo + chunk])
        logs = model.run_training(feats, params[lo:lo + chunk], n_updates, batch, test_frac)
    a, means, packed = model.predict_mogs(fn(states[:1], actions[:1]))[0]
    smp = mog_sample_host(a, means, packed, n_posterior_samples)
    return logs, smp
