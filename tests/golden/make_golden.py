"""Generate the committed golden vectors from the LIVE reference.

Run inside the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Imports the unmodified reference through oracle/load_reference.py (signatory /
ghalton supplied by oracle/refshim), drives it on small seeded inputs, records
every random draw it consumes (torch.rand_like noise, numpy minibatch indices,
numpy uniforms / normals) and writes inputs + outputs to tests/golden/*.npz.
The GPU box has no /root/reference: tests there read these files only.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, '..', '..')))

from oracle import load_reference  # noqa: E402

ref_bs = load_reference.load('bayes_sim')
ref_mdnn = load_reference.load('models.mdnn')
ref_mdrff = load_reference.load('models.mdrff')
ref_sum = load_reference.load('utils.summarizers')
ref_pdf = load_reference.load('utils.pdf')


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


class RecordRandLike(object):
    """Patch torch.rand_like so that every draw is recorded."""
    def __init__(self):
        self.draws = []

    def __enter__(self):
        self._orig = torch.rand_like

        def rec(t, *a, **k):
            out = self._orig(t, *a, **k)
            self.draws.append(out.detach().cpu().numpy().copy())
            return out
        torch.rand_like = rec
        return self

    def __exit__(self, *exc):
        torch.rand_like = self._orig


class RecordRandint(object):
    def __init__(self):
        self.draws = []

    def __enter__(self):
        self._orig = np.random.randint

        def rec(*a, **k):
            out = self._orig(*a, **k)
            self.draws.append(np.array(out).copy())
            return out
        np.random.randint = rec
        return self

    def __exit__(self, *exc):
        np.random.randint = self._orig


def synth(seed, n, t1, d, a):
    g = torch.Generator('cpu').manual_seed(seed)
    states = torch.randn(n, t1, d, generator=g).clamp_(-100, 100)
    actions = torch.rand(n, t1, a, generator=g)
    return states, actions


def golden_summarizers():
    out = {}
    cases = {  # name: (seed, N, T1, D, A)
        'pendulum': (11, 7, 21, 3, 1),
        'cartpole': (12, 5, 21, 4, 1),
        'ant': (13, 3, 51, 60, 8),
        'humanoid': (14, 2, 11, 108, 21),
        'exact10': (15, 4, 10, 3, 2),
        'short6': (16, 3, 6, 4, 2),      # T1 < W: corr uses all 6 steps
        'single_pad': (17, 1, 6, 3, 1),  # N == 1: summary_start pads to 10
    }
    for name, (seed, n, t1, d, a) in cases.items():
        s, ac = synth(seed, n, t1, d, a)
        out[name + '.states'] = s.numpy()
        out[name + '.actions'] = ac.numpy()
        with quiet():
            if t1 >= 10 or n == 1:
                out[name + '.summary_start'] = ref_sum.summary_start(s, ac).numpy()
                out[name + '.summary_waypts'] = ref_sum.summary_waypts(s, ac).numpy()
            out[name + '.summary_corr'] = ref_sum.summary_corr(s, ac).numpy()
            out[name + '.summary_corrdiff'] = ref_sum.summary_corrdiff(s, ac).numpy()
        assert torch.equal(ref_sum.summary_start(s, ac) if (t1 >= 10 or n == 1)
                           else torch.zeros(1),
                           ref_sum.summary_waypts(s, ac) if (t1 >= 10 or n == 1)
                           else torch.zeros(1))
    out['signature_depth.in'] = np.array([1, 5, 6, 22, 23, 69, 110, 111, 130, 232])
    out['signature_depth.out'] = np.array(
        [ref_sum.signature_depth(int(c)) for c in out['signature_depth.in']])
    np.savez_compressed(os.path.join(HERE, 'summarizers.npz'), **out)
    print('summarizers.npz', len(out), 'arrays')


def state_to_np(model):
    return {k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}


def golden_mdn():
    out = {}
    cases = {
        # name: (cls, input_dim, P, K, full_cov, hidden, B)
        'diag': ('MDNN', 12, 3, 4, False, (16, 16), 9),
        'full': ('MDNN', 6, 4, 3, True, (8,), 7),
        'full_big': ('MDNN', 10, 13, 10, True, (32, 32), 12),
        'p1': ('MDNN', 5, 1, 2, True, (8, 8), 6),
        'rff': ('MDRFF', 7, 3, 5, False, (), 10),
        'rff_full': ('MDRFF', 150, 2, 3, True, (), 8),
    }
    for name, (cls, din, p, k, full, hidden, b) in cases.items():
        torch.manual_seed(100 + len(name))
        np.random.seed(200 + len(name))
        lows = np.linspace(0.1, 0.5, p)
        highs = lows + np.linspace(1.0, 3.0, p)
        kwargs = dict(input_dim=din, output_dim=p, output_lows=lows,
                      output_highs=highs, n_gaussians=k, full_covariance=full,
                      hidden_layers=hidden, activation=torch.nn.Tanh, lr=1e-3,
                      device='cpu')
        with quiet():
            if cls == 'MDNN':
                model = ref_mdnn.MDNN(**kwargs)
            else:
                model = ref_mdrff.MDRFF(n_feat=20, sigma=4.0, kernel='RBF', **kwargs)
        x = torch.randn(b, din)
        y_raw = torch.from_numpy(lows + (highs - lows) * np.random.rand(b, p)).float()
        y = model.normalize_samples(y_raw)
        out[name + '.meta'] = np.array([din, p, k, int(full), b] + list(hidden))
        out[name + '.lows'] = lows
        out[name + '.highs'] = highs
        out[name + '.x'] = x.numpy()
        out[name + '.y'] = y.numpy()
        out[name + '.y_raw'] = y_raw.numpy()
        if cls == 'MDRFF':
            out[name + '.rff.freqs'] = model.rff.freqs.numpy()
            out[name + '.rff.sigma'] = model.rff.sigma.numpy()
            out[name + '.rff.features'] = model.rff.to_features(x).numpy()
        for key, val in state_to_np(model).items():
            out[name + '.init.' + key] = val
        # three Adam steps on the same batch with fresh noise each step
        opt = torch.optim.Adam(model.parameters(), lr=model.lr)
        for step in range(3):
            with RecordRandLike() as rec:
                opt.zero_grad()
                w, mu, ld, low = model(x)
                loss = model.mdn_loss_fn(w, mu, ld, low, y)
                loss.backward()
            assert len(rec.draws) == 1
            tag = '%s.step%d.' % (name, step)
            out[tag + 'noise'] = rec.draws[0]
            out[tag + 'weights'] = w.detach().numpy().copy()
            out[tag + 'mu'] = mu.detach().numpy().copy()
            out[tag + 'L_d'] = ld.detach().numpy().copy()
            if low is not None:
                out[tag + 'L'] = low.detach().numpy().copy()
            out[tag + 'loss'] = np.array(loss.item())
            for key, prm in model.named_parameters():
                out[tag + 'grad.' + key] = prm.grad.detach().numpy().copy()
            opt.step()
            for key, val in state_to_np(model).items():
                out[tag + 'after.' + key] = val
        # predict_MoGs at R = 1 (the only R the reference supports with full cov)
        xs = x[:1] if full else x[:3]
        with RecordRandLike() as rec:
            mogs = model.predict_MoGs(xs)
        out[name + '.predict.noise'] = rec.draws[0]
        out[name + '.predict.xs'] = xs.numpy()
        out[name + '.predict.a'] = np.stack([m.a for m in mogs])
        for fld in ('m', 'C', 'S', 'P', 'Pm'):
            out[name + '.predict.' + fld] = np.stack(
                [np.stack([getattr(g, fld) for g in m.xs]) for m in mogs])
        out[name + '.predict.logdetP'] = np.stack(
            [np.array([g.logdetP for g in m.xs]) for m in mogs])
    np.savez_compressed(os.path.join(HERE, 'mdn.npz'), **out)
    print('mdn.npz', len(out), 'arrays')


def golden_mdn_bench():
    """A full training update at BASELINE configs[1]'s dimensions (minibatch 100, corrdiff
    width 302, hidden 128 x 128, P = 13, K = 10; diagonal and full covariance): the reference's
    own loop body (mdnn.py:221-234: randint rows -> forward -> loss -> backward -> Adam) for
    four updates, with the minibatch rows and the eps-noise recorded.  Written to its own file
    (mdn_bench.npz) because the parameter vectors are large."""
    out = {}
    for name, full, n_steps in (('bench_diag', False, 4), ('bench_full', True, 3)):
        din, p, k, hidden, b, n_rows = 302, 13, 10, (128, 128), 100, 128
        torch.manual_seed(7 + int(full))
        np.random.seed(17 + int(full))
        lows = np.full(p, 0.1)
        highs = np.full(p, 2.0)
        with quiet():
            model = ref_mdnn.MDNN(input_dim=din, output_dim=p, output_lows=lows,
                                  output_highs=highs, n_gaussians=k, full_covariance=full,
                                  hidden_layers=hidden, activation=torch.nn.Tanh, lr=1e-3,
                                  device='cpu')
        x = torch.randn(n_rows, din)
        y_raw = torch.from_numpy(lows + (highs - lows) * np.random.rand(n_rows, p)).float()
        y = model.normalize_samples(y_raw)
        out[name + '.meta'] = np.array([din, p, k, int(full), b] + list(hidden))
        out[name + '.lows'], out[name + '.highs'] = lows, highs
        out[name + '.x'], out[name + '.y_raw'] = x.numpy(), y_raw.numpy()
        for key, val in state_to_np(model).items():
            out[name + '.init.' + key] = val
        opt = torch.optim.Adam(model.parameters(), lr=model.lr)
        idx_all, noise_all, losses = [], [], []
        for step in range(n_steps):
            with RecordRandLike() as rec, RecordRandint() as ri:
                ids = np.random.randint(0, n_rows, b)
                opt.zero_grad()
                w, mu, ld, low = model(x[ids])
                loss = model.mdn_loss_fn(w, mu, ld, low, y[ids])
                loss.backward()
                opt.step()
            assert len(rec.draws) == 1 and len(ri.draws) == 1
            idx_all.append(ri.draws[0])
            noise_all.append(rec.draws[0])
            losses.append(loss.item())
        out[name + '.idx'] = np.stack(idx_all).astype(np.int64)
        out[name + '.noise'] = np.stack(noise_all)
        out[name + '.loss'] = np.array(losses)
        for key, val in state_to_np(model).items():
            out[name + '.after.' + key] = val
    np.savez_compressed(os.path.join(HERE, 'mdn_bench.npz'), **out)
    print('mdn_bench.npz', len(out), 'arrays')


def golden_pdf():
    out = {}
    rs = np.random.RandomState(7)
    for name, p, k, dt in (('f32', 3, 4, np.float32), ('f64', 5, 6, np.float64),
                           ('p1', 1, 3, np.float32), ('p13', 13, 10, np.float32)):
        a = rs.rand(k) + 0.05
        a = (a / a.sum()).astype(dt)
        ms = [rs.randn(p).astype(dt) for _ in range(k)]
        nl = p * (p - 1) // 2
        ls = [np.concatenate([0.3 + rs.rand(p), 0.2 * rs.randn(nl)]).astype(dt)
              for _ in range(k)]
        mog = ref_pdf.MoG(a=a, ms=ms, Ls=ls)
        out[name + '.a'] = a
        out[name + '.ms'] = np.stack(ms)
        out[name + '.Ls'] = np.stack(ls)
        for fld in ('C', 'S', 'P', 'Pm'):
            out[name + '.' + fld] = np.stack([getattr(g, fld) for g in mog.xs])
        out[name + '.logdetP'] = np.array([g.logdetP for g in mog.xs])
        n = 257
        np.random.seed(31)
        smp = mog.gen(n_samples=n)
        np.random.seed(31)
        u = np.random.rand(n, 1)
        z = np.random.randn(n, p)     # == concatenated per-component randn blocks
        out[name + '.gen.u'] = u
        out[name + '.gen.z'] = z
        out[name + '.gen.samples'] = smp
        np.random.seed(32)
        out[name + '.discrete.idx'] = ref_pdf.discrete_sample(a, 100)
        np.random.seed(32)
        out[name + '.discrete.u'] = np.random.rand(100, 1)
        xq = (rs.randn(33, p) * 1.5)
        out[name + '.eval.x64'] = xq
        out[name + '.eval.log64'] = mog.eval(xq, log=True)
        out[name + '.eval.lin64'] = mog.eval(xq, log=False)
        out[name + '.eval.log32'] = mog.eval(xq.astype(np.float32), log=True)
        one = mog.gen(n_samples=1)
        assert one.shape == (1, p)
        # the posterior consumer: ParamsGenerator.sample (sim/params_generator.py:115-118),
        # once per environment; its draws are replayed to record (u, z) per environment
        n_env = 41
        lows, highs = np.full(p, -0.8), np.full(p, 0.9)
        np.random.seed(33)
        envs = np.stack([np.clip(mog.gen(n_samples=1)[0], lows, highs) for _ in range(n_env)])
        np.random.seed(33)
        eu, ez = [], []
        for _ in range(n_env):
            eu.append(np.random.rand(1, 1))
            ez.append(np.random.randn(1, p))
        out[name + '.envs.lows'], out[name + '.envs.highs'] = lows, highs
        out[name + '.envs.u'] = np.concatenate(eu)
        out[name + '.envs.z'] = np.concatenate(ez)
        out[name + '.envs.samples'] = envs
    # pruning (pdf.py:562-570)
    a = np.array([0.001, 0.5, 0.002, 0.497])
    mog = ref_pdf.MoG(a=a, ms=[np.zeros(2)] * 4, Ls=[np.ones(2)] * 4)
    mog.prune_negligible_components(threshold=0.005)
    out['prune.a_in'] = a
    out['prune.a_out'] = mog.a
    np.savez_compressed(os.path.join(HERE, 'pdf.npz'), **out)
    print('pdf.npz', len(out), 'arrays')


def golden_pdf_marginal():
    """Pairwise 2-D marginal densities on regular grids as utils/plot.py:38-44 evaluates them
    (get_2d_posterior_data: posterior.eval(X.T, ii=dims, log=False) on an np.mgrid), pair by
    pair through the LIVE reference (MoG.eval -> Gaussian.eval marginal branch,
    pdf.py:334-339, which jitters the covariance with numpy's global RNG)."""
    out = {}
    rs = np.random.RandomState(11)
    p, k, nbins = 5, 6, 32
    a = rs.rand(k) + 0.05
    a = a / a.sum()
    ms = [0.3 + rs.rand(p) for _ in range(k)]
    nl = p * (p - 1) // 2
    ls = [np.concatenate([0.2 + 0.3 * rs.rand(p), 0.1 * rs.randn(nl)]) for _ in range(k)]
    mog = ref_pdf.MoG(a=a, ms=ms, Ls=ls)
    out['a'], out['ms'], out['Ls'] = a, np.stack(ms), np.stack(ls)
    pairs = [(i, j) for i in range(p) for j in range(i + 1, p)]
    lims = np.array([[-0.2 + 0.05 * q, 1.6, 0.0, 1.8 - 0.03 * q] for q in range(len(pairs))])
    out['pairs'], out['lims'], out['nbins'] = np.array(pairs), lims, np.array(nbins)
    np.random.seed(91)
    grids, logs = [], []
    for (i, j), (xmin, xmax, ymin, ymax) in zip(pairs, lims):
        xi, yi = np.mgrid[xmin:xmax:nbins * 1j, ymin:ymax:nbins * 1j]
        X = np.concatenate((xi.reshape(1, nbins * nbins), yi.reshape(1, nbins * nbins)), axis=0)
        grids.append(mog.eval(X.T, ii=[i, j], log=False).reshape(nbins, nbins))
    out['density'] = np.stack(grids)
    np.random.seed(92)
    for (i, j), (xmin, xmax, ymin, ymax) in zip(pairs, lims):
        xi, yi = np.mgrid[xmin:xmax:nbins * 1j, ymin:ymax:nbins * 1j]
        X = np.concatenate((xi.reshape(1, nbins * nbins), yi.reshape(1, nbins * nbins)), axis=0)
        logs.append(mog.eval(X.T, ii=[i, j], log=True).reshape(nbins, nbins))
    out['logdensity'] = np.stack(logs)
    np.savez_compressed(os.path.join(HERE, 'pdf_marginal.npz'), **out)
    print('pdf_marginal.npz', len(out), 'arrays')


def golden_pdf_host():
    """Host-side pdf surface that is not on the device path (SURVEY 8.b list): Uniform,
    Gaussian algebra / KL, MoG moments / projection / sampled KL, fit_mog.  Every value
    comes from the live reference (py2-named operators are called explicitly)."""
    out = {}
    rs = np.random.RandomState(11)
    # ---- Uniform (pdf.py:79-192)
    lb, ub = np.array([0.0, 1.0, -2.0]), np.array([1.0, 3.0, 0.5])
    uni = ref_pdf.Uniform(lb, ub)
    np.random.seed(41)
    out['uni.lb'], out['uni.ub'] = lb, ub
    out['uni.gen'] = uni.gen(n_samples=7)                      # Q8: scrambled dimensions
    xq = np.stack([rs.uniform(-0.5, 1.5, 9), rs.uniform(0.5, 3.5, 9), rs.uniform(-2.5, 1.0, 9)], 1)
    xq[0] = [0.5, 2.0, -1.0]                                     # at least one point inside
    out['uni.x'] = xq
    out['uni.eval_lin'] = uni.eval(xq, log=False)
    out['uni.eval_log'] = uni.eval(xq, log=True)
    out['uni.eval_marg'] = uni.eval(xq[:, [0, 2]], ii=[0, 2], log=False)
    # ---- Gaussian algebra (pdf.py:344-411)
    def rand_gauss(p):
        a = rs.randn(p, p)
        return rs.randn(p), np.dot(a, a.T) + p * np.eye(p)
    m1, s1 = rand_gauss(3)
    m2, _ = rand_gauss(3)
    s2 = 3.0 * s1 + np.eye(3)                                    # broader: g1 / g2 stays proper
    g1, g2 = ref_pdf.Gaussian(m=m1, S=s1), ref_pdf.Gaussian(m=m2, S=s2)
    out['g.m1'], out['g.S1'], out['g.m2'], out['g.S2'] = m1, s1, m2, s2
    for name, g in (('mul', g1 * g2), ('div', g1.__div__(g2)), ('pow', g1 ** 2.5)):
        out['g.%s.m' % name], out['g.%s.S' % name] = g.m, g.S
        out['g.%s.logdetP' % name] = np.array(g.logdetP)
    out['g.kl'] = np.array(g1.kl(g2))
    xg = rs.randn(6, 3)
    out['g.x'] = xg
    out['g.eval'] = g1.eval(xg)
    # ---- MoG moments / projection / sampled KL / algebra with a Gaussian (pdf.py:501-582)
    a = np.array([0.2, 0.5, 0.3])
    ms = [rs.randn(3) for _ in range(3)]
    ss = [s1 * (0.5 + 0.3 * i) for i in range(3)]                # all narrower than g2
    mog = ref_pdf.MoG(a=a, ms=ms, Ss=ss)
    out['mog.a'], out['mog.ms'], out['mog.Ss'] = a, np.stack(ms), np.stack(ss)
    # (MoG.calc_mean_and_cov / project_to_gaussian raise AttributeError in the reference --
    #  pdf.py:553 reads a non-existent attribute `sigma` -- so they cannot be recorded)
    try:
        mog.calc_mean_and_cov()
        raise SystemExit('reference calc_mean_and_cov unexpectedly works: record it')
    except AttributeError:
        pass
    prod = mog * g2
    out['mog.mul.a'] = prod.a
    out['mog.mul.ms'] = np.stack([x.m for x in prod.xs])
    out['mog.mul.Ss'] = np.stack([x.S for x in prod.xs])
    # MoG.__div__ divides its components with `/`, which Python 3 does not map to the
    # py2-named Gaussian.__div__: alias it on the live class for this call only
    ref_pdf.Gaussian.__truediv__ = ref_pdf.Gaussian.__div__
    try:
        quot = mog.__div__(g2)
    finally:
        del ref_pdf.Gaussian.__truediv__
    out['mog.div.a'] = quot.a
    out['mog.div.ms'] = np.stack([x.m for x in quot.xs])
    out['mog.div.Ss'] = np.stack([x.S for x in quot.xs])
    # ---- fit_mog (pdf.py:584-642): EM from a seeded start on seeded data
    data = np.concatenate([rs.randn(150, 2) * 0.4 + [2.0, 0.0], rs.randn(100, 2) * 0.7 + [-1.5, 1.0]])
    out['fit.x'] = data
    np.random.seed(43)
    fit = ref_pdf.fit_mog(data, n_components=2, tol=1e-7, maxiter=200)
    out['fit.a'] = fit.a
    out['fit.ms'] = np.stack([x.m for x in fit.xs])
    out['fit.Ss'] = np.stack([x.S for x in fit.xs])
    np.savez_compressed(os.path.join(HERE, 'pdf_host.npz'), **out)
    print('pdf_host.npz', len(out), 'arrays')


def golden_rff_host():
    """Row a7 (host construction of the RFF frequencies, rff.py:53-120,135-184): every
    kernel class's sample_freqs (numpy stream) and inv_cdf, and draw_freqs without the
    (absent, un-pinnable) ghalton sequence."""
    ref_rff = load_reference.load('models.rff')
    out = {}
    u = np.linspace(0.03, 0.97, 21).reshape(7, 3)
    out['u'] = u
    for name in ('RFFKernelRBF', 'RFFKernelMatern12', 'RFFKernelMatern32', 'RFFKernelMatern52'):
        kern = getattr(ref_rff, name)()
        np.random.seed(51)
        out[name + '.sample'] = kern.sample_freqs((5, 4))
        out[name + '.inv_cdf'] = kern.inv_cdf(u)
        np.random.seed(52)
        out[name + '.draw'] = ref_rff.RFF.draw_freqs(kern, 6, 3, False)
    np.savez_compressed(os.path.join(HERE, 'rff_host.npz'), **out)
    print('rff_host.npz', len(out), 'arrays')


def load_pendulum(fnm, limit=None):
    loaded = np.load(fnm)
    params = loaded['params']
    data = loaded['data']
    if params.ndim == 1:
        params, data = params.reshape(1, -1), data.reshape(1, -1)
    if limit:
        params, data = params[:limit], data[:limit]
    return params, data


def golden_bayessim():
    """End-to-end BayesSim.run_training + predict on a slice of the reference's
    own Pendulum fixture (tests/data/*.npz), all random draws recorded."""
    out = {}
    data_dir = os.path.join(load_reference.REFERENCE_ROOT, 'bayes_sim_ig', 'tests', 'data')
    params, data = load_pendulum(
        os.path.join(data_dir, 'pendulum_train_data_ones_policy_nornd.npz'), 250)
    tparams, tdata = load_pendulum(
        os.path.join(data_dir, 'pendulum_true_data_ones_policy_nornd.npz'))
    out['pendulum.params'] = params.astype(np.float32)
    out['pendulum.data'] = data.astype(np.float32)
    out['pendulum.true_params'] = tparams.astype(np.float32)
    out['pendulum.true_data'] = tdata.astype(np.float32)
    lows, highs = np.array([0.01] * 2), np.array([2.0] * 2)
    cls = ref_bs.BayesSim
    saved = (cls.NUM_GRAD_UPDATES, cls.MINIBATCH_SIZE)
    cls.NUM_GRAD_UPDATES, cls.MINIBATCH_SIZE = 20, 32
    try:
        for name, model_class, summ in (('mdnn_start', 'MDNN', 'summary_start'),
                                        ('mdrff_corrdiff', 'MDRFF', 'summary_corrdiff')):
            torch.manual_seed(2)
            np.random.seed(2)
            cfg = {'modelClass': model_class, 'summarizerFxn': summ,
                   'trainTrajLen': 10, 'components': 10,
                   'hiddenLayers': (24, 24), 'lr': 5e-4}
            with quiet():
                bsim = cls(model_cfg=cfg, obs_dim=3, act_dim=1, params_dim=2,
                           params_lows=lows, params_highs=highs, prior=None,
                           proposal=None, device='cpu')
            sa = torch.from_numpy(data).float().reshape(params.shape[0], -1, 4)
            st, ac = sa[:, :, :3], sa[:, :, 3:]
            prm = torch.from_numpy(params).float()
            for key, val in state_to_np(bsim.model).items():
                out[name + '.init.' + key] = val
            if model_class == 'MDRFF':
                out[name + '.rff.freqs'] = bsim.model.rff.freqs.numpy()
                out[name + '.rff.sigma'] = bsim.model.rff.sigma.numpy()
            with quiet(), RecordRandLike() as rl, RecordRandint() as ri:
                logs = bsim.run_training(prm, st, ac)
            out[name + '.train.idx'] = np.stack(ri.draws)
            # 20 training forwards + 6 test forwards (epochs 0,4,8,12,16,19)
            out[name + '.train.n_noise'] = np.array(len(rl.draws))
            for i, dr in enumerate(rl.draws):
                out[name + '.train.noise%02d' % i] = dr
            out[name + '.train.train_loss'] = np.array(logs['train_loss'])
            out[name + '.train.test_loss'] = np.array(logs['test_loss'])
            for key, val in state_to_np(bsim.model).items():
                out[name + '.after.' + key] = val
            tsa = torch.from_numpy(tdata).float().reshape(1, -1, 4)
            with quiet(), RecordRandLike() as rl:
                post = bsim.predict(tsa[:, :, :3], tsa[:, :, 3:])
            out[name + '.predict.noise'] = rl.draws[0]
            out[name + '.predict.a'] = post.a
            out[name + '.predict.m'] = np.stack([g.m for g in post.xs])
            out[name + '.predict.S'] = np.stack([g.S for g in post.xs])
            out[name + '.predict.nll_true'] = -post.eval(tparams.astype(np.float32))
    finally:
        cls.NUM_GRAD_UPDATES, cls.MINIBATCH_SIZE = saved
    np.savez_compressed(os.path.join(HERE, 'bayessim.npz'), **out)
    print('bayessim.npz', len(out), 'arrays')


def seeded_uniforms(call_index, shape):
    """The i-th torch.rand_like draw of a replayed run (shared with the GPU test)."""
    g = torch.Generator('cpu').manual_seed(424242 + int(call_index))
    return torch.rand(tuple(shape), generator=g)


def golden_predict_multi():
    """BayesSim.predict with R = 2 real trajectories (reference bayes_sim.py:148-179): the
    per-trajectory mixtures are resampled (numpy RNG) and ONE unconditional MDNN is refitted
    to the 10 000 pooled samples with 500 Adam updates.  Too many draws to store, so they are
    made reproducible instead: numpy is seeded, torch.rand_like is replaced by
    seeded_uniforms(call index), and the refit network's initial weights are recorded."""
    out = {}
    gb = np.load(os.path.join(HERE, 'bayessim.npz'))
    lows, highs = np.array([0.01] * 2), np.array([2.0] * 2)
    cfg = {'modelClass': 'MDNN', 'summarizerFxn': 'summary_start', 'trainTrajLen': 10,
           'components': 10, 'hiddenLayers': (24, 24), 'lr': 5e-4}
    with quiet():
        bsim = ref_bs.BayesSim(model_cfg=cfg, obs_dim=3, act_dim=1, params_dim=2,
                               params_lows=lows, params_highs=highs, prior=None,
                               proposal=None, device='cpu')
    trained = {k[len('mdnn_start.after.'):]: torch.from_numpy(gb[k])
               for k in gb.files if k.startswith('mdnn_start.after.')}
    bsim.model.load_state_dict(trained)
    data = gb['pendulum.data']
    sa = torch.from_numpy(data[:2]).float().reshape(2, -1, 4)
    out['states'], out['actions'] = sa[:, :, :3].numpy(), sa[:, :, 3:].numpy()
    recorded = {}
    orig_mdnn = ref_bs.MDNN

    class RecMDNN(orig_mdnn):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            recorded['init'] = state_to_np(self)

        def run_training(self, x_data, y_data, n_updates, batch_size, test_frac=0.2):
            recorded['pool'] = y_data.detach().numpy().copy()
            recorded['n_updates'] = n_updates
            logs = super().run_training(x_data, y_data, n_updates, batch_size, test_frac)
            recorded['logs'] = logs
            return logs
    calls = {'n': 0}
    orig_rand_like = torch.rand_like

    def fake_rand_like(t, *a, **k):
        u = seeded_uniforms(calls['n'], t.shape).to(t.dtype)
        calls['n'] += 1
        return u
    ref_bs.MDNN = RecMDNN
    torch.rand_like = fake_rand_like
    try:
        np.random.seed(77)
        torch.manual_seed(78)
        with quiet():
            post = bsim.predict(sa[:, :, :3], sa[:, :, 3:])
    finally:
        ref_bs.MDNN = orig_mdnn
        torch.rand_like = orig_rand_like
    out['n_rand_like_calls'] = np.array(calls['n'])
    out['n_updates'] = np.array(recorded['n_updates'])
    for key, val in recorded['init'].items():
        out['refit.init.' + key] = val
    pool = recorded['pool']
    out['pool.shape'] = np.array(pool.shape)
    out['pool.head'] = pool[:16]
    out['pool.tail'] = pool[-16:]
    out['pool.mean'] = pool.astype(np.float64).mean(axis=0)
    out['pool.cov'] = np.cov(pool.astype(np.float64).T)
    out['refit.train_loss'] = np.array(recorded['logs']['train_loss'])
    out['refit.test_loss'] = np.array(recorded['logs']['test_loss'])
    out['post.a'] = post.a
    out['post.m'] = np.stack([g.m for g in post.xs])
    out['post.S'] = np.stack([g.S for g in post.xs])
    xs_eval = np.random.RandomState(5).rand(64, 2) * 2.0
    out['post.eval_x'] = xs_eval
    out['post.eval_logp'] = post.eval(xs_eval.astype(np.float32))
    np.savez_compressed(os.path.join(HERE, 'predict_multi.npz'), **out)
    print('predict_multi.npz', len(out), 'arrays;', calls['n'], 'rand_like calls')


if __name__ == '__main__':
    torch.set_num_threads(1)   # reproducible reduction order
    golden_summarizers()
    golden_mdn()
    if 'bench' in sys.argv[1:] or len(sys.argv) == 1:
        golden_mdn_bench()
    golden_pdf()
    golden_pdf_marginal()
    golden_pdf_host()
    golden_rff_host()
    golden_bayessim()
    golden_predict_multi()
