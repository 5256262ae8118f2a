#!/usr/bin/env python
"""Benchmark of the BayesSimIG hot path on B200 (contract: see DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one pass of the hot path over one batch of synthetic rollouts of
BASELINE.json configs[1]: 4096 Cartpole-shaped trajectories per GPU ->
summary_corrdiff -> MDNN fit (reference constants: chunks of <= 1000
trajectories, 100 Adam updates of minibatch 100 and 6 held-out evaluations per
chunk) -> posterior for one held-out trajectory -> 10 000 posterior samples.
metric = fit trajectories/sec (whole job, all GPUs).  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# ------------------------------------------------------------------ workload (configs[1])
TASK = dict(name='cartpole', D=4, A=1, T1=21, P=13, K=10)
N_TRAJ = 4096
CHUNK = 1000
HIDDEN = (128, 128)
LR = 1e-4
SUMMARIZER = 'summary_corrdiff'
N_POSTERIOR_SAMPLES = 10000
METRIC = 'mdnn_fit_trajectories_per_sec'
UNIT = 'trajectories/s'


def synth(seed, n, task, pin=False):
    """Seeded synthetic rollouts shared by both implementations."""
    g = torch.Generator('cpu').manual_seed(seed)
    states = torch.randn(n, task['T1'], task['D'], generator=g).clamp_(-100, 100)
    actions = torch.rand(n, task['T1'], task['A'], generator=g)
    lows = np.full(task['P'], 0.1)
    highs = np.full(task['P'], 2.0)
    params = torch.from_numpy(lows).float() + torch.rand(n, task['P'], generator=g) * \
        torch.from_numpy(highs - lows).float()
    if pin and torch.cuda.is_available():
        states, actions, params = states.pin_memory(), actions.pin_memory(), params.pin_memory()
    return states, actions, params, lows, highs


def workload_config(n_gpus):
    return {'workload': 'configs[1]: Cartpole-shaped %d trajectories/GPU [T1=21,D=4,A=1,P=13], '
                        'summary_corrdiff(F=302) + MDNN[128,128] K=10 fit (chunks of 1000: 100 Adam '
                        'updates x minibatch 100 + 6 test evals) + predict + 10000 posterior samples'
                        % N_TRAJ,
            'trajectories_per_gpu': N_TRAJ, 'chunk': CHUNK, 'n_updates_per_chunk': 100,
            'minibatch': 100, 'posterior_samples': N_POSTERIOR_SAMPLES,
            'posterior_sampler': 'device RNG (MoG.gen(method="philox")); the reference arm draws with numpy',
            'parallelism': 'dp%d' % n_gpus,
            'gradient_exchange': ('none' if n_gpus == 1 else
                                  os.environ.get('BSIG_DP_EXCHANGE', 'p2p') +
                                  (' (one-shot all-reduce over NVLink peer memory fused into the '
                                   'Adam kernel)' if os.environ.get('BSIG_DP_EXCHANGE', 'p2p') == 'p2p'
                                   else ' (NCCL all-reduce between two CUDA graphs per update)')),
            'l2_policy': 'L2 flushed (256 MiB write) between timed steps'}


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.gpu_index), '--query-gpu=' + self.QUERY,
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        try:
            for line in open(self.path):
                f = [c.strip() for c in line.split(',')]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(names, f[5:9]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out['reasons'] = sorted(reasons)
        return out


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


# --------------------------------------------------------------------------- reference arm
def reference_kind():
    """'reference' when the UNMODIFIED reference modules can be imported (from
    /root/reference in the build container, else from the verbatim staged copy under
    oracle/_ref that travels to the GPU box), 'port' (oracle/torch_port.py) otherwise."""
    try:
        from oracle import load_reference
        return 'reference' if load_reference.reference_available() else 'port'
    except Exception:
        return 'port'


_REF_BSIM = {}


def _reference_bsim(lows, highs):
    """The reference's own BayesSim on the CPU (bayes_sim.py:27-82), built once."""
    from oracle import load_reference
    key = 'cpu'
    if key not in _REF_BSIM:
        ref_bs = load_reference.load('bayes_sim')
        cfg = {'modelClass': 'MDNN', 'summarizerFxn': SUMMARIZER, 'trainTrajLen': TASK['T1'] - 1,
               'components': TASK['K'], 'hiddenLayers': list(HIDDEN), 'lr': LR}
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):     # the width probe prints
            _REF_BSIM[key] = ref_bs.BayesSim(cfg, TASK['D'], TASK['A'], TASK['P'], lows, highs,
                                             prior=None, proposal=None, device='cpu')
    return _REF_BSIM[key]


def cpu_pipeline_rate(n_traj, threads, seed=0):
    """Time the reference CPU path on n_traj trajectories of the workload: per chunk of
    <= 1000 trajectories BayesSim.run_training (summary_corrdiff + 100 Adam updates + 6
    test evals, bayes_sim.py:91-114), then BayesSim.predict for one trajectory and
    MoG.gen(10000).  Runs the unmodified reference when it is importable (kind
    'reference'), the torch port otherwise (kind 'port')."""
    import contextlib
    import io
    torch.set_num_threads(threads)
    states, actions, params, lows, highs = synth(seed, n_traj, TASK)
    torch.manual_seed(seed)
    np.random.seed(seed)
    if reference_kind() == 'reference':
        bsim = _reference_bsim(lows, highs)
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            for lo in range(0, n_traj, CHUNK):
                bsim.run_training(params[lo:lo + CHUNK], states[lo:lo + CHUNK],
                                  actions[lo:lo + CHUNK])
            post = bsim.predict(states[:1], actions[:1])
            post.gen(N_POSTERIOR_SAMPLES)
        dt = time.perf_counter() - t0
        return n_traj / dt, dt
    from oracle import torch_port
    width = 10 * (TASK['D'] - 1) * 10 * TASK['A'] + 2
    model = torch_port.PortModel(width, TASK['P'], lows, highs, TASK['K'], False, HIDDEN, LR)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        torch_port.fit_pipeline(states, actions, params, model, SUMMARIZER, chunk=CHUNK,
                                n_posterior_samples=N_POSTERIOR_SAMPLES)
    dt = time.perf_counter() - t0
    return n_traj / dt, dt


def cpu_sample_note(kind, sample):
    src = ('the UNMODIFIED reference modules (bayes_sim.py, models/mdnn.py, utils/summarizers.py, '
           'utils/pdf.py; verbatim copy staged by oracle/stage_reference.py) with device="cpu"'
           if kind == 'reference' else
           'oracle/torch_port.py (same torch ops as the reference CPU path)')
    return ('%s; %d Cartpole trajectories per step = 1 chunk of the workload: corrdiff + 100 Adam '
            'updates + 6 test evals + predict + 10000 samples' % (src, sample))


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    kind = reference_kind()
    # bounded sample: one 1000-trajectory chunk per step (about 1.5-6 s of CPU work)
    sample = 1000
    for _ in range(args.warmup):
        cpu_pipeline_rate(200, threads)
    rates, times = [], []
    for s in range(args.steps):
        r, dt = cpu_pipeline_rate(sample, threads, seed=s)
        rates.append(r)
        times.append(dt)
    value = sample * len(times) / sum(times)
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * float(np.mean(times)), 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args.gpus),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': kind,
                             'sample': cpu_sample_note(kind, sample)},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0,
                    'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------- B200 arm
def flush_l2(buf):
    buf.add_(1.0)


def time_kernel(fn, flush, reps=20, warm=3):
    """Average device time of fn() in ms, CUDA events on the current stream, L2
    flushed before every timed launch."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def kernel_rooflines(dev, flush, peak_gbs, peak_src):
    """Stand-alone CUDA-event timings of the streaming kernels at sizes that
    exceed L2, as achieved algorithmic GB/s against the measured HBM peak."""
    from bayes_sim_ig_b200 import _lib
    from bayes_sim_ig.utils import summarizers as S
    import contextlib
    import io
    out = {}
    st = lambda: _lib.stream_ptr(dev)

    # DRAM traffic per call (dram__bytes_read.sum + dram__bytes_write.sum, summed over the
    # launches of a call) of the SHIPPED kernels at exactly these sizes: one ncu metrics pass,
    # round 2, profiles/r2/roofline_traffic.summary.txt.  Writes that stay in the 126 MB L2
    # past the end of the launch are not counted by ncu, hence traffic < algorithmic bytes
    # for the write-dominated kernels.
    traffic = {'summary_corrdiff_shadowhand': 1680.89e6, 'summary_corrdiff_cartpole_1M': 1578.41e6,
               'summary_start_humanoid': 640.26e6, 'signature_depth3_cartpole_1M': 1464.50e6,
               'mdn_nll_fused_fwd_bwd': 167.93e6 + 668.56e6 + 436.51e6,
               'rff_projection_ant_64k_simt_fp32': 209.72e6,
               'rff_projection_ant_64k_tcgen05_tf32': 194.06e6,
               'rff_projection_ant_64k_tcgen05_tf32x3': 197.49e6, 'adam_13.5M': 319.93e6}

    def entry(name, algo_bytes, ms, note):
        ach = algo_bytes / (ms * 1e-3) / 1e9
        out[name] = {'bound': 'hbm', 'achieved': round(ach, 1), 'peak': peak_gbs, 'unit': 'GB/s',
                     'frac': round(ach / peak_gbs, 4), 'ms': round(ms, 4),
                     'traffic': int(traffic[name]) if name in traffic else None,
                     'algorithmic_bytes': int(algo_bytes), 'shape': note, 'peak_source': peak_src}

    # summary_corrdiff, ShadowHand-shaped (F = 105002): 4*[W(D+A) + F] B/traj
    n, t1, d, a = 4096, 51, 211, 20          # 1.7 GB of summaries: far beyond the 126 MB L2
    s = torch.randn(n, t1, d, device=dev)
    ac = torch.rand(n, t1, a, device=dev)
    width = 5 * (d - 1) * 5 * a + 2
    feats = torch.empty((n, width), device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)

    def cross():
        _lib.call('bsig_summary_crosscorr', s.data_ptr(), ac.data_ptr(), feats.data_ptr(), n, t1, t1,
                  d, a, 5, 1, flag.data_ptr(), st())
    entry('summary_corrdiff_shadowhand', n * 4 * (5 * (d + a) + width), time_kernel(cross, flush),
          'N=%d T1=51 D=211 A=20 -> F=%d' % (n, width))
    del feats
    # summary_corrdiff, Cartpole-shaped at 1M trajectories (F = 302)
    n, t1, d, a = 1 << 20, 21, 4, 1
    s = torch.randn(n, t1, d, device=dev)
    ac = torch.rand(n, t1, a, device=dev)
    feats = torch.empty((n, 302), device=dev)

    def cross2():
        _lib.call('bsig_summary_crosscorr', s.data_ptr(), ac.data_ptr(), feats.data_ptr(), n, t1, t1,
                  d, a, 10, 1, flag.data_ptr(), st())
    entry('summary_corrdiff_cartpole_1M', n * 4 * (10 * (d + a) + 302), time_kernel(cross2, flush),
          'N=%d T1=21 D=4 A=1 -> F=302' % n)
    # summary_start, Humanoid-shaped
    n2, t2, d2, a2 = 1 << 16, 11, 108, 21
    s2 = torch.randn(n2, t2, d2, device=dev)
    ac2 = torch.rand(n2, t2, a2, device=dev)
    o2 = torch.empty((n2, 10 * (d2 + a2)), device=dev)

    def start():
        _lib.call('bsig_summary_start', s2.data_ptr(), ac2.data_ptr(), o2.data_ptr(), n2, t2, t2, d2,
                  a2, 10, st())
    entry('summary_start_humanoid', n2 * 2 * 4 * 10 * (d2 + a2), time_kernel(start, flush),
          'N=%d T1=11 D=108 A=21 -> F=1290' % n2)
    # path signature depth 3, Cartpole-shaped (C = 6 -> 258)
    sig = torch.empty((n, 258), device=dev)

    def sigk():
        _lib.call('bsig_signature_fwd', s.data_ptr(), ac.data_ptr(), sig.data_ptr(), n, t1, t1, t1,
                  d, a, 3, st())
    entry('signature_depth3_cartpole_1M', n * 4 * (t1 * (d + a) + 258), time_kernel(sigk, flush),
          'N=%d L=21 C=6 depth 3 -> 258' % n)
    del s, ac, feats, sig
    # fused head + mixture NLL forward/backward, B = 262144, P = 13, K = 10 (diag)
    b, p, k = 1 << 18, 13, 10
    nh = k * (1 + 2 * p)
    z = torch.randn(b, nh, device=dev) * 0.3
    dz = torch.empty_like(z)
    noise = torch.rand(b, p, k, device=dev)
    y = torch.rand(b, p, device=dev)
    loss = torch.zeros(1, device=dev)
    ws = torch.zeros(_lib.load().bsig_mdn_ws_bytes(b), dtype=torch.uint8, device=dev)

    def nll():
        _lib.call('bsig_mdn_nll_fused', z.data_ptr(), noise.data_ptr(), y.data_ptr(), None,
                  loss.data_ptr(), dz.data_ptr(), b, p, k, 0, ws.data_ptr(), ws.numel(),
                  flag.data_ptr(), st())
    # algorithmic: read z + noise + y, write dz.  3 launches (exp-sum for the batch-global
    # eps, the streaming NLL kernel, the batch-global eps-gradient fix-up) whose exact
    # semantics need 1.78x the algorithmic bytes: z_d read three times, dz_d written twice
    entry('mdn_nll_fused_fwd_bwd', 4 * b * (2 * nh + p * k + p), time_kernel(nll, flush),
          'B=%d P=13 K=10 diag, fwd+bwd (3 launches; minimum traffic of the exact '
          'batch-global eps semantics = 1.78x algorithmic)' % b)
    # RFF projection + sincos, Ant-shaped summary_start (configs[2]): N=65536, d=680 -> 200
    del z, dz, noise, y
    n3, d3, nf = 1 << 16, 680, 100
    x3 = torch.randn(n3, d3, device=dev)
    coeff = torch.randn(nf, d3, device=dev) / 4.0
    feat = torch.empty(n3, 2 * nf, device=dev)
    wsl = torch.empty(max(_lib.load().bsig_linear_ws_bytes(n3, nf, d3), 16), dtype=torch.uint8,
                      device=dev)
    for eng, tag in ((0, 'simt_fp32'), (1, 'tcgen05_tf32'), (2, 'tcgen05_tf32x3')):
        def rff(eng=eng):
            _lib.call('bsig_rff_features', x3.data_ptr(), d3, None, coeff.data_ptr(), feat.data_ptr(),
                      n3, d3, nf, 0.1, eng, wsl.data_ptr(), wsl.numel(), st())
        ms = time_kernel(rff, flush)
        entry('rff_projection_ant_64k_' + tag, 4 * (n3 * (d3 + 2 * nf) + nf * d3), ms,
              'N=65536 d=680 -> 200 features, engine %s' % tag)
        out['rff_projection_ant_64k_' + tag]['tflops'] = round(2.0 * n3 * d3 * nf / (ms * 1e-3) / 1e12, 2)
    del x3, feat
    # Adam over 13.5 M parameters (ShadowHand MDNN): 28 B/param
    cnt = 13540748
    pr, g, m, v = (torch.zeros(cnt, device=dev) for _ in range(4))

    def adam():
        _lib.call('bsig_adam_step', pr.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), cnt, 1,
                  1e-4, 0.9, 0.999, 1e-8, 1.0, st())
    entry('adam_13.5M', 28 * cnt, time_kernel(adam, flush), 'flat fp32 buffer of %d params' % cnt)
    del pr, g, m, v
    # fused cross-correlation -> first layer (csrc/corr_layer.cu) at the
    # ShadowHand shape: minibatch 100, s = 1050, q = 100 (F = 105002), 128 hidden units.  The
    # summary rows are generated inside the GEMMs; what has to move is the weight (forward:
    # 4 B/param read) and weight + both Adam moments (weight gradient + optimiser: 24 B/param).
    cs, cq, mb, nh0 = 1050, 100, 100, 128
    fdim = cs * cq + 2
    ldf = (cs + cq + 2 + 3) // 4 * 4
    fac = torch.randn(800, ldf, device=dev)
    rows = torch.randint(0, 800, (mb,), device=dev)
    w0 = torch.randn(nh0, fdim, device=dev) / 300.0
    b0 = torch.zeros(nh0, device=dev)
    y0 = torch.empty(mb, nh0, device=dev)
    dy0 = torch.randn(mb, nh0, device=dev)
    ea, es = torch.zeros_like(w0), torch.zeros_like(w0)
    wsc = torch.empty(_lib.load().bsig_corr_linear_ws_bytes(mb, nh0, cs, cq) + 256, dtype=torch.uint8,
                      device=dev)
    step_no = [0]

    def warm_small():          # what the training step keeps L2-resident between kernels
        fac.sum(); dy0.sum(); rows.sum()

    def corr_fwd():
        _lib.call('bsig_corr_linear_fwd', fac.data_ptr(), ldf, rows.data_ptr(), cs, cq, w0.data_ptr(),
                  b0.data_ptr(), y0.data_ptr(), mb, nh0, 1, wsc.data_ptr(), wsc.numel(), st())

    def corr_wgrad():
        step_no[0] += 1
        _lib.call('bsig_corr_linear_wgrad', dy0.data_ptr(), fac.data_ptr(), ldf, rows.data_ptr(), cs,
                  cq, mb, nh0, None, w0.data_ptr(), ea.data_ptr(), es.data_ptr(), step_no[0], 1e-4,
                  0.9, 0.999, 1e-8, 1.0, st())

    def flush_keep_small():
        flush()
        warm_small()
    shape = 'minibatch 100 of ShadowHand factors (s=1050, q=100, F=%d), 128 outputs' % fdim
    traffic['corr_fused_first_layer_fwd'] = 56.2e6
    traffic['corr_fused_first_layer_wgrad_adam'] = 278.7e6
    entry('corr_fused_first_layer_fwd', 4 * (nh0 * fdim + mb * (cs + cq + 2) + mb * nh0),
          time_kernel(corr_fwd, flush_keep_small), shape + ': generated x tiles (TMEM) x W by cp.async, '
          'tcgen05 TF32x3 split-K + reduce (2 launches)')
    entry('corr_fused_first_layer_wgrad_adam', 24 * nh0 * fdim + 4 * mb * (nh0 + cs + cq + 2),
          time_kernel(corr_wgrad, flush_keep_small), shape + ': dW on tcgen05 from generated x^T tiles, '
          'Adam of W / exp_avg / exp_avg_sq in the epilogue (1 launch, the gradient never reaches HBM)')
    for key in ('corr_fused_first_layer_fwd', 'corr_fused_first_layer_wgrad_adam'):
        out[key]['tflops'] = round(2.0 * mb * nh0 * fdim / (out[key]['ms'] * 1e-3) / 1e12, 2)
    return out


def dominant_kernel_roofline(dev, peak_gbs, peak_src):
    """`gemm_small_kernel` is the dominant kernel of the benchmarked step (8 of the 10
    launches of every Adam update, ~70 % of the device time in the ncu launch list,
    profiles/).  Its eight per-update launches are timed here with CUDA events at
    the step's own shapes; operands are L2-resident exactly as in the real step
    (a 360 KB model), so no L2 flush: this is a latency-bound kernel and the
    HBM fraction says so."""
    from bayes_sim_ig_b200 import _lib
    st = lambda: _lib.stream_ptr(dev)
    b, f, h, nh = 100, 302, 128, 270
    x = torch.randn(800, f, device=dev)
    rows = torch.randint(0, 800, (b,), device=dev)
    w1, b1 = torch.randn(h, f, device=dev), torch.randn(h, device=dev)
    w2, b2 = torch.randn(h, h, device=dev), torch.randn(h, device=dev)
    wh, bh = torch.randn(nh, h, device=dev), torch.randn(nh, device=dev)
    h1, h2 = torch.empty(b, h, device=dev), torch.empty(b, h, device=dev)
    z, dz = torch.empty(b, nh, device=dev), torch.randn(b, nh, device=dev)
    dh2, dh1 = torch.empty(b, h, device=dev), torch.empty(b, h, device=dev)
    dwh, dbh = torch.empty(nh, h, device=dev), torch.empty(nh, device=dev)
    dw2, db2 = torch.empty(h, h, device=dev), torch.empty(h, device=dev)
    dw1, db1 = torch.empty(h, f, device=dev), torch.empty(h, device=dev)
    ws = torch.empty(1 << 20, dtype=torch.uint8, device=dev)
    wp, wn = ws.data_ptr(), ws.numel()

    def eight():
        c = _lib.call
        c('bsig_linear_fwd', x.data_ptr(), f, rows.data_ptr(), w1.data_ptr(), b1.data_ptr(),
          h1.data_ptr(), b, h, f, 1, 0, wp, wn, st())
        c('bsig_linear_fwd', h1.data_ptr(), h, None, w2.data_ptr(), b2.data_ptr(), h2.data_ptr(),
          b, h, h, 1, 0, wp, wn, st())
        c('bsig_linear_fwd', h2.data_ptr(), h, None, wh.data_ptr(), bh.data_ptr(), z.data_ptr(),
          b, nh, h, 0, 0, wp, wn, st())
        c('bsig_linear_wgrad', dz.data_ptr(), h2.data_ptr(), h, None, dwh.data_ptr(), dbh.data_ptr(),
          b, nh, h, 0, wp, wn, st())
        c('bsig_linear_dgrad', dz.data_ptr(), wh.data_ptr(), h2.data_ptr(), dh2.data_ptr(), b, nh, h,
          1, 0, wp, wn, st())
        c('bsig_linear_wgrad', dh2.data_ptr(), h1.data_ptr(), h, None, dw2.data_ptr(), db2.data_ptr(),
          b, h, h, 0, wp, wn, st())
        c('bsig_linear_dgrad', dh2.data_ptr(), w2.data_ptr(), h1.data_ptr(), dh1.data_ptr(), b, h, h,
          1, 0, wp, wn, st())
        c('bsig_linear_wgrad', dh1.data_ptr(), x.data_ptr(), f, rows.data_ptr(), dw1.data_ptr(),
          db1.data_ptr(), b, h, f, 0, wp, wn, st())
    # algorithmic bytes of the eight GEMMs: 4*(MK + NK + MN) each
    shapes = [(b, h, f), (b, h, h), (b, nh, h), (nh, h, b), (b, h, nh), (h, h, b), (b, h, h),
              (h, f, b)]
    algo = sum(4 * (m * k + n * k + m * n) for m, n, k in shapes)
    graph = torch.cuda.CUDAGraph()
    eight()
    torch.cuda.synchronize()
    with torch.cuda.graph(graph):
        for _ in range(25):
            eight()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        graph.replay()
    e1.record()
    e1.synchronize()
    us_per_launch = e0.elapsed_time(e1) * 1e3 / (4 * 25 * 8)
    ach = (algo / 8) / (us_per_launch * 1e-6) / 1e9
    return {'kernel': 'gemm_small_kernel', 'bound': 'hbm', 'achieved': round(ach, 2),
            'peak': peak_gbs, 'unit': 'GB/s', 'frac': round(ach / peak_gbs, 5),
            # dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full capture
            # profiles/r1_ncu_step_kernels.summary.txt (operands live in L2: ~0 B read, the
            # rest is write-back)
            'traffic': 15000,
            'us_per_launch': round(us_per_launch, 2), 'algorithmic_bytes': int(algo / 8),
            'peak_source': peak_src,
            'note': 'dominant kernel of the step (8 of 9 launches per Adam update); minibatch-100 '
                    'layer GEMMs with L2-resident operands are latency-bound, not bandwidth-bound: '
                    'see `rooflines` for the HBM-bound kernels at sizes that exceed L2'}


def extra_configs(dev):
    """Other BASELINE.json configurations, measured once each (not the headline):
    configs[2] -- Ant-shaped 65 536 trajectories, summary_start (F=680) + MDRFF fit with the
    reference's constants (66 run_training calls = 6 600 Adam updates; the RFF projection of
    the whole-batch predict runs on the tcgen05 engine)."""
    import contextlib
    import io
    from bayes_sim_ig.bayes_sim import BayesSim
    task = dict(name='ant', D=60, A=8, T1=51, P=17, K=10)
    n = 1 << 16
    states, actions, params, lows, highs = synth(7, n, task)
    states, actions, params = states.to(dev), actions.to(dev), params.to(dev)
    cfg = {'modelClass': 'MDRFF', 'summarizerFxn': 'summary_start', 'trainTrajLen': 50,
           'components': 10, 'hiddenLayers': [128, 128], 'lr': 1e-4}
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        bsim = BayesSim(cfg, task['D'], task['A'], task['P'], lows, highs, prior=None,
                        proposal=None, device=str(dev))

        def fit():
            for lo in range(0, n, CHUNK):
                logs = bsim.run_training(params[lo:lo + CHUNK], states[lo:lo + CHUNK],
                                         actions[lo:lo + CHUNK])
            return logs
        fit()                                   # capture + warm up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        logs = fit()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        feats = bsim.summarizer_fxn(states, actions)
        mogs_in = bsim.model.rff.to_features(feats)      # 65536 x 680 -> 200 on tcgen05 (warm-up)
        torch.cuda.synchronize()
        reps = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            mogs_in = bsim.model.rff.to_features(feats)
            e1.record()
            e1.synchronize()
            reps.append(e0.elapsed_time(e1))
        t_rff = 1e-3 * float(np.median(reps))
    out['mdrff_ant_64k'] = {
        'config': 'configs[2]: Ant-shaped 65536 trajectories [T1=51,D=60,A=8,P=17], '
                  'summary_start(F=680) + MDRFF(n_feat=200, sigma=4, RBF) fit, reference constants',
        'fit_trajectories_per_s': n / dt, 'seconds': dt, 'final_test_loss': logs['test_loss'][-1],
        'rff_features_whole_batch_ms': 1e3 * t_rff, 'rff_timing': 'median of 5 warm calls, CUDA events',
        'features_shape': list(mogs_in.shape)}
    del bsim, states, actions, params, feats, mogs_in
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    for fn in (extra_shadowhand, extra_signature_mdrff, extra_scaled_mode):
        try:
            out.update(fn(dev))
        except Exception as exc:          # an extra must never take the headline line down
            out[fn.__name__] = {'error': repr(exc)[:300]}
            torch.cuda.empty_cache()
    return out


def extra_shadowhand(dev):
    """configs[3], single-GPU slice: ShadowHand-shaped rollouts [T1=51, D=211, A=20, P=32],
    summary_corrdiff (F = 105 002, 420 KB per trajectory) + MDNN[128,128] K=10 diag
    (13.5 M parameters), one reference-sized call: 1000 trajectories, 100 Adam updates of
    minibatch 100.  At this shape an update is HBM-bound: 28 B/param of Adam traffic plus
    the 54 MB first-layer weight read (forward) and gradient write."""
    import contextlib
    import io
    from bayes_sim_ig.bayes_sim import BayesSim
    task = dict(name='shadowhand', D=211, A=20, T1=51, P=32, K=10)
    n = 1000
    states, actions, params, lows, highs = synth(11, n, task)
    states, actions, params = states.to(dev), actions.to(dev), params.to(dev)
    cfg = {'modelClass': 'MDNN', 'summarizerFxn': 'summary_corrdiff', 'trainTrajLen': 50,
           'components': 10, 'hiddenLayers': [128, 128], 'lr': 1e-4}
    with contextlib.redirect_stdout(io.StringIO()):
        bsim = BayesSim(cfg, task['D'], task['A'], task['P'], lows, highs, prior=None,
                        proposal=None, device=str(dev))
        bsim.run_training(params, states, actions)          # capture + warm up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        logs = bsim.run_training(params, states, actions)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    n_par = int(bsim.model.flat_params.numel())
    fused = any(getattr(pl, 'corr', None) is not None for pl in bsim.model._plans.values())
    # the same call with the summary materialised ([1000, 105002] fp32 = 420 MB) and the generic
    # tcgen05 layer GEMMs + flat Adam: what the fused first layer replaces
    os.environ['BSIG_FUSED_CORR'] = '0'
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            bsim2 = BayesSim(cfg, task['D'], task['A'], task['P'], lows, highs, prior=None,
                             proposal=None, device=str(dev))
            bsim2.run_training(params, states, actions)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            bsim2.run_training(params, states, actions)
            torch.cuda.synchronize()
            dt_mat = time.perf_counter() - t0
        del bsim2
    finally:
        os.environ.pop('BSIG_FUSED_CORR', None)
    res = {'shadowhand_corrdiff_mdnn_1k': {
        'config': 'configs[3] (one GPU, one reference-sized call): ShadowHand-shaped 1000 '
                  'trajectories, summary_corrdiff(F=105002) + MDNN[128,128] K=10, 100 Adam updates x '
                  'minibatch 100 + 6 test evals',
        'fit_trajectories_per_s': n / dt, 'seconds': dt, 'parameters': n_par,
        'ms_per_update': 1e3 * dt / 100, 'final_test_loss': logs['test_loss'][-1],
        'fused_corr_first_layer': bool(fused),
        'ms_per_update_materialised_summary': 1e3 * dt_mat / 100}}
    del bsim, states, actions, params
    torch.cuda.empty_cache()
    return res


def extra_scaled_mode(dev):
    """SURVEY 8.d "scaled mode": the reference's 10 passes over the data, but with a large
    per-GPU minibatch B_g and n_updates = 10 N / B_g in ONE run_training call over all N
    trajectories (the class constants are overridable; `model.run_training` takes them as
    arguments).  At these sizes every layer GEMM (forward, dgrad, wgrad: >= 2^26 MACs) runs
    on the tcgen05 engine (TF32x3, MN-major operands for the backward); the same call on the
    fp32 SIMT engine is timed beside it."""
    import contextlib
    import io
    from bayes_sim_ig.bayes_sim import BayesSim
    res = {}
    n = 1 << 18
    states, actions, params, lows, highs = synth(21, n, TASK)
    states, actions, params = states.to(dev), actions.to(dev), params.to(dev)
    cfg = {'modelClass': 'MDNN', 'summarizerFxn': SUMMARIZER, 'trainTrajLen': TASK['T1'] - 1,
           'components': TASK['K'], 'hiddenLayers': list(HIDDEN), 'lr': LR}
    with contextlib.redirect_stdout(io.StringIO()):
        bsim = BayesSim(cfg, TASK['D'], TASK['A'], TASK['P'], lows, highs, prior=None,
                        proposal=None, device=str(dev))
        feats = bsim.summarizer_fxn(states, actions)
        for b_g in (4096, 16384):
            n_updates = 10 * int(n * 0.8) // b_g
            entry = {}
            for tag, engine in (('tcgen05_tf32x3', -1), ('simt_fp32', 0)):
                bsim.model.gemm_engine = engine
                bsim.model.run_training(feats, params, n_updates, b_g, 0.2)     # capture + warm up
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                logs = bsim.model.run_training(feats, params, n_updates, b_g, 0.2)
                e1.record()
                e1.synchronize()
                ms = e0.elapsed_time(e1)
                entry[tag] = {'ms_per_call': ms, 'ms_per_update': ms / n_updates,
                              'fit_trajectories_per_s': n / (ms * 1e-3),
                              'final_test_loss': logs['test_loss'][-1]}
                bsim.model._plans = {}
                torch.cuda.empty_cache()
            entry['n_updates'] = n_updates
            entry['speedup_tensor_core_vs_simt'] = (entry['simt_fp32']['ms_per_call'] /
                                                    entry['tcgen05_tf32x3']['ms_per_call'])
            res['scaled_mode_B%d' % b_g] = entry
    res['scaled_mode_config'] = ('Cartpole-shaped %d trajectories on one GPU, summary_corrdiff(F=302) '
                                 '+ MDNN[128,128] K=10, 10 passes, minibatch B_g, one call' % n)
    del bsim, states, actions, params, feats
    torch.cuda.empty_cache()
    return res


def extra_signature_mdrff(dev):
    """configs[4], single-GPU slice: depth-3 path-signature summarizer on Cartpole-shaped
    rollouts (C = 6 -> 258 features) + MDRFF fit (reference constants, chunks of 1000) and a
    posterior-sampling sweep with the device RNG."""
    import contextlib
    import io
    from bayes_sim_ig.bayes_sim import BayesSim
    res = {}
    for n in (4096, 65536):
        states, actions, params, lows, highs = synth(13, n, TASK)
        states, actions, params = (states * 0.3).to(dev), actions.to(dev), params.to(dev)
        cfg = {'modelClass': 'MDRFF', 'summarizerFxn': 'summary_signatory', 'trainTrajLen': 20,
               'components': 10, 'hiddenLayers': [128, 128], 'lr': 1e-4}
        with contextlib.redirect_stdout(io.StringIO()):
            bsim = BayesSim(cfg, TASK['D'], TASK['A'], TASK['P'], lows, highs, prior=None,
                            proposal=None, device=str(dev))

            def fit():
                for lo in range(0, n, CHUNK):
                    logs = bsim.run_training(params[lo:lo + CHUNK], states[lo:lo + CHUNK],
                                             actions[lo:lo + CHUNK])
                return logs
            fit()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            logs = fit()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            post = bsim.predict(states[:1], actions[:1])
            sweep = {}
            for ns in (10000, 1000000):
                post.gen(ns, method='philox')
                torch.cuda.synchronize()
                reps = []
                for _ in range(5):
                    t1 = time.perf_counter()
                    smp = post.gen(ns, method='philox')
                    reps.append(time.perf_counter() - t1)
                sweep[str(ns)] = ns / float(np.median(reps))     # median of 5 warm calls
        res['signature_mdrff_%d' % n] = {
            'config': 'configs[4] (one GPU): Cartpole-shaped %d trajectories, summary_signatory '
                      '(depth 3, 258 features) + MDRFF fit, reference constants; posterior sampling '
                      'with the device RNG (samples/s incl. D2H copy)' % n,
            'fit_trajectories_per_s': n / dt, 'seconds': dt,
            'final_test_loss': logs['test_loss'][-1], 'posterior_samples_per_s': sweep,
            'sample_shape': list(smp.shape)}
        del bsim, states, actions, params
        torch.cuda.empty_cache()
    return res


# ----------------------------------------------------------------- multi-GPU extras / checks
def _max_over_ranks(ms, dev, world):
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def dp_parity_check(world, rank, dev):
    """Data-parallel parity where the driver can see it (runs inside `bench.py --gpus N`):
    G ranks x minibatch 16 with injected rows / noise must (a) leave bit-identical replicas
    and (b) reproduce the losses (1e-4) and the parameters (all but < 0.1 % of the entries within
    3e-5 at lr 1e-3, none further than 2*lr per update) of ONE process that trains on the
    concatenated G*16 minibatch (the only semantic difference is the rank-local
    mean inside the 1e-5 eps-noise term).  Bench-shape model (F=302,
    128x128, P=13, K=10: the fused peer-memory exchange path), three updates."""
    import torch.distributed as dist
    from bayes_sim_ig.models.mdnn import MDNN
    from bayes_sim_ig_b200 import data_parallel
    from bayes_sim_ig_b200.models.train_engine import run_training_captured
    import contextlib
    import io
    f, p, k, bg, n_upd = 302, 13, 10, 16, 3
    rs = np.random.RandomState(99)                   # identical on every rank
    n_rows = bg * world
    x = rs.randn(n_rows, f).astype(np.float32)
    y = (0.1 + 1.9 * rs.rand(n_rows, p)).astype(np.float32)
    noise = rs.rand(n_upd, n_rows, p, k).astype(np.float32)
    lows, highs = np.full(p, 0.1), np.full(p, 2.0)

    def make():
        torch.manual_seed(1234)
        return MDNN(f, p, lows, highs, k, False, (128, 128), torch.nn.Tanh, 1e-3, device=str(dev))
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        model = make()
        data_parallel.enable(model)
        lo, hi = rank * bg, (rank + 1) * bg
        inj = dict(idx=np.tile(np.arange(bg), (n_upd, 1)), noise_train=noise[:, lo:hi],
                   noise_test=None)
        logs = run_training_captured(model, torch.from_numpy(x[lo:hi]).to(dev),
                                     torch.from_numpy(y[lo:hi]).to(dev), n_upd, bg, 0.0,
                                     injected=inj)
        plan = list(model._plans.values())[0]
        out['exchange'] = 'p2p' if plan.p2p is not None else 'nccl'
        bits = model.flat_params.view(torch.int32).to(torch.int64)
        h = torch.stack([bits.sum(), (bits * torch.arange(1, bits.numel() + 1, device=dev)).sum()])
        hmin, hmax = h.clone(), h.clone()
        dist.all_reduce(hmin, op=dist.ReduceOp.MIN)
        dist.all_reduce(hmax, op=dist.ReduceOp.MAX)
        out['replicas_bit_identical'] = bool(torch.equal(hmin, hmax))
        loss = torch.tensor(logs['train_loss'], device=dev, dtype=torch.float64)
        dist.all_reduce(loss)
        loss /= world
        if rank == 0:
            single = make()                          # one process, concatenated minibatch
            inj1 = dict(idx=np.tile(np.arange(n_rows), (n_upd, 1)), noise_train=noise,
                        noise_test=None)
            logs1 = run_training_captured(single, torch.from_numpy(x).to(dev),
                                          torch.from_numpy(y).to(dev), n_upd, n_rows, 0.0,
                                          injected=inj1)
            # (data_parallel.enable pads the flat buffer to 4 * world floats: compare the parameters)
            npar = int(single.n_params)
            diff = (single.flat_params[:npar] - model.flat_params[:npar]).abs()
            err = float(diff.max().item())
            frac = float((diff > 3e-5).float().mean().item())
            lerr = float(np.abs(np.asarray(logs1['train_loss']) - loss.cpu().numpy()).max())
            out['max_abs_param_diff_vs_single_process'] = err
            out['frac_params_off_by_more_than_3e-5'] = frac
            out['max_abs_loss_diff_vs_single_process'] = lerr
            # Adam moves an entry whose gradient is below the rounding noise by +-lr with the sign
            # of that noise, so a handful of the 90 k parameters may sit up to 2*lr*n_upd apart
            # between two equivalent computations (rank-local eps-noise mean, summation order):
            # the criterion is the losses of every update (1e-4), the share of entries off by more
            # than 3e-5 (< 0.1 %) and the hard bound 2*lr per update.
            out['ok'] = bool(out['replicas_bit_identical'] and lerr <= 1e-4 and frac < 1e-3 and
                             err <= 2 * 1e-3 * n_upd + 1e-6)
    ok = torch.tensor([1 if out.get('ok', True) and out['replicas_bit_identical'] else 0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    assert int(ok.item()) == 1, 'data-parallel parity check failed: %r' % (out,)
    return out


def multi_gpu_extras(world, rank, dev, flush, peak_gbs):
    """BASELINE.json's multi-GPU configurations at N GPUs (weak scaling, every rank the same
    amount of work, device time = max over ranks):
      * sharded summarizers (no collective): corrdiff and depth-3 signature on Cartpole-shaped
        rollouts, 2^20 trajectories per GPU -> aggregate algorithmic GB/s;
      * configs[3]: ShadowHand-shaped rollouts, 1000 trajectories per GPU and call,
        summary_corrdiff (F = 105 002) + MDNN[128,128] (13.5 M parameters) data parallel
        with an NCCL all-reduce of the 54 MB gradient per update;
      * configs[4]: depth-3 signature + MDRFF fit, 4096 trajectories per GPU, data parallel
        over the fused peer-memory exchange."""
    import contextlib
    import io
    import torch.distributed as dist
    from bayes_sim_ig.bayes_sim import BayesSim
    from bayes_sim_ig_b200 import _lib, data_parallel
    out = {}
    st = lambda: _lib.stream_ptr(dev)

    def timed_kernel(fn, reps=10):
        for _ in range(3):
            fn()
        tot = 0.0
        for _ in range(reps):
            flush()
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            tot += _max_over_ranks(e0.elapsed_time(e1), dev, world)
        return tot / reps
    # ---- sharded summarizers
    n, t1, d, a = 1 << 20, 21, 4, 1
    g = torch.Generator('cpu').manual_seed(500 + rank)
    s = torch.randn(n, t1, d, generator=g).to(dev) * 0.3
    ac = torch.rand(n, t1, a, generator=g).to(dev)
    feats = torch.empty((n, 302), device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    ms = timed_kernel(lambda: _lib.call('bsig_summary_crosscorr', s.data_ptr(), ac.data_ptr(),
                                        feats.data_ptr(), n, t1, t1, d, a, 10, 1, flag.data_ptr(), st()))
    by = n * 4 * (10 * (d + a) + 302)
    out['summary_corrdiff_cartpole_sharded'] = {
        'trajectories_per_gpu': n, 'ms': ms, 'aggregate_GBps': world * by / (ms * 1e-3) / 1e9,
        'per_gpu_frac_of_hbm_peak': by / (ms * 1e-3) / 1e9 / peak_gbs}
    del feats
    sig = torch.empty((n, 258), device=dev)
    ms = timed_kernel(lambda: _lib.call('bsig_signature_fwd', s.data_ptr(), ac.data_ptr(),
                                        sig.data_ptr(), n, t1, t1, t1, d, a, 3, st()))
    by = n * 4 * (t1 * (d + a) + 258)
    out['signature_depth3_cartpole_sharded'] = {
        'trajectories_per_gpu': n, 'ms': ms, 'aggregate_GBps': world * by / (ms * 1e-3) / 1e9,
        'per_gpu_frac_of_hbm_peak': by / (ms * 1e-3) / 1e9 / peak_gbs}
    del s, ac, sig
    torch.cuda.empty_cache()

    def timed_fit(bsim, params, states, actions, n_traj, reps=3):
        chunks = range(0, n_traj, CHUNK)

        def fit():
            for lo in chunks:
                bsim.run_training(params[lo:lo + CHUNK], states[lo:lo + CHUNK], actions[lo:lo + CHUNK])
        with contextlib.redirect_stdout(io.StringIO()):
            fit()
            fit()
            tot = 0.0
            for _ in range(reps):
                dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fit()
                e1.record()
                e1.synchronize()
                tot += _max_over_ranks(e0.elapsed_time(e1), dev, world)
        return tot / reps
    # ---- configs[3]: ShadowHand corrdiff + 13.5 M-parameter MDNN, data parallel
    task = dict(name='shadowhand', D=211, A=20, T1=51, P=32, K=10)
    states, actions, params, lows, highs = synth(2000 + rank, 1000, task)
    states, actions, params = states.to(dev), actions.to(dev), params.to(dev)
    cfg = {'modelClass': 'MDNN', 'summarizerFxn': 'summary_corrdiff', 'trainTrajLen': 50,
           'components': 10, 'hiddenLayers': [128, 128], 'lr': 1e-4}
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        bsim = BayesSim(cfg, task['D'], task['A'], task['P'], lows, highs, prior=None,
                        proposal=None, device=str(dev))
    data_parallel.enable(bsim.model)
    ms = timed_fit(bsim, params, states, actions, 1000)
    plan = list(bsim.model._plans.values())[0]
    out['shadowhand_corrdiff_mdnn_dp'] = {
        'config': 'configs[3]: ShadowHand-shaped 1000 trajectories per GPU and call, '
                  'summary_corrdiff(F=105002) + MDNN[128,128] K=10 (13.5 M parameters), 100 Adam '
                  'updates x minibatch 100 per GPU + 6 test evals, data parallel',
        'gradient_exchange': 'p2p' if plan.p2p is not None else (
            'nccl reduce-scatter (54 MB) -> Adam on the owned 1/%d slice -> all-gather of the weights, '
            'between CUDA graphs' % world if plan._sharded() else
            'nccl all-reduce (54 MB) between two CUDA graphs per update'),
        'first_layer': 'fused cross-correlation (dW stored for the exchange)'
                       if getattr(plan, 'corr', None) is not None else 'materialised summary',
        'ms_per_call': ms, 'ms_per_update': ms / 100,
        'fit_trajectories_per_s': world * 1000 / (ms * 1e-3),
        'parameters': int(bsim.model.flat_params.numel())}
    del bsim, states, actions, params
    torch.cuda.empty_cache()
    # ---- configs[4]: depth-3 signature + MDRFF, data parallel
    n4 = 4096
    states, actions, params, lows, highs = synth(3000 + rank, n4, TASK)
    states, actions, params = (states * 0.3).to(dev), actions.to(dev), params.to(dev)
    cfg = {'modelClass': 'MDRFF', 'summarizerFxn': 'summary_signatory', 'trainTrajLen': 20,
           'components': 10, 'hiddenLayers': [128, 128], 'lr': 1e-4}
    torch.manual_seed(0)
    np.random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        bsim = BayesSim(cfg, TASK['D'], TASK['A'], TASK['P'], lows, highs, prior=None,
                        proposal=None, device=str(dev))
    data_parallel.enable(bsim.model)
    ms = timed_fit(bsim, params, states, actions, n4)
    out['signature_mdrff_dp'] = {
        'config': 'configs[4]: Cartpole-shaped 4096 trajectories per GPU, summary_signatory '
                  '(depth 3, 258 features) + MDRFF fit, reference constants, data parallel',
        'ms_per_step': ms, 'fit_trajectories_per_s': world * n4 / (ms * 1e-3)}
    del bsim, states, actions, params
    torch.cuda.empty_cache()
    return out


def run_b200(args):
    import contextlib
    import io
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py (impl b200) needs a CUDA device'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    from bayes_sim_ig_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from bayes_sim_ig_b200 import build as _b
        _b.build()
    from bayes_sim_ig.bayes_sim import BayesSim
    from bayes_sim_ig_b200.models import train_engine

    states_h, actions_h, params_h, lows, highs = synth(1000 + rank, N_TRAJ, TASK, pin=True)
    states_d, actions_d, params_d = states_h.to(dev), actions_h.to(dev), params_h.to(dev)
    torch.manual_seed(0)
    np.random.seed(0)
    cfg = {'modelClass': 'MDNN', 'summarizerFxn': SUMMARIZER, 'trainTrajLen': TASK['T1'] - 1,
           'components': TASK['K'], 'hiddenLayers': list(HIDDEN), 'lr': LR}
    bsim = BayesSim(cfg, TASK['D'], TASK['A'], TASK['P'], lows, highs, prior=None, proposal=None,
                    device=str(dev))
    if world > 1:
        from bayes_sim_ig_b200 import data_parallel
        data_parallel.enable(bsim.model)
    sink = io.StringIO()

    def step(states, actions, params):
        """The public-API pass: BayesSim.run_training per chunk, predict, MoG.gen."""
        with contextlib.redirect_stdout(sink):
            for lo in range(0, N_TRAJ, CHUNK):
                logs = bsim.run_training(params[lo:lo + CHUNK], states[lo:lo + CHUNK],
                                         actions[lo:lo + CHUNK])
            post = bsim.predict(states[:1], actions[:1])
            smp = post.gen(N_POSTERIOR_SAMPLES, method='philox')   # device RNG (Philox)
        return logs, smp

    flush_buf = torch.zeros(64 * 1024 * 1024, device=dev)     # 256 MiB > 126 MB L2
    flush = lambda: flush_l2(flush_buf)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        tot = 0.0
        for _ in range(k):
            flush()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            tot += e0.elapsed_time(e1)
        t = torch.tensor([tot], device=dev, dtype=torch.float64)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        step(states_d, actions_d, params_d)
    barrier()
    dp_check = dp_parity_check(world, rank, dev) if world > 1 else None
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = _lib.load().bsig_launch_count() + train_engine.replayed_launches()
    ms_res = timed(lambda: step(states_d, actions_d, params_d), args.steps)
    launches = _lib.load().bsig_launch_count() + train_engine.replayed_launches() - launches0
    ms_e2e = timed(lambda: step(states_h, actions_h, params_h), args.steps)
    clocks = sampler.stop()
    value = world * N_TRAJ * args.steps / (ms_res * 1e-3)
    e2e_value = world * N_TRAJ * args.steps / (ms_e2e * 1e-3)
    h2d = int(states_h.numel() + actions_h.numel() + params_h.numel()) * 4 + \
        2 * (N_TRAJ // CHUNK + 1) * 100 * 100 * 8
    d2h = (N_TRAJ // CHUNK + 1) * 13 * 4 + N_POSTERIOR_SAMPLES * TASK['P'] * 4 + \
        TASK['K'] * (1 + 2 * TASK['P']) * 4

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms_res / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': workload_config(world), 'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h, 'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': int(launches)}
    if dp_check is not None:
        line['dp_check'] = dp_check
    if world > 1:
        peak, src = measured_peaks()
        try:
            extra = multi_gpu_extras(world, rank, dev, flush, peak)
        except Exception as exc:                  # an extra must never take the headline down
            extra = {'error': repr(exc)[:300]}
        line['extra'] = extra
    if rank == 0 and world == 1:
        peak, src = measured_peaks()
        roofs = kernel_rooflines(dev, flush, peak, src)
        line['roofline'] = dominant_kernel_roofline(dev, peak, src)
        line['roofline']['kernels'] = roofs     # per-kernel table, under the key the driver keeps
        line['rooflines'] = roofs
        line['extra'] = extra_configs(dev)
        threads = os.cpu_count() or 1
        cpu_pipeline_rate(200, threads)         # warm up the CPU thread pool / allocator
        rate, dt = cpu_pipeline_rate(1000, threads)
        kind = reference_kind()
        line['cpu_baseline'] = {
            'value': rate, 'unit': UNIT, 'cores': threads, 'kind': kind,
            'sample': cpu_sample_note(kind, 1000) + '; %.2f s' % dt}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
