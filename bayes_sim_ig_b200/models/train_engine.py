"""CUDA-graph training engine behind ``MDNN.run_training``.

Restates the reference loop (models/mdnn.py:180-243) as one recorded stream of
kernels per call:

    for step in range(n_updates):
        rows = idx[step]                         # np.random.randint, host (mdnn.py:221)
        [MDRFF: feat = rff(x_train[rows])]       # fused gather + GEMM + sincos
        h_l = tanh(h_{l-1} W_l^T + b_l)          # gather fused into layer 0
        z   = h W_heads^T + b_heads              # the 3-4 heads as ONE GEMM
        loss, dz = fused head epilogue + mixture NLL forward/backward
        dW_heads, db_heads, dh, dW_l, db_l ...   # wgrad / dgrad (+ fused dtanh)
        Adam over the flat parameter buffer      # fresh moments every call (Q9)
        every n_updates//5 steps: test loss on the held-out 20 % (forward only)

Random inputs (minibatch rows, eps-noise uniforms) are generated up front for
the whole call; the 2 x n_log losses and the finiteness flag are read back with
one device->host copy when the graph has finished.
"""
import contextlib
import gc
import os

import numpy as np
import torch

from .. import _lib
from .. import data_parallel

ACT_NONE, ACT_TANH = 0, 1


_REPLAYED = [0]

# Tests only: callable(plan) -> (noise_train [n_updates,B,P,K], noise_test [n_log,n_test,P,K])
# that replaces the device RNG of a public run_training call, so that a whole
# BayesSim.predict refit can be replayed against the reference draw for draw.
NOISE_HOOK = None


def replayed_launches():
    """Kernel launches executed through graph replays so far (bench accounting)."""
    return _REPLAYED[0]


def log_steps(n_updates):
    every = max(n_updates // 5, 1)
    return [e for e in range(n_updates) if e % every == 0 or e + 1 == n_updates]


@contextlib.contextmanager
def _quiet_gc():
    """No cyclic garbage collection while a stream is capturing: a collected TrainPlan of
    a discarded model would destroy its CUDA graph in the middle of the capture, which
    invalidates it (seen when several models are fitted one after the other)."""
    gc.collect()
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was_enabled:
            gc.enable()


class TrainPlan(object):
    """Persistent buffers + the captured graph for one problem shape."""

    def __init__(self, model, n_train, n_test, n_updates, batch, in_dim, corr=None):
        self.model = model
        dev = model.flat_params.device
        self.dev = dev
        self.n_train, self.n_test, self.n_updates, self.batch = n_train, n_test, n_updates, batch
        self.in_dim = in_dim
        p, k = model.output_dim, model.n_gaussians
        self.p, self.k = p, k
        f32 = dict(dtype=torch.float32, device=dev)
        self.logs = log_steps(n_updates)
        n_log = len(self.logs)
        # training rows live in a zero-padded buffer whose pitch is a multiple of 4 floats, so
        # that the persistent kernel can fetch a minibatch row with one 16-byte aligned bulk copy
        self.x_ld = (in_dim + 3) // 4 * 4
        # factored cross-correlation summaries (utils.summarizers.CorrFactors):
        # only the factor rows are stored; the first layer's forward and weight-gradient GEMMs
        # generate the summary tiles on the fly (csrc/corr_layer.cu)
        self.corr = corr                     # None or (s, q, ldf)
        if corr is not None:
            self.fac_train = torch.zeros((n_train, corr[2]), **f32)
            self.fac_test = torch.zeros((max(n_test, 1), corr[2]), **f32)
            self.x_train_buf = self.x_train = self.x_test = None
        else:
            self.x_train_buf = torch.zeros((n_train, self.x_ld), **f32)
            self.x_train = self.x_train_buf[:, :in_dim]
            self.x_test = torch.empty((max(n_test, 1), in_dim), **f32)
        self.y_train = torch.empty((n_train, p), **f32)
        self.y_test = torch.empty((max(n_test, 1), p), **f32)
        self.y_stage = torch.empty((n_train + n_test, p), **f32)
        self.idx = torch.empty((n_updates, batch), dtype=torch.int64, device=dev)
        self.noise_train = torch.empty((n_updates, batch, p, k), **f32)
        self.noise_test = torch.empty((n_log, max(n_test, 1), p, k), **f32)
        trunk = model._trunk_layers()
        self.rff = getattr(model, 'rff', None)
        widths = [lin.weight.shape[0] for lin in trunk]
        feat_dim = model.head_in if not trunk else trunk[0].weight.shape[1]
        self.feat_dim = feat_dim
        nh = model.n_head

        def act_set(rows):
            d = {}
            d['feat'] = torch.empty((rows, feat_dim), **f32) if self.rff is not None else None
            d['h'] = [torch.empty((rows, w), **f32) for w in widths]
            d['z'] = torch.empty((rows, nh), **f32)
            return d
        self.tr = act_set(batch)
        self.te = act_set(max(n_test, 1))
        self.dz = torch.empty((batch, nh), **f32)
        self.dh = [torch.empty((batch, w), **f32) for w in widths]
        self.grads = torch.zeros_like(model.flat_params)
        # data parallel on <= 8 GPUs with a small model: gradients live in peer-mapped
        # buffers and the exchange is fused into the Adam kernel (csrc/p2p.cu)
        self.p2p = data_parallel.p2p_comm_for(model)
        self.exp_avg = torch.zeros_like(model.flat_params)
        self.exp_avg_sq = torch.zeros_like(model.flat_params)
        lib = _lib.load()
        rows_max = max(batch, n_test, 1)
        shapes = [(nh, model.head_in)] + list(zip(widths, [feat_dim] + widths[:-1]))
        if self.rff is not None and corr is None:
            shapes.append((self.rff.freqs.shape[0], in_dim))
        ws_bytes = max(lib.bsig_linear_ws_bytes(rows, n_out, k_in)
                       for rows in (batch, rows_max) for n_out, k_in in shapes)
        self.ws_gemm = torch.empty(int(ws_bytes) + 256, dtype=torch.uint8, device=dev)
        self.ws_mdn = torch.zeros(lib.bsig_mdn_ws_bytes(rows_max), dtype=torch.uint8, device=dev)
        self.corr_adam = False
        if corr is not None:
            n0 = widths[0] if self.rff is None else int(self.rff.freqs.shape[0])
            self.ws_corr = torch.empty(int(lib.bsig_corr_linear_ws_bytes(rows_max, n0, corr[0],
                                                                         corr[1])) + 256,
                                       dtype=torch.uint8, device=dev)
            # single GPU: Adam of the first-layer weight runs in the weight-gradient epilogue
            # (the gradient never reaches HBM); the flat Adam pass then starts behind that weight
            self.corr_adam = (self.rff is None and data_parallel.world_of(model) == 1 and
                              (n0 * in_dim) % 4 == 0)
        # weight-gradient GEMMs run on a side stream, concurrently with the dgrad chain,
        # when every GEMM of the step is a single-launch (workspace-free) kernel
        def single_launch(m_, n_, k_):
            return m_ * n_ <= 128 * 1024 and k_ <= 8192
        self.fork_wgrad = corr is None and all(
            single_launch(batch, n_out, k_in) and single_launch(batch, k_in, n_out)
            and single_launch(n_out, k_in, batch) for n_out, k_in in shapes)
        self.side = torch.cuda.Stream(device=dev) if self.fork_wgrad else None
        # large first layers (>= 2^26 multiply-adds: they run on the tcgen05 engine, which
        # cannot gather rows through TMA): the minibatch rows are gathered ONCE per update
        # into a dense, 16-byte aligned buffer that both the forward and the weight-gradient
        # GEMM read, instead of being staged inside each of them
        self.xg = None
        if (corr is None and self.rff is None and trunk and int(model.gemm_engine) != 0 and
                batch * widths[0] * in_dim >= (1 << 26)):
            self.xg = torch.zeros((batch, self.x_ld), **f32)
        self.loss_buf = torch.zeros(2 * n_log + 1, **f32)   # [train..., test..., scratch]
        self.flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self.graph = None
        # single GPU, two hidden layers: the three weight-gradient GEMMs and Adam are ONE
        # launch (csrc/optim.cu: wgrad3_adam_kernel) on the critical path instead of three
        # side-stream launches + the Adam launch.  Opt-in (BSIG_FUSED_WGRAD=1): measured equal to the default
        # (4.80 vs 4.75 ms per 100 updates, profiles/r2/), it only lowers the launch count
        self.fused_wgrad_adam = (corr is None and self.rff is None and len(trunk) == 2 and
                                 self.p2p is None and batch <= 128 and
                                 data_parallel.world_of(model) == 1 and
                                 os.environ.get('BSIG_FUSED_WGRAD', '0') == '1')
        # single GPU, minibatch-sized layers: Adam rides in the epilogue of the LAST backward
        # kernel (the first layer's weight-gradient GEMM: by then every other gradient is
        # complete and nothing reads the old weights any more), which also updates all other
        # parameters with its idle CTAs -- one launch less on the critical path of an update.
        # BSIG_FUSED_L0_ADAM=0 restores the separate Adam launch.
        self.fused_l0_adam = (corr is None and self.rff is None and len(trunk) >= 1 and
                              self.p2p is None and data_parallel.world_of(model) == 1 and
                              self.fork_wgrad and not self.fused_wgrad_adam and
                              os.environ.get('BSIG_FUSED_L0_ADAM', '1') != '0')
        self._setup_persistent()
        if self.persistent:
            self.fused_l0_adam = False

    # ------------------------------------------------- persistent cluster kernel (default)
    def _setup_persistent(self):
        """Opt-in (BSIG_PERSISTENT=1): all updates between two logging steps as ONE launch of
        csrc/train_persistent.cu (model resident in the shared memory of a 16-CTA cluster),
        for shapes inside the kernel's envelope (bsig_train_persistent_query) on a single GPU.
        Parity-green on hardware (tests/test_gpu_persistent.py) but NOT the default: measured
        at 65-78 us per update against 47 us for the launch-per-GEMM path -- sixteen SMs of
        fp32 FFMA are instruction-issue bound at this problem size (profiles/r2/)."""
        import ctypes
        self.persistent = False
        m = self.model
        if os.environ.get('BSIG_PERSISTENT', '0') != '1' or self.corr is not None:
            return
        if data_parallel.world_of(m) > 1:
            return
        layers, head = self._views()
        dense = layers + [head]
        if len(dense) > 3:
            return
        f32 = dict(dtype=torch.float32, device=self.dev)
        d = _lib.TpDesc()
        d.n_layers = len(dense)
        flat0 = m.flat_params.data_ptr()
        for i, lay in enumerate(dense):
            d.in_dim[i], d.out_dim[i] = int(lay['k']), int(lay['n'])
            d.w_off[i] = (lay['w'].data_ptr() - flat0) // 4
            d.b_off[i] = (lay['b'].data_ptr() - flat0) // 4
        d.batch, d.p, d.k = self.batch, self.p, self.k
        d.full_cov = 1 if m.full_covariance else 0
        if self.rff is not None:
            # random Fourier features are not trained: computed once per call for the whole
            # training split (instead of once per minibatch) and handed to the kernel as x
            self.feat_train = torch.zeros((self.n_train, self.feat_dim), **f32)
            d.x, d.ldx = self.feat_train.data_ptr(), self.feat_dim
        else:
            d.x, d.ldx = self.x_train_buf.data_ptr(), self.x_ld
        need, smem = ctypes.c_int64(0), ctypes.c_int64(0)
        lib = _lib.load()
        if lib.bsig_train_persistent_query(ctypes.byref(d), ctypes.byref(need),
                                           ctypes.byref(smem)) != 0:
            self.persistent_reason = _lib.last_error()
            return
        self.tp_scratch = torch.zeros(int(need.value) + 64, **f32)
        slots = np.full(self.n_updates, -1, dtype=np.int32)
        for slot, step in enumerate(self.logs):
            slots[step] = slot
        self.tp_slots = torch.from_numpy(slots).to(self.dev)
        t = np.arange(1, self.n_updates + 1, dtype=np.float64)
        b1, b2 = 0.9, 0.999
        coef = np.stack([float(m.lr) / (1.0 - b1 ** t), 1.0 / np.sqrt(1.0 - b2 ** t)], axis=1)
        self.tp_coef = torch.from_numpy(coef.astype(np.float32)).to(self.dev)
        d.y = self.y_train.data_ptr()
        d.idx = self.idx.data_ptr()
        d.noise = self.noise_train.data_ptr()
        d.params = flat0
        d.exp_avg, d.exp_avg_sq = self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
        d.scratch, d.scratch_floats = self.tp_scratch.data_ptr(), self.tp_scratch.numel()
        d.loss_buf = self.loss_buf.data_ptr()
        d.loss_slot = self.tp_slots.data_ptr()
        d.adam_coef = self.tp_coef.data_ptr()
        d.flag = self.flag.data_ptr()
        d.beta1, d.beta2, d.eps = b1, b2, 1e-8
        self.tp_prof = None
        if os.environ.get('BSIG_TP_PROF') == '1':       # per-phase cycle counters (profiles/)
            self.tp_prof = torch.zeros(32, dtype=torch.int64, device=self.dev)
            d.prof = self.tp_prof.data_ptr()
        self.tp_desc = d
        self.tp_smem = int(smem.value)
        self.tp_lr = float(m.lr)
        self.persistent = True

    def _enqueue_segment(self, step0, step1, st):
        import ctypes
        _lib.call('bsig_train_persistent', ctypes.byref(self.tp_desc), step0, step1, st)

    def _enqueue_rff_train_features(self, st):
        coeff = self.rff.coeff()
        _lib.call('bsig_rff_features', self.x_train_buf.data_ptr(), self.x_ld, None,
                  coeff.data_ptr(), self.feat_train.data_ptr(), self.n_train, self.rff.d,
                  coeff.shape[0], float(self.rff.a), int(self.rff.gemm_engine),
                  self.ws_gemm.data_ptr(), self.ws_gemm.numel(), st)

    # ------------------------------------------------------------- kernel sequence
    def _views(self, gbase=None):
        """Per-layer weight / bias tensors and the DEVICE POINTERS of their gradients
        inside the flat gradient buffer that starts at ``gbase``."""
        m = self.model
        flat = m.flat_params
        gbase = self.grads.data_ptr() if gbase is None else gbase
        out, off = [], 0
        for lin in m._trunk_layers():
            nw, nb = lin.weight.numel(), lin.bias.numel()
            out.append(dict(w=flat[off:off + nw], b=flat[off + nw:off + nw + nb],
                            dw=gbase + 4 * off, db=gbase + 4 * (off + nw),
                            n=lin.weight.shape[0], k=lin.weight.shape[1]))
            off += nw + nb
        nhw = m.n_head * m.head_in
        head = dict(w=flat[m._head_w_off:m._head_w_off + nhw],
                    b=flat[m._head_b_off:m._head_b_off + m.n_head],
                    dw=gbase + 4 * m._head_w_off, db=gbase + 4 * m._head_b_off,
                    n=m.n_head, k=m.head_in)
        return out, head

    def _forward(self, acts, x, rows, n_rows, st, ld=None):
        """x (optionally row-gathered, row pitch ld) -> acts['z']; returns the head input."""
        m = self.model
        eng = int(m.gemm_engine)
        wsp, wsn = self.ws_gemm.data_ptr(), self.ws_gemm.numel()
        rows_p = None if rows is None else rows.data_ptr()
        layers, head = self._views()
        cur, ld = x, (x.shape[1] if ld is None else ld)
        if self.rff is not None:
            coeff = self.rff.coeff()
            self._coeff_ref = coeff        # the captured graph holds this tensor's address
            if self.corr is not None:
                cs, cq, cld = self.corr        # projection of the never-materialised summary
                _lib.call('bsig_corr_rff_features', cur.data_ptr(), cld, rows_p, cs, cq,
                          coeff.data_ptr(), acts['feat'].data_ptr(), n_rows, coeff.shape[0],
                          float(self.rff.a), self.ws_corr.data_ptr(), self.ws_corr.numel(), st)
            else:
                _lib.call('bsig_rff_features', cur.data_ptr(), ld, rows_p, coeff.data_ptr(),
                          acts['feat'].data_ptr(), n_rows, self.rff.d, coeff.shape[0],
                          float(self.rff.a), int(self.rff.gemm_engine), wsp, wsn, st)
            cur, ld, rows_p = acts['feat'], acts['feat'].shape[1], None
        for li, lay in enumerate(layers):
            if li == 0 and self.corr is not None:
                cs, cq, cld = self.corr
                _lib.call('bsig_corr_linear_fwd', cur.data_ptr(), cld, rows_p, cs, cq,
                          lay['w'].data_ptr(), lay['b'].data_ptr(), acts['h'][0].data_ptr(),
                          n_rows, lay['n'], ACT_TANH, self.ws_corr.data_ptr(),
                          self.ws_corr.numel(), st)
                cur, ld, rows_p = acts['h'][0], lay['n'], None
                continue
            _lib.call('bsig_linear_fwd', cur.data_ptr(), ld, rows_p, lay['w'].data_ptr(),
                      lay['b'].data_ptr(), acts['h'][li].data_ptr(), n_rows, lay['n'], lay['k'],
                      ACT_TANH, eng, wsp, wsn, st)
            cur, ld, rows_p = acts['h'][li], lay['n'], None
        _lib.call('bsig_linear_fwd', cur.data_ptr(), ld, rows_p, head['w'].data_ptr(),
                  head['b'].data_ptr(), acts['z'].data_ptr(), n_rows, head['n'], head['k'],
                  ACT_NONE, eng, wsp, wsn, st)
        return cur, ld, rows_p

    def _enqueue_step(self, step, st):
        m = self.model
        eng = int(m.gemm_engine)
        wsp, wsn = self.ws_gemm.data_ptr(), self.ws_gemm.numel()
        b, p, k = self.batch, self.p, self.k
        rows = self.idx[step]
        layers, head = self._views(self.p2p.local_grads(step) if self.p2p is not None else None)
        slot = self.logs.index(step) if step in self.logs else None
        n_log = len(self.logs)
        loss_ptr = self.loss_buf.data_ptr() + 4 * (slot if slot is not None else 2 * n_log)
        if self.corr is not None:
            x0, x0_rows = self.fac_train, rows
        elif self.xg is not None:
            _lib.call('bsig_gather_rows', self.x_train_buf.data_ptr(), self.x_ld, rows.data_ptr(),
                      self.xg.data_ptr(), b, self.x_ld, st)
            x0, x0_rows = self.xg, None
        else:
            x0, x0_rows = self.x_train_buf, rows
        hin, hin_ld, hin_rows = self._forward(self.tr, x0, x0_rows, b, st, ld=self.x_ld)
        _lib.call('bsig_mdn_nll_fused', self.tr['z'].data_ptr(), self.noise_train[step].data_ptr(),
                  self.y_train.data_ptr(), rows.data_ptr(), loss_ptr, self.dz.data_ptr(),
                  b, p, k, 1 if m.full_covariance else 0, self.ws_mdn.data_ptr(),
                  self.ws_mdn.numel(), self.flag.data_ptr(), st)
        # backward: the dgrad chain stays on the main stream; each wgrad (+ fused bias
        # gradient) is forked to the side stream as soon as its dY exists
        main = torch.cuda.current_stream(self.dev)
        side = self.side if self.fork_wgrad else None

        def wgrad(dy, xin, xld, xrows, lay, n_out, k_in):
            if self.fused_wgrad_adam:
                return                          # formed by bsig_wgrad3_adam_step in _enqueue_update
            if side is not None:
                side.wait_stream(main)
                stream = side.cuda_stream
            else:
                stream = st
            _lib.call('bsig_linear_wgrad', dy.data_ptr(), xin.data_ptr(), xld, xrows,
                      lay['dw'], lay['db'], b, n_out, k_in, eng, wsp, wsn,
                      stream)

        wgrad(self.dz, hin, hin_ld, hin_rows, head, head['n'], head['k'])
        dcur = self.dz
        nxt = head
        for li in reversed(range(len(layers))):
            lay = layers[li]
            # d pre-activation of layer li = (dcur @ W_next) * (1 - h_li^2)
            _lib.call('bsig_linear_dgrad', dcur.data_ptr(), nxt['w'].data_ptr(),
                      self.tr['h'][li].data_ptr(), self.dh[li].data_ptr(), b, nxt['n'], nxt['k'],
                      ACT_TANH, eng, wsp, wsn, st)
            if li == 0 and self.corr is not None:
                self._enqueue_corr_wgrad(step, lay, rows, st)
                dcur, nxt = self.dh[li], lay
                continue
            if li == 0 and self.fused_l0_adam:
                # (side stream, in order behind the other weight gradients; dgrad of layer 1,
                # the last reader of any weight, precedes it through wait_stream)
                side.wait_stream(main)
                _lib.call('bsig_linear_wgrad_adam', self.dh[0].data_ptr(),
                          self.x_train_buf.data_ptr(), self.x_ld, rows.data_ptr(), b, lay['n'],
                          lay['k'], m.flat_params.data_ptr(), self.grads.data_ptr(),
                          self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                          m.flat_params.numel(), step + 1, float(m.lr), 0.9, 0.999, 1e-8,
                          side.cuda_stream)
                dcur, nxt = self.dh[li], lay
                continue
            if li > 0:
                xin, xld, xrows = self.tr['h'][li - 1], layers[li - 1]['n'], None
            elif self.rff is not None:
                xin, xld, xrows = self.tr['feat'], self.feat_dim, None
            elif self.xg is not None:
                xin, xld, xrows = self.xg, self.x_ld, None
            else:
                xin, xld, xrows = self.x_train_buf, self.x_ld, rows.data_ptr()
            wgrad(self.dh[li], xin, xld, xrows, lay, lay['n'], lay['k'])
            dcur, nxt = self.dh[li], lay
        if side is not None and not self.fused_wgrad_adam:
            main.wait_stream(side)

    def _enqueue_corr_wgrad(self, step, lay, rows, st):
        """Weight gradient of the first layer from the factored summary rows; single GPU:
        with Adam of that weight in the epilogue (fresh moments per call, step count as in
        _enqueue_update); data parallel: the gradient is stored for the exchange."""
        m = self.model
        cs, cq, cld = self.corr
        nw0 = lay['n'] * lay['k']
        _lib.call('bsig_linear_colsum', self.dh[0].data_ptr(), lay['db'], self.batch, lay['n'], st)
        if self.corr_adam:
            _lib.call('bsig_corr_linear_wgrad', self.dh[0].data_ptr(), self.fac_train.data_ptr(),
                      cld, rows.data_ptr(), cs, cq, self.batch, lay['n'], None,
                      lay['w'].data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                      step + 1, float(m.lr), 0.9, 0.999, 1e-8, 1.0, st)
        else:
            _lib.call('bsig_corr_linear_wgrad', self.dh[0].data_ptr(), self.fac_train.data_ptr(),
                      cld, rows.data_ptr(), cs, cq, self.batch, lay['n'], lay['dw'],
                      None, None, None, 1, 0.0, 0.9, 0.999, 1e-8, 1.0, st)
        assert nw0 == m._offsets[1]       # the first-layer weight opens the flat buffer

    def _enqueue_update(self, step, st):
        """Second half of a step: Adam (gradient mean over the replicas folded in) and,
        on logging steps, the held-out loss."""
        m = self.model
        p, k = self.p, self.k
        slot = self.logs.index(step) if step in self.logs else None
        n_log = len(self.logs)
        world = data_parallel.world_of(m)
        if self.fused_wgrad_adam:
            layers, head = self._views()
            h1, h2 = self.tr['h']
            nw0, nw1 = layers[0]['n'] * layers[0]['k'], layers[1]['n'] * layers[1]['k']
            off1 = nw0 + layers[0]['n']
            _lib.call('bsig_wgrad3_adam_step',
                      self.dh[0].data_ptr(), self.x_train_buf.data_ptr(), self.x_ld,
                      self.idx[step].data_ptr(), layers[0]['n'], layers[0]['k'], 0, nw0,
                      self.dh[1].data_ptr(), h1.data_ptr(), layers[1]['n'], layers[1]['k'],
                      off1, off1 + nw1,
                      self.dz.data_ptr(), h2.data_ptr(), head['n'], head['k'],
                      m._head_w_off, m._head_b_off,
                      m.flat_params.data_ptr(), self.exp_avg.data_ptr(),
                      self.exp_avg_sq.data_ptr(), self.batch, step + 1, float(m.lr), 0.9, 0.999,
                      1e-8, st)
        elif self.fused_l0_adam:
            pass                                # applied by bsig_linear_wgrad_adam in _enqueue_step
        elif self.p2p is not None:
            # one kernel: all-reduce over NVLink peer memory (1/world folded in) + Adam
            self.p2p.adam_allreduce(m, self.exp_avg, self.exp_avg_sq, step, step + 1, st)
        else:
            # (factored first layer, single GPU: its weight was updated in the wgrad epilogue)
            skip = m._offsets[1] if self.corr_adam else 0
            _lib.call('bsig_adam_step', m.flat_params.data_ptr() + 4 * skip,
                      self.grads.data_ptr() + 4 * skip, self.exp_avg.data_ptr() + 4 * skip,
                      self.exp_avg_sq.data_ptr() + 4 * skip, m.flat_params.numel() - skip,
                      step + 1, float(m.lr), 0.9, 0.999, 1e-8, 1.0 / world, st)
        self._enqueue_eval(step, st)

    def _enqueue_eval(self, step, st):
        """Held-out loss after update `step` (logging steps only; mdnn.py:236-241)."""
        m = self.model
        p, k = self.p, self.k
        slot = self.logs.index(step) if step in self.logs else None
        n_log = len(self.logs)
        if slot is not None and self.n_test > 0:
            self._forward(self.te, self.fac_test if self.corr is not None else self.x_test,
                          None, self.n_test, st)
            _lib.call('bsig_mdn_nll_fused', self.te['z'].data_ptr(),
                      self.noise_test[slot].data_ptr(), self.y_test.data_ptr(), None,
                      self.loss_buf.data_ptr() + 4 * (n_log + slot), None, self.n_test, p, k,
                      1 if m.full_covariance else 0, self.ws_mdn.data_ptr(), self.ws_mdn.numel(),
                      self.flag.data_ptr(), st)

    def enqueue_all(self):
        st = _lib.stream_ptr(self.dev)
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        if self.persistent:
            self.tp_scratch.zero_()           # Adam moments of the persistent kernel (Q9)
            if self.rff is not None:
                self._enqueue_rff_train_features(st)
            prev = 0
            for step in self.logs:            # one launch per stretch between two logging steps
                self._enqueue_segment(prev, step + 1, st)
                self._enqueue_eval(step, st)
                prev = step + 1
            if prev < self.n_updates:
                self._enqueue_segment(prev, self.n_updates, st)
            return
        for step in range(self.n_updates):
            self._enqueue_step(step, st)
            self._enqueue_update(step, st)

    def warm_up(self):
        """Run update 0 once OUTSIDE graph capture, then undo its effects.  The first
        launch of a kernel loads its code lazily and may grow the context's local-memory
        pool -- neither is allowed while a stream is capturing (a first call that went
        straight into capture failed for the ShadowHand-sized layers, whose engine had
        never been launched before).  Step 0 is a logging step, so the held-out
        evaluation kernels are covered too."""
        m = self.model
        saved = m.flat_params.detach().clone()
        saved_loss, saved_flag = self.loss_buf.clone(), self.flag.clone()
        st = _lib.stream_ptr(self.dev)
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        if self.persistent:
            self.tp_scratch.zero_()
            if self.rff is not None:
                self._enqueue_rff_train_features(st)
            self._enqueue_segment(0, 1, st)
            self._enqueue_eval(0, st)
        elif self.p2p is not None:
            # Peer-memory exchange: the warm-up stays RANK-LOCAL.  Plan keys contain the
            # rank-local split sizes, so with uneven shards one rank can build a new plan while
            # its peers replay cached ones -- a rendezvous here (the exchange kernel advances a
            # device-side epoch and waits for every peer) would hang or desynchronise them.
            # The exchange kernel's code is loaded without launching it; update 0 runs with the
            # plain local Adam on this rank's own gradient buffer (parity 0, which the graph's
            # update 0 overwrites before it publishes anything) and is undone below.
            _lib.call('bsig_p2p_preload')
            self._enqueue_step(0, st)
            _lib.call('bsig_adam_step', m.flat_params.data_ptr(), self.p2p.local_grads(0),
                      self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), m.flat_params.numel(),
                      1, float(m.lr), 0.9, 0.999, 1e-8, 1.0, st)
            self._enqueue_eval(0, st)
        else:
            # (NCCL data parallel as well: no collective here -- the warm-up only has to launch
            # every kernel of the graphs once, its result is undone below, and a collective
            # would have to be matched by peers that may be replaying cached plans)
            self._enqueue_step(0, st)
            self._enqueue_update(0, st)
        torch.cuda.synchronize(self.dev)
        m.flat_params.copy_(saved)
        self.loss_buf.copy_(saved_loss)
        self.flag.copy_(saved_flag)

    def capture(self):
        graph = torch.cuda.CUDAGraph()
        with _quiet_gc(), torch.cuda.graph(graph, capture_error_mode='relaxed'):
            self.enqueue_all()
        self.graph = graph

    # ---- data parallel: two graphs per step with the NCCL all-reduce launched between
    # them (collectives stay out of graph capture; 2 graph launches + 1 collective per
    # step keep the host far ahead of the ~50 us the device needs)
    # Sharded exchange (ZeRO-1 style) whenever the flat buffer splits evenly: reduce-scatter of
    # the gradients, Adam on this rank's 1/world slice only (a 13.5 M-parameter Adam pass is
    # 74 us of HBM traffic: divided by the world size instead of replicated), all-gather of the
    # updated weights -- the same bytes on the wire as the all-reduce it replaces.
    # Default from 4 ranks up (measured on configs[3]: 0.428 vs 0.452 ms per update at 8 GPUs,
    # 0.356 vs 0.347 at 2, where halving the Adam pass does not pay for the second collective);
    # BSIG_DP_SHARDED=1 / 0 forces it on / off.
    def _sharded(self):
        world = data_parallel.world_of(self.model)
        n = self.model.flat_params.numel()
        want = os.environ.get('BSIG_DP_SHARDED', '1' if world >= 4 else '0') != '0'
        return world > 1 and n % (4 * world) == 0 and want

    def _enqueue_adam_shard(self, step, st):
        m = self.model
        world = data_parallel.world_of(m)
        shard = m.flat_params.numel() // world
        off = 4 * shard * torch.distributed.get_rank(getattr(m, '_dp_group', None))
        _lib.call('bsig_adam_step', m.flat_params.data_ptr() + off, self.grads.data_ptr() + off,
                  self.exp_avg.data_ptr() + off, self.exp_avg_sq.data_ptr() + off, shard,
                  step + 1, float(m.lr), 0.9, 0.999, 1e-8, 1.0 / world, st)

    def capture_dp(self):
        self.dp_graphs = []
        pool = None
        sharded = self._sharded()
        for step in range(self.n_updates):
            halves = [self._enqueue_step]
            if sharded:
                halves += [self._enqueue_adam_shard, self._enqueue_eval]
            else:
                halves += [self._enqueue_update]
            graphs = []
            for half in halves:
                if half == self._enqueue_eval and not (step in self.logs and self.n_test > 0):
                    graphs.append(None)
                    continue
                g = torch.cuda.CUDAGraph()
                with _quiet_gc(), torch.cuda.graph(g, pool=pool, capture_error_mode='relaxed'):
                    half(step, _lib.stream_ptr(self.dev))
                pool = g.pool()
                graphs.append(g)
            self.dp_graphs.append(graphs)

    def replay_dp(self):
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        m = self.model
        if not self._sharded():
            for fwd_bwd, update in self.dp_graphs:
                fwd_bwd.replay()
                data_parallel.allreduce_gradients(m, self.grads)
                update.replay()
            return
        group = getattr(m, '_dp_group', None)
        world = data_parallel.world_of(m)
        rank = torch.distributed.get_rank(group)
        shard = m.flat_params.numel() // world
        g_mine = self.grads[rank * shard:(rank + 1) * shard]
        p_mine = m.flat_params.data[rank * shard:(rank + 1) * shard]
        for fwd_bwd, adam, evalg in self.dp_graphs:
            fwd_bwd.replay()
            torch.distributed.reduce_scatter_tensor(g_mine, self.grads, group=group)   # in place
            adam.replay()
            torch.distributed.all_gather_into_tensor(m.flat_params.data, p_mine, group=group)
            if evalg is not None:
                evalg.replay()


def _stage_inputs(plan, model, x_data, y_data):
    n_train, n_test = plan.n_train, plan.n_test
    dev = plan.dev
    y_data = y_data.detach()
    if plan.corr is not None:
        fac = x_data.fac.detach()
        plan.fac_train.copy_(fac[:n_train], non_blocking=True)
        if n_test > 0:
            plan.fac_test[:n_test].copy_(fac[n_train:], non_blocking=True)
    else:
        x_data = x_data.detach()
        plan.x_train.copy_(x_data[:n_train], non_blocking=True)
        if n_test > 0:
            plan.x_test[:n_test].copy_(x_data[n_train:], non_blocking=True)
    y_dev = y_data.to(device=dev, dtype=torch.float32).contiguous()
    st = _lib.stream_ptr(dev)
    if model.output_lows is not None:
        _lib.call('bsig_normalize_rows', y_dev.data_ptr(), model.output_lows.data_ptr(),
                  model.output_highs.data_ptr(), plan.y_stage.data_ptr(), y_dev.shape[0],
                  y_dev.shape[1], st)
    else:
        plan.y_stage.copy_(y_dev)
    plan.y_train.copy_(plan.y_stage[:n_train])
    if n_test > 0:
        plan.y_test[:n_test].copy_(plan.y_stage[n_train:])


def _corr_layout(model, x_data, batch_size, n_test):
    """(s, q, ldf) when ``x_data`` is a CorrFactors object the fused first-layer kernels can
    consume for this model (an MDNN trunk whose first layer is at most 128 wide, or an MDRFF
    with at most 128 frequency rows; minibatch of at most 128 rows, csrc/corr_layer.cu), else
    None."""
    if not hasattr(x_data, 'fac') or not hasattr(x_data, 'materialize'):
        return None
    trunk = model._trunk_layers()
    rff = getattr(model, 'rff', None)
    if rff is not None:
        n0 = int(rff.freqs.shape[0])           # MDRFF: the projection x . (freqs / sigma)^T
        if rff.d != x_data.shape[1] or not getattr(rff, 'to_features', None) == rff._to_cos_sin_features:
            return None
    elif trunk:
        n0 = trunk[0].weight.shape[0]
    else:
        return None
    ok = _lib.load().bsig_corr_linear_applicable(batch_size, max(batch_size, n_test, 1), n0,
                                                 x_data.s, x_data.q)
    return (x_data.s, x_data.q, int(x_data.fac.shape[1])) if ok else None


def run_training_captured(model, x_data, y_data, n_updates, batch_size, test_frac=0.2,
                          use_graph=True, injected=None):
    """See MDNN.run_training.  ``injected`` (tests only) = dict(idx=[n_updates,B]
    int64, noise_train=[n_updates,B,P,K], noise_test=[n_log,n_test,P,K]) replaces
    the generated random inputs so that runs can be compared draw for draw."""
    assert (x_data.shape[0] == y_data.shape[0])
    model._ensure_flat()
    model.train()
    dev = model.flat_params.device
    n_tot = x_data.shape[0]
    n_train = max(int(n_tot * (1.0 - test_frac)), 1)
    n_test = n_tot - n_train
    in_dim = x_data.shape[1]
    dp = data_parallel.world_of(model) > 1
    corr = _corr_layout(model, x_data, batch_size, n_test)
    if corr is None and hasattr(x_data, 'materialize'):
        x_data = x_data.materialize()          # shape outside the fused kernels' envelope
    key = (n_train, n_test, n_updates, batch_size, in_dim, bool(use_graph), dp,
           float(model.lr), int(model.gemm_engine), corr)
    with torch.cuda.device(dev):
        plan = model._plans.get(key)
        if plan is None:
            plan = TrainPlan(model, n_train, n_test, n_updates, batch_size, in_dim, corr)
            model._plans[key] = plan
        _stage_inputs(plan, model, x_data, y_data)
        if injected is None:
            # same generator and call pattern as the reference: numpy's global RNG,
            # one randint(0, n_train, batch) per update (mdnn.py:221)
            # (one vectorised call draws the identical stream and leaves the identical
            # generator state as n_updates calls of size batch_size -- checked in
            # tests/test_cabi_and_host.py)
            ids = np.random.randint(0, n_train, (n_updates, batch_size))
            plan.idx.copy_(torch.from_numpy(ids.astype(np.int64)), non_blocking=False)
            if NOISE_HOOK is None:
                plan.noise_train.uniform_(0.0, 1.0)
                plan.noise_test.uniform_(0.0, 1.0)
            else:
                tr, te = NOISE_HOOK(plan)
                plan.noise_train.copy_(torch.as_tensor(tr).reshape(plan.noise_train.shape))
                if n_test > 0:
                    plan.noise_test.copy_(torch.as_tensor(te).reshape(plan.noise_test.shape))
        else:
            plan.idx.copy_(torch.as_tensor(injected['idx'], dtype=torch.int64))
            plan.noise_train.copy_(torch.as_tensor(injected['noise_train']).reshape(
                plan.noise_train.shape))
            if n_test > 0:
                plan.noise_test.copy_(torch.as_tensor(injected['noise_test']).reshape(
                    plan.noise_test.shape))
        plan.flag.zero_()
        if n_test == 0:
            plan.loss_buf.fill_(float('nan'))
        if dp and plan.p2p is not None and plan.n_updates % 2 == 1:
            # gradient buffers alternate with the update parity; after an odd number of
            # updates the next call starts on the buffer the last exchange is still
            # reading on slower ranks: all ranks must have retired the previous call
            torch.distributed.barrier(group=getattr(model, '_dp_group', None))
        if use_graph and dp and plan.p2p is None:
            if getattr(plan, 'dp_graphs', None) is None:
                plan.warm_up()
                before = _lib.load().bsig_launch_count()
                plan.capture_dp()
                plan.launches_per_replay = _lib.load().bsig_launch_count() - before
            plan.replay_dp()
            _REPLAYED[0] += int(getattr(plan, 'launches_per_replay', 0))
        elif use_graph:
            if (plan.graph is not None and plan.rff is not None and
                    plan.rff.coeff() is not getattr(plan, '_coeff_ref', None)):
                # freqs / sigma were edited since the capture: the cached coefficient tensor
                # the graph points at has been replaced (RFF.coeff) -> record the call again
                plan.graph = None
            fresh = plan.graph is None
            if fresh:
                plan.warm_up()
                before = _lib.load().bsig_launch_count()
                plan.capture()
                plan.launches_per_replay = _lib.load().bsig_launch_count() - before
            with _lib.nvtx_range('bsig.train_graph'):
                plan.graph.replay()
            _REPLAYED[0] += int(getattr(plan, 'launches_per_replay', 0))
        else:
            if plan.persistent:
                plan.enqueue_all()
            else:
                plan.exp_avg.zero_()
                plan.exp_avg_sq.zero_()
            st = _lib.stream_ptr(dev)
            for step in range(plan.n_updates if not plan.persistent else 0):
                plan._enqueue_step(step, st)
                if plan.p2p is None:
                    data_parallel.allreduce_gradients(model, plan.grads)
                plan._enqueue_update(step, st)
        n_log = len(plan.logs)
        host = torch.cat([plan.loss_buf[:2 * n_log], plan.flag.float()]).cpu().numpy()
    if plan.p2p is not None:
        missing = plan.p2p.peer_timeout()
        if missing:
            raise _lib.BsigError('data-parallel exchange timed out waiting for rank %d: a peer '
                                 'process died or stalled; the results of this call are invalid'
                                 % (missing - 1))
    assert (host[-1] == 0), 'non-finite value in MDNN training (forward / log-likelihood)'
    train_loss = [float(v) for v in host[:n_log]]
    test_loss = [float(v) for v in host[n_log:2 * n_log]]
    for tr, te in zip(train_loss, test_loss):
        print(f'loss: train {tr:0.4f} test {te:0.4f}')
    return {'train_loss': train_loss, 'test_loss': test_loss}
