"""GPU parity of the dense-layer engines (SIMT small / tiled, tcgen05 TF32 and
TF32x3) against a float64 torch reference, through the C ABI."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
SIMT, TC_TF32, TC_X3 = 0, 1, 2
# max-norm relative tolerances.  fp32 FFMA: 1e-5.  The 3xTF32 split recovers the
# operand mantissas, but the tensor core's fp32 accumulator truncates on every
# add, an error that grows linearly with K (measured 1.9e-5 at K = 680):
# 2e-5 + 4e-8*K.  Plain TF32 has 10-bit operand mantissas: 4e-3.


def tol_for(engine, k):
    return {SIMT: 1e-5, TC_X3: 2e-5 + 4e-8 * k, TC_TF32: 4e-3}[engine]


def _lib():
    from bayes_sim_ig_b200 import _lib
    return _lib


def _ws(m, n, k):
    nbytes = _lib().load().bsig_linear_ws_bytes(m, n, k)
    return torch.empty(max(nbytes, 16), dtype=torch.uint8, device=DEV)


def _rel(got, ref):
    return float((got.double() - ref).abs().max() / ref.abs().max())


SHAPES = [(300, 100, 680), (128, 128, 128), (1000, 270, 128), (65, 20, 36), (257, 128, 100),
          (100, 128, 302), (4096, 128, 1292), (100, 128, 11804),
          (38001, 128, 160)]     # >= 148 tiles of 256 rows: the plain TF32 engine's two-accumulator form


@pytest.mark.parametrize('engine', [SIMT, TC_TF32, TC_X3])
@pytest.mark.parametrize('shape', SHAPES)
def test_linear_forward_engines(engine, shape):
    lib = _lib()
    m, n, k = shape
    g = torch.Generator('cpu').manual_seed(m + n + k)
    x = torch.randn(m, k, generator=g).to(DEV)
    w = (torch.randn(n, k, generator=g) / np.sqrt(k)).to(DEV)
    b = torch.randn(n, generator=g).to(DEV)
    y = torch.empty(m, n, device=DEV)
    ws = _ws(m, n, k)
    for act in (0, 1):
        lib.call('bsig_linear_fwd', x.data_ptr(), k, None, w.data_ptr(), b.data_ptr(), y.data_ptr(),
                 m, n, k, act, engine, ws.data_ptr(), ws.numel(), lib.stream_ptr(DEV))
        ref = x.double() @ w.double().T + b.double()
        ref = torch.tanh(ref) if act else ref
        err = _rel(y, ref)
        assert err < tol_for(engine, k), (engine, shape, act, err)


@pytest.mark.parametrize('engine', [SIMT, TC_TF32, TC_X3])
@pytest.mark.parametrize('shape', [(300, 100, 680), (4099, 100, 1292), (70, 10, 40), (512, 100, 52),
                                   (40003, 100, 300)])
def test_rff_features_engines(engine, shape):
    lib = _lib()
    m, nf, d = shape
    g = torch.Generator('cpu').manual_seed(d)
    x = torch.randn(m, d, generator=g).to(DEV)
    coeff = (torch.randn(nf, d, generator=g) / 4.0).to(DEV)
    out = torch.empty(m, 2 * nf, device=DEV)
    ws = _ws(m, nf, d)
    scale = float(np.sqrt(1.0 / nf))
    lib.call('bsig_rff_features', x.data_ptr(), d, None, coeff.data_ptr(), out.data_ptr(), m, d, nf,
             scale, engine, ws.data_ptr(), ws.numel(), lib.stream_ptr(DEV))
    inner = x.double() @ coeff.double().T
    ref = scale * torch.cat([torch.cos(inner), torch.sin(inner)], dim=1)
    # an error in the phase `inner` (|inner| up to ~40, d fp32 products) passes
    # straight into the feature: absolute tolerance on the phase, times the scale
    phase_tol = {SIMT: 1e-4, TC_X3: 1e-4 + 5e-7 * d, TC_TF32: 0.15}[engine]
    err = float((out.double() - ref).abs().max())
    assert err < phase_tol * scale, (engine, shape, err / scale)


@pytest.mark.parametrize('shape', [(100, 128, 302), (100, 270, 128), (257, 64, 100), (1000, 128, 700),
                                   (100, 128, 20002)])
def test_dgrad_wgrad_gather(shape):
    lib = _lib()
    m, n, k = shape
    g = torch.Generator('cpu').manual_seed(7 * m + n)
    xfull = torch.randn(m + 50, k, generator=g).to(DEV)
    rows = torch.randint(0, m + 50, (m,), generator=g).to(DEV)
    w = (torch.randn(n, k, generator=g) / np.sqrt(k)).to(DEV)
    dy = torch.randn(m, n, generator=g).to(DEV)
    h = torch.tanh(torch.randn(m, k, generator=g)).to(DEV)
    dx = torch.empty(m, k, device=DEV)
    dw = torch.empty(n, k, device=DEV)
    db = torch.empty(n, device=DEV)
    ws = _ws(m, n, k)
    st = lib.stream_ptr(DEV)
    lib.call('bsig_linear_dgrad', dy.data_ptr(), w.data_ptr(), h.data_ptr(), dx.data_ptr(), m, n, k,
             1, SIMT, ws.data_ptr(), ws.numel(), st)
    ref = (dy.double() @ w.double()) * (1 - h.double() ** 2)
    assert _rel(dx, ref) < 1e-5
    lib.call('bsig_linear_wgrad', dy.data_ptr(), xfull.data_ptr(), k, rows.data_ptr(), dw.data_ptr(),
             db.data_ptr(), m, n, k, SIMT, ws.data_ptr(), ws.numel(), st)
    xg = xfull[rows].double()
    assert _rel(dw, dy.double().T @ xg) < 1e-5
    assert _rel(db, dy.double().sum(0)) < 1e-5
    y = torch.empty(m, n, device=DEV)
    b = torch.zeros(n, device=DEV)
    lib.call('bsig_linear_fwd', xfull.data_ptr(), k, rows.data_ptr(), w.data_ptr(), b.data_ptr(),
             y.data_ptr(), m, n, k, 0, SIMT, ws.data_ptr(), ws.numel(), st)
    assert _rel(y, xg @ w.double().T) < 1e-5


@pytest.mark.parametrize('engine', [TC_TF32, TC_X3])
@pytest.mark.parametrize('shape', [(4096, 128, 302), (1000, 270, 128), (257, 64, 100),
                                   (100, 128, 20002), (5000, 128, 128), (384, 270, 128)])
def test_dgrad_wgrad_gather_on_tensor_cores(engine, shape):
    """dgrad (MN-major weight operand, fused tanh'), wgrad (both operands MN-major, minibatch
    gather, split-K for the skinny output) and a gathered forward on the tcgen05 engine,
    including widths that are 2 mod 4 (302, 270, 20002: staged through a padded copy),
    against float64.  Same tolerances as the forward engine tests, with the length of the
    reduction of each GEMM (n for dgrad, the batch for wgrad, k for forward)."""
    lib = _lib()
    m, n, k = shape
    g = torch.Generator('cpu').manual_seed(11 * m + n)
    xfull = torch.randn(m + 50, k, generator=g).to(DEV)
    rows = torch.randint(0, m + 50, (m,), generator=g).to(DEV)
    w = (torch.randn(n, k, generator=g) / np.sqrt(k)).to(DEV)
    dy = torch.randn(m, n, generator=g).to(DEV)
    h = torch.tanh(torch.randn(m, k, generator=g)).to(DEV)
    dx = torch.empty(m, k, device=DEV)
    dw = torch.empty(n, k, device=DEV)
    db = torch.empty(n, device=DEV)
    ws = _ws(m, n, k)
    st = lib.stream_ptr(DEV)
    lib.call('bsig_linear_dgrad', dy.data_ptr(), w.data_ptr(), h.data_ptr(), dx.data_ptr(), m, n, k,
             1, engine, ws.data_ptr(), ws.numel(), st)
    ref = (dy.double() @ w.double()) * (1 - h.double() ** 2)
    assert _rel(dx, ref) < tol_for(engine, n), ('dgrad', _rel(dx, ref))
    lib.call('bsig_linear_wgrad', dy.data_ptr(), xfull.data_ptr(), k, rows.data_ptr(), dw.data_ptr(),
             db.data_ptr(), m, n, k, engine, ws.data_ptr(), ws.numel(), st)
    xg = xfull[rows].double()
    assert _rel(dw, dy.double().T @ xg) < tol_for(engine, m), ('wgrad', _rel(dw, dy.double().T @ xg))
    assert _rel(db, dy.double().sum(0)) < 1e-5
    y = torch.empty(m, n, device=DEV)
    b = torch.randn(n, generator=g).to(DEV)
    lib.call('bsig_linear_fwd', xfull.data_ptr(), k, rows.data_ptr(), w.data_ptr(), b.data_ptr(),
             y.data_ptr(), m, n, k, 1, engine, ws.data_ptr(), ws.numel(), st)
    ref = torch.tanh(xg @ w.double().T + b.double())
    assert _rel(y, ref) < tol_for(engine, k), ('fwd', _rel(y, ref))


@pytest.mark.parametrize('shape', [(100, 128, 302), (100, 24, 30), (16, 16, 1), (37, 130, 65), (128, 7, 9)])
def test_weight_gradient_with_adam_epilogue_is_bit_identical_to_the_two_kernel_form(shape):
    """bsig_linear_wgrad_adam (last backward kernel of a single-GPU update: first-layer weight
    gradient with Adam of the WHOLE flat parameter buffer in its epilogue) against
    bsig_linear_wgrad + bsig_adam_step on the same buffers: same arithmetic, same parameters bit for bit."""
    lib = _lib()
    m, n, k = shape
    g = torch.Generator('cpu').manual_seed(m * 7 + n * 3 + k)
    tail = 1000 + (m % 3)                              # parameters of the other layers
    n_params = n * k + n + tail
    flat = torch.randn(n_params, generator=g).to(DEV)
    grads = torch.randn(n_params, generator=g).to(DEV)
    ea = (0.01 * torch.randn(n_params, generator=g)).to(DEV)
    es = (1e-4 * torch.rand(n_params, generator=g)).to(DEV)
    x = torch.randn(m + 11, k, generator=g).to(DEV)
    rows = torch.randint(0, m + 11, (m,), generator=g).to(DEV)
    dy = torch.randn(m, n, generator=g).to(DEV)
    ws = _ws(m, n, k)
    # reference: separate weight gradient (+ bias gradient) into the flat gradient buffer, then Adam
    p_ref, g_ref, m_ref, v_ref = flat.clone(), grads.clone(), ea.clone(), es.clone()
    lib.call('bsig_linear_wgrad', dy.data_ptr(), x.data_ptr(), k, rows.data_ptr(), g_ref.data_ptr(),
             g_ref.data_ptr() + 4 * n * k, m, n, k, SIMT, ws.data_ptr(), ws.numel(), lib.stream_ptr(DEV))
    # (bsig_adam_step wants 16-byte aligned buffers: clones are)
    lib.call('bsig_adam_step', p_ref.data_ptr(), g_ref.data_ptr(), m_ref.data_ptr(), v_ref.data_ptr(),
             n_params, 5, 1e-3, 0.9, 0.999, 1e-8, 1.0, lib.stream_ptr(DEV))
    p, mm, vv = flat.clone(), ea.clone(), es.clone()
    lib.call('bsig_linear_wgrad_adam', dy.data_ptr(), x.data_ptr(), k, rows.data_ptr(), m, n, k,
             p.data_ptr(), grads.data_ptr(), mm.data_ptr(), vv.data_ptr(), n_params, 5, 1e-3, 0.9,
             0.999, 1e-8, lib.stream_ptr(DEV))
    # (the second moment may differ in its last bit: the compiler contracts v*b2 + (1-b2)*g*g into
    # different fma shapes in the two kernels)
    assert torch.equal(p, p_ref) and torch.equal(mm, m_ref)
    assert torch.allclose(vv, v_ref, rtol=3e-7, atol=0.0)
