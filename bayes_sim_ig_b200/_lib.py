"""ctypes binding of libbsig_b200.so (the C ABI declared in include/bsig.h).

There is no CPU fallback anywhere in this package: if the shared library is
missing the import fails loudly, and every kernel entry point refuses tensors
that are not fp32/contiguous/on a CUDA device.
"""
import ctypes
import os

import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, 'libbsig_b200.so')

_c_ptr = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int
_f32 = ctypes.c_float
_u64 = ctypes.c_uint64



class TpDesc(ctypes.Structure):
    """bsig_tp_desc of include/bsig.h (persistent training kernel)."""
    _fields_ = [('n_layers', ctypes.c_int32), ('in_dim', ctypes.c_int32 * 3),
                ('out_dim', ctypes.c_int32 * 3), ('batch', ctypes.c_int32),
                ('p', ctypes.c_int32), ('k', ctypes.c_int32), ('full_cov', ctypes.c_int32),
                ('w_off', ctypes.c_int64 * 3), ('b_off', ctypes.c_int64 * 3),
                ('x', _c_ptr), ('ldx', _i64), ('y', _c_ptr), ('idx', _c_ptr), ('noise', _c_ptr),
                ('params', _c_ptr), ('exp_avg', _c_ptr), ('exp_avg_sq', _c_ptr),
                ('scratch', _c_ptr), ('scratch_floats', _i64), ('loss_buf', _c_ptr),
                ('loss_slot', _c_ptr), ('adam_coef', _c_ptr), ('flag', _c_ptr),
                ('beta1', _f32), ('beta2', _f32), ('eps', _f32), ('prof', _c_ptr)]


# name -> (restype, argtypes); must list every symbol of include/bsig.h
SIGNATURES = {
    'bsig_train_persistent_query': (_int, [ctypes.POINTER(TpDesc), ctypes.POINTER(_i64),
                                           ctypes.POINTER(_i64)]),
    'bsig_train_persistent': (_int, [ctypes.POINTER(TpDesc), _i64, _i64, _c_ptr]),
    'bsig_last_error': (ctypes.c_char_p, []),
    'bsig_version': (_int, []),
    'bsig_launch_count': (_i64, []),
    'bsig_set_pdl': (_int, [_int]),
    'bsig_device_info': (_int, [ctypes.POINTER(_int)] * 3),
    'bsig_summary_start': (_int, [_c_ptr, _c_ptr, _c_ptr] + [_i64] * 6 + [_c_ptr]),
    'bsig_summary_start_tm': (_int, [_c_ptr, _c_ptr, _c_ptr] + [_i64] * 6 + [_c_ptr]),
    'bsig_summary_crosscorr': (_int, [_c_ptr, _c_ptr, _c_ptr] + [_i64] * 6 + [_int, _c_ptr, _c_ptr]),
    'bsig_summary_crosscorr_tm': (_int, [_c_ptr, _c_ptr, _c_ptr] + [_i64] * 6 + [_int, _c_ptr, _c_ptr]),
    'bsig_corr_factors': (_int, [_c_ptr, _c_ptr, _c_ptr] + [_i64] * 7 + [_int, _int, _c_ptr, _c_ptr]),
    'bsig_corr_linear_applicable': (_int, [_i64] * 5),
    'bsig_corr_linear_ws_bytes': (_i64, [_i64] * 4),
    'bsig_corr_linear_fwd': (_int, [_c_ptr, _i64, _c_ptr, _i64, _i64, _c_ptr, _c_ptr, _c_ptr, _i64,
                                    _i64, _int, _c_ptr, _i64, _c_ptr]),
    'bsig_corr_rff_features': (_int, [_c_ptr, _i64, _c_ptr, _i64, _i64, _c_ptr, _c_ptr, _i64, _i64,
                                      _f32, _c_ptr, _i64, _c_ptr]),
    'bsig_corr_linear_wgrad': (_int, [_c_ptr, _c_ptr, _i64, _c_ptr, _i64, _i64, _i64, _i64,
                                      _c_ptr, _c_ptr, _c_ptr, _c_ptr, _i64, _f32, _f32, _f32,
                                      _f32, _f32, _c_ptr]),
    'bsig_signature_fwd': (_int, [_c_ptr, _c_ptr, _c_ptr] + [_i64] * 6 + [_int, _c_ptr]),
    'bsig_signature_bwd': (_int, [_c_ptr] * 5 + [_i64] * 6 + [_int, _c_ptr]),
    'bsig_linear_ws_bytes': (_i64, [_i64] * 3),
    'bsig_linear_fwd': (_int, [_c_ptr, _i64, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _i64, _i64, _i64,
                               _int, _int, _c_ptr, _i64, _c_ptr]),
    'bsig_linear_dgrad': (_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _i64, _i64, _i64, _int, _int,
                                 _c_ptr, _i64, _c_ptr]),
    'bsig_linear_wgrad': (_int, [_c_ptr, _c_ptr, _i64, _c_ptr, _c_ptr, _c_ptr, _i64, _i64, _i64,
                                 _int, _c_ptr, _i64, _c_ptr]),
    'bsig_linear_wgrad_adam': (_int, [_c_ptr, _c_ptr, _i64, _c_ptr, _i64, _i64, _i64, _c_ptr, _c_ptr,
                                      _c_ptr, _c_ptr, _i64, _i64, _f32, _f32, _f32, _f32, _c_ptr]),
    'bsig_linear_colsum': (_int, [_c_ptr, _c_ptr, _i64, _i64, _c_ptr]),
    'bsig_tanh_bwd': (_int, [_c_ptr, _c_ptr, _c_ptr, _i64, _c_ptr]),
    'bsig_rff_features': (_int, [_c_ptr, _i64, _c_ptr, _c_ptr, _c_ptr, _i64, _i64, _i64, _f32,
                                 _int, _c_ptr, _i64, _c_ptr]),
    'bsig_mdn_ws_bytes': (_i64, [_i64]),
    'bsig_mdn_head_fwd': (_int, [_c_ptr] * 4 + [_i64] * 3 + [_int, _c_ptr, _i64, _c_ptr, _c_ptr]),
    'bsig_mdn_head_bwd': (_int, [_c_ptr] * 8 + [_i64] * 3 + [_int, _c_ptr, _i64, _c_ptr]),
    'bsig_mog_nll_fwd': (_int, [_c_ptr, _c_ptr, _i64, _c_ptr, _i64, _c_ptr, _i64, _c_ptr, _c_ptr,
                                _c_ptr, _i64, _i64, _i64, _c_ptr, _i64, _c_ptr, _c_ptr]),
    'bsig_mog_nll_bwd': (_int, [_c_ptr, _c_ptr, _i64, _c_ptr, _i64, _c_ptr, _i64, _c_ptr, _c_ptr,
                                _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _i64, _i64, _i64,
                                _c_ptr, _i64, _c_ptr]),
    'bsig_mdn_nll_fused': (_int, [_c_ptr] * 6 + [_i64] * 3 + [_int, _c_ptr, _i64, _c_ptr, _c_ptr]),
    'bsig_adam_step': (_int, [_c_ptr] * 4 + [_i64, _i64, _f32, _f32, _f32, _f32, _f32, _c_ptr]),
    'bsig_p2p_alloc': (_int, [ctypes.POINTER(_c_ptr), _i64, ctypes.c_char_p]),
    'bsig_p2p_open': (_int, [ctypes.c_char_p, ctypes.POINTER(_c_ptr)]),
    'bsig_p2p_close': (_int, [_c_ptr]),
    'bsig_p2p_read': (_int, [_c_ptr, _c_ptr, _i64]),
    'bsig_p2p_free': (_int, [_c_ptr]),
    'bsig_p2p_set_timeout_ms': (_int, [_i64]),
    'bsig_p2p_preload': (_int, []),
    'bsig_adam_allreduce_step': (_int, [_c_ptr, ctypes.POINTER(_c_ptr), ctypes.POINTER(_c_ptr),
                                        _c_ptr, _int, _int, _c_ptr, _c_ptr, _i64, _i64,
                                        _f32, _f32, _f32, _f32, _c_ptr]),
    'bsig_wgrad3_adam_step': (_int, [_c_ptr, _c_ptr, _i64, _c_ptr, _i64, _i64, _i64, _i64,
                                     _c_ptr, _c_ptr, _i64, _i64, _i64, _i64,
                                     _c_ptr, _c_ptr, _i64, _i64, _i64, _i64,
                                     _c_ptr, _c_ptr, _c_ptr, _i64, _i64, _f32, _f32, _f32, _f32,
                                     _c_ptr]),
    'bsig_gather_rows': (_int, [_c_ptr, _i64, _c_ptr, _c_ptr, _i64, _i64, _c_ptr]),
    'bsig_finite_flag': (_int, [_c_ptr, _i64, _c_ptr, _c_ptr]),
    'bsig_normalize_rows': (_int, [_c_ptr] * 4 + [_i64, _i64, _c_ptr]),
    'bsig_mog_denorm': (_int, [_c_ptr, _c_ptr, _i64, _c_ptr, _i64, _c_ptr, _i64, _c_ptr, _c_ptr,
                               _c_ptr, _c_ptr, _c_ptr, _i64, _i64, _i64, _c_ptr]),
    'bsig_mog_sample': (_int, [_c_ptr, _int] + [_c_ptr] * 7 + [_i64] * 3 + [_c_ptr]),
    'bsig_mog_sample_philox': (_int, [_c_ptr] * 5 + [_u64, _i64, _i64, _i64, _c_ptr]),
    'bsig_mog_sample_envs': (_int, [_c_ptr, _int] + [_c_ptr] * 8 + [_i64] * 3 + [_c_ptr]),
    'bsig_mog_sample_envs_philox': (_int, [_c_ptr] * 7 + [_u64, _i64, _i64, _i64, _c_ptr]),
    'bsig_mog_marginal_grid': (_int, [_c_ptr] * 5 + [_i64] * 3 + [_int, _c_ptr]),
    'bsig_mog_logpdf': (_int, [_c_ptr, _int] + [_c_ptr] * 6 + [_i64] * 3 + [_int, _c_ptr]),
}

_lib = None


class BsigError(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes handle; raises if the .so is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            'libbsig_b200.so is not built (%s). Run `python -m bayes_sim_ig_b200.build` '
            '(needs nvcc); this package has no CPU fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)     # AttributeError if the symbol is missing
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error():
    return load().bsig_last_error().decode('utf-8', 'replace')


def call(name, *args):
    """Invoke an int-returning entry point and raise on a non-zero status."""
    rc = getattr(load(), name)(*args)
    if rc != 0:
        raise BsigError('%s failed (%d): %s' % (name, rc, last_error()))


import contextlib


@contextlib.contextmanager
def nvtx_range(name):
    """NVTX range around a stage of the path (summarize / train / predict / sample): shows up
    in nsys / ncu --nvtx timelines; a no-op cost of two driver calls otherwise."""
    torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t, dtype=torch.float32, allow_none=False):
    """Device pointer of a tensor after checking what the kernels assume."""
    if t is None:
        if allow_none:
            return None
        raise BsigError('missing tensor argument')
    if not t.is_cuda:
        raise BsigError('bsig kernels need CUDA tensors (got %s); there is no CPU fallback'
                        % t.device)
    if t.dtype != dtype:
        raise BsigError('expected dtype %s, got %s' % (dtype, t.dtype))
    if not t.is_contiguous():
        raise BsigError('expected a contiguous tensor')
    return t.data_ptr()


def require_cuda(device):
    dev = torch.device(device)
    if dev.type != 'cuda':
        raise BsigError("bayes_sim_ig_b200 runs on CUDA devices only (device=%r): the hot path "
                        "is hand-written sm_100a kernels with no CPU fallback" % (device,))
    if not torch.cuda.is_available():
        raise BsigError('CUDA is not available in this process')
    return dev
