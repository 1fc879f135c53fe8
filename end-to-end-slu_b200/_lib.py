"""ctypes binding of the C-ABI shared library (include/slu_b200.h).  No CPU fallback: if the
library is missing or a kernel launch fails this raises -- the CUDA path must fail loudly."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libslu_b200.so")
_lib = None

_P = ctypes.c_void_p
_I = ctypes.c_int
_L = ctypes.c_long
_F = ctypes.c_float
_U = ctypes.c_ulonglong
SIGNATURES = {
    # name: argtypes (all return int = cudaError_t, 0 on success)
    "slu_sinc_filters_fwd": [_P, _P, _P, _P],
    "slu_sinc_filters_bwd": [_P, _P, _P, _P, _P, _P],
    "slu_sincconv_fwd_simt": [_P, _P, _I, _I, _P, _P, _P],
    "slu_sincconv_bwd_simt": [_P, _P, _P, _I, _I, _P, _P],
    "slu_sincconv_fwd_tc": [_P, _P, _I, _I, _P, _P, _P, _P],
    "slu_sincconv_bwd_tc": [_P, _P, _P, _I, _I, _P, _P],
    "slu_sincconv_bwd_jac_tc": [_P, _P, _P, _P, _I, _I, _P, _P, _P],
    "slu_sinc_filters_jac": [_P, _P, _P, _P],
    "slu_set_sinc_persistent": [_I],
    "slu_debug_sinc_trace": [_P],
    "slu_debug_wgrad_mode": [_I],
    "slu_debug_gemm_mode": [_I],
    "slu_debug_wgrad_trace": [_P],
    "slu_gru_fwd_simt": [_P, _P, _P, _P, _F, _U, _P, _I, _I, _I, _P, _P, _P, _P],
    "slu_gru_bwd_simt": [_P, _P, _F, _U, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P],
    "slu_gru_fwd_tc": [_P, _P, _P, _P, _F, _U, _P, _I, _I, _I, _P, _P, _P, _P],
    "slu_gru_bwd_tc": [_P, _P, _F, _U, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P],
    "slu_bigru_bwd_tc": [_P, _P, _F, _U, _P, _P, _P, _P, _P, _I, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _I, _P],
    "slu_set_gru_precision": [_I],
    "slu_gru_rows_per_cta": [_I],
    "slu_debug_gru_phase_clocks": [_P],
    "slu_intent_head_fwd": [_P, _P, _P, _P, _I, _I, _I, _P, _I, _P, _P, _P, _P, _P, _P, _P],
    "slu_intent_head_bwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _I, _P, _P, _P, _P],
    "slu_h2d_async": [_P, _P, ctypes.c_size_t, _P, _I],
    "slu_h2d_ready": [_P],
    "slu_h2d_pending": [_P],
    "slu_h2d_wait": [],
    "slu_stream_fork": [_P, _I, _P],
    "slu_stream_join": [_P, _I],
    "slu_dropout_mask": [_P, _L, _F, ctypes.c_ulonglong, _P],
    "slu_dropout_mask_gru": [_P, _I, _I, _F, ctypes.c_ulonglong, _P],
    "slu_leaky_bwd_bias": [_P, _P, _F, _P, _P, _L, _I, _P],
    "slu_gemm_tc": [_P, _L, _P, _P, _P, _L, _I, _I, _I, _I, _I, _I, _I, _F, _P],
    "slu_presplit_bf16": [_P, _L, _L, _L, _I, _I, _I, _P, _P],
    "slu_presplit_multi": [_P, _I, _P],
    "slu_wgrad2_tc": [_P, _L, _I, _P, _L, _I, _P, _L, _I, _I, _I, _I, _I, _P, _L, _L, _L, _P],
    "slu_wgrad_tc": [_P, _L, _I, _P, _L, _I, _I, _I, _I, _I, _P, _L, _L, _L, _P],
    "slu_skinny_gemm": [_P, _L, _P, _L, _L, _P, _P, _L, _I, _I, _I, _P],
    "slu_attn_step_fwd": [_P, _L, _P, _P, _I, _I, _I, _I, _F, _P, _P, _P],
    "slu_attn_step_bwd": [_P, _P, _P, _L, _P, _P, _I, _I, _I, _I, _F, _P, _L, _P, _P, _P],
    "slu_grucell_fwd": [_P, _L, _P, _L, _P, _L, _P, _P, _I, _I, _F, _U, _P, _I, _P, _P, _P, _P],
    "slu_grucell_bwd": [_P, _P, _P, _P, _P, _P, _I, _I, _F, _U, _P, _I, _P, _L, _P, _L, _P, _P],
    "slu_seed_advance": [_P, _P],
    "slu_ce_count": [_P, _L, _P, _P],
    "slu_ce_rows": [_P, _L, _I, _P, _L, _P, _I, _P, _P, _P],
    "slu_ce_finish": [_P, _P, _L, _P, _P, _P],
    "slu_colsum_acc": [_P, _L, _L, _I, _P, _P],
    "slu_scale": [_P, _P, _L, _P, _P],
    "slu_adam_multi": [_P, _I, ctypes.c_double, ctypes.c_double, _F, _F, _P],
    "slu_f64_hilo_split": [_P, _P, _P, _I, _P],
    "slu_f64_hilo_merge": [_P, _P, _P, _I, _P],
    "slu_tc_selftest": [_P, _P, _P, _I, _I, _P],
    "slu_tc_selftest_ts": [_P, _P, _P, _I, _I, _P],
}


def load():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError("slu_b200: %s not built -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(nvcc, sm_100a).  There is no fallback path." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the symbol is missing: loud by design
            fn.argtypes = argtypes
            fn.restype = _I
        _lib = lib
        lib.slu_set_gru_precision(1 if os.environ.get("SLU_GRU_PRECISION", "bf16x3") == "fp16" else 0)
    return _lib


def ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "slu_b200 kernels take contiguous CUDA tensors"
    return t.data_ptr()


def stream():
    """Raw handle of torch's current stream on the current device (what every launch of this library is queued on)."""
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


ERR_TOO_LARGE = 100001      # SLU_ERR_TOO_LARGE in include/slu_b200.h
stats = {"calls": 0}        # number of C-ABI kernel launches issued by this process
_prof = None                # name -> [(start_event, end_event)] while profiling
_prof_detail = None         # [(name, small-int args, start, end)] per launch, when asked for
_fn = {}
_HOST_ONLY = ("slu_set_sinc_persistent", "slu_debug_sinc_trace", "slu_debug_wgrad_mode", "slu_debug_gemm_mode", "slu_debug_wgrad_trace", "slu_gru_rows_per_cta", "slu_h2d_async", "slu_h2d_ready", "slu_h2d_pending", "slu_h2d_wait", "slu_stream_fork", "slu_stream_join", "slu_set_gru_precision", "slu_debug_gru_phase_clocks")


def call(name, *args):
    fn = _fn.get(name)
    if fn is None:
        fn = _fn[name] = getattr(load(), name)
    if name not in _HOST_ONLY:
        stats["calls"] += 1
    if _prof is not None and name not in _HOST_ONLY:
        st = torch.cuda.ExternalStream(args[-1]) if args[-1] else torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        err = fn(*args)
        e1.record(st)
        _prof.setdefault(name, []).append((e0, e1))
        if _prof_detail is not None:
            _prof_detail.append((name, tuple(a for a in args if isinstance(a, int) and 0 <= a < (1 << 24)), e0, e1))
    else:
        err = fn(*args)
    if err != 0:
        if err == ERR_TOO_LARGE:
            raise RuntimeError("slu_b200: %s: a size exceeds the kernel's 32-bit index range (the GRU kernels take B*T < 2^21 "
                               "frames per launch) -- split the batch" % name)
        raise RuntimeError("slu_b200: %s failed with cudaError %d" % (name, err))


def fork(n):
    """-> (main stream handle, [n side-stream handles]) with the side streams waiting on the current stream."""
    main = stream()
    side = (ctypes.c_void_p * n)()
    call("slu_stream_fork", main, n, side)
    return main, list(side)


def join(main, n):
    call("slu_stream_join", main, n)


def profile_begin(detail=False):
    """Start recording CUDA events around every launch (on the stream it is queued on)."""
    global _prof, _prof_detail
    _prof = {}
    _prof_detail = [] if detail else None


def profile_detail():
    """-> [(entry point, (small integer arguments: sizes / strides), device ms)] in launch order; call before profile_end."""
    torch.cuda.synchronize()
    return [(n, a, e0.elapsed_time(e1)) for n, a, e0, e1 in (_prof_detail or [])]


def profile_end():
    """-> {entry point: [device ms per launch]}"""
    global _prof
    torch.cuda.synchronize()
    out = {k: [a.elapsed_time(b) for a, b in v] for k, v in _prof.items()}
    _prof = None
    return out
