"""Autograd wrappers over the C-ABI kernels (CUDA tensors only).

Layouts are the library's own (time-major NLC everywhere):
  waveform [B,T] -> sinc frames [B,L1,80] -> conv blocks [B,L1,60] -> GRU stacks [B,T_l,256].
"""
import ctypes
import os

import torch

from . import _lib, grads

H = 128
# Which persistent-GRU kernel family runs the recurrence: "tc" = tcgen05 (weights stationary in TMEM),
# "simt" = fp32 CUDA-core variant.  Both are sm_100a kernels of this library with identical contracts.
GRU_IMPL = os.environ.get("SLU_GRU_IMPL", "tc")
# SincConv: "tc" = tcgen05 6-tap framing GEMM, "simt" = fp32 CUDA-core kernel.
SINC_IMPL = os.environ.get("SLU_SINC_IMPL", "tc")


# Weight-gradient launches of one layer are independent of each other and of the input-gradient GEMM: they go to side
# streams (forked after the producer kernel, joined before the autograd node returns), so the small grids share the GPU.
OVERLAP = os.environ.get("SLU_OVERLAP", "1") != "0"
# SLU_FUSED_BWD=1: one C-ABI call per GRU layer backward (slu_bigru_bwd_tc) instead of driving its launches one by one from
# Python -- ~0.3 ms less host time per step.  Off by default: on a single GPU the step is GPU-bound and the A/B on one box
# measured 2.98 ms (one call) vs 2.94 ms (launch by launch); it is meant for host-bound set-ups.
FUSED_BWD = os.environ.get("SLU_FUSED_BWD", "0") != "0"


class _Fork:
    """Side streams of the library (slu_stream_fork / slu_stream_join): `stream(i)` is where independent launch i goes.
    All tensors involved are allocated on the current stream BEFORE the fork and stay referenced until after the join,
    so the caching allocator never sees cross-stream reuse."""

    def __init__(self, n):
        self.n = n if OVERLAP else 0
        if self.n:
            self.main, self.side = _lib.fork(self.n)
        else:
            self.main, self.side = _lib.stream(), []

    def stream(self, i):
        return self.side[i % self.n] if self.n else self.main

    def join(self):
        if self.n:
            _lib.join(self.main, self.n)


def _f32(t):
    return t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()


def _reserve(ctx, numel, f64=False):
    """Forward side of the gradient arena (grads.py): remember the arena of this forward pass and a slot of `numel` values."""
    arena = grads.current()
    ctx.arena = arena
    return arena.reserve(numel, f64) if arena is not None else None


def _zeros(ctx, key, shape, f64=False):
    """Backward side: the zero-filled slot as a tensor of `shape` (a private torch.zeros when there is no arena / slot)."""
    v = ctx.arena.view(key, shape, f64) if getattr(ctx, "arena", None) is not None else None
    if v is None:
        v = torch.zeros(shape, device=ctx.arena.device if getattr(ctx, "arena", None) is not None else torch.device("cuda"),
                        dtype=torch.float64 if f64 else torch.float32)
    return v


def _eptr(t, off=0):
    assert t.is_cuda and t.dtype == torch.float32
    return t.data_ptr() + 4 * off


def gemm_tc(A, lda, w_img, M, N, K, out, bias=None, taps=1, tap_pad=0, T=0, act=0, slope=0.0):
    """out[M,N] = sum_tap A[(m+tap-tap_pad)*lda + k] * W(n,tap,k) (+bias)(LeakyReLU) -- slu_gemm_tc, see include/slu_b200.h."""
    _lib.call("slu_gemm_tc", _eptr(A), lda, w_img.data_ptr(), None if bias is None else _eptr(bias), _eptr(out), N, M, N, K, taps,
              tap_pad, T, act, float(slope), _lib.stream())
    return out


def presplit(W, w_off, sn, sk, stap, taps, N, K):
    """Weights -> bf16 hi/lo operand image for gemm_tc's B side (slu_presplit_bf16)."""
    Kp = (K + 31) // 32 * 32
    img = torch.empty(2 * taps * N * Kp, device=W.device, dtype=torch.bfloat16)
    _lib.call("slu_presplit_bf16", _eptr(W, w_off), sn, sk, stap, taps, N, K, img.data_ptr(), _lib.stream())
    return img


class _Job(ctypes.Structure):
    _fields_ = [("W", ctypes.c_void_p), ("sn", ctypes.c_long), ("sk", ctypes.c_long), ("stap", ctypes.c_long),
                ("taps", ctypes.c_int), ("N", ctypes.c_int), ("K", ctypes.c_int), ("pad", ctypes.c_int), ("img", ctypes.c_void_p)]


# weight-operand forms (arguments of slu_presplit_bf16) of the layers' GEMMs
def _form_nt(w):                     # linear_nt: w [N, K] row-major
    N, K = w.shape
    return (0, K, 1, 0, 1, N, K)


def _form_nn(w):                     # matmul_nn: w [K, N] row-major, read transposed
    K, N = w.shape
    return (0, 1, N, 0, 1, N, K)


def _form_conv_fwd(w):               # Conv1d weight [Cout, Cin, k] as k taps of [Cout, Cin]
    Cout, Cin, k = w.shape
    return (0, Cin * k, k, 1, k, Cout, Cin)


def _form_conv_dx(w):                # input gradient: taps walked backwards, [Cin, Cout] per tap
    Cout, Cin, k = w.shape
    return (k - 1, k, Cin * k, -1, k, Cin, Cout)


def presplit_many(items):
    """items: [(weight tensor, form tuple)] -> list of operand images, made by ONE launch per 16 items (slu_presplit_multi)."""
    imgs = []
    for base in range(0, len(items), 16):
        chunk = items[base:base + 16]
        jobs = (_Job * len(chunk))()
        for i, (w, (w_off, sn, sk, stap, taps, N, K)) in enumerate(chunk):
            Kp = (K + 31) // 32 * 32
            img = torch.empty(2 * taps * N * Kp, device=w.device, dtype=torch.bfloat16)
            jobs[i] = _Job(_eptr(w, w_off), sn, sk, stap, taps, N, K, 0, img.data_ptr())
            imgs.append(img)
        _lib.call("slu_presplit_multi", jobs, len(chunk), _lib.stream())
    return imgs


def wgrad_tc(G, g_off, ldg, M, X, x_off, ldx, N, B, T, out, o_off, s_m, s_n=1, s_tap=0, taps=1, shift0=0, stream=None):
    """out[...] += G^T . X over frames (slu_wgrad_tc); `out` must be pre-zeroed for a plain gradient."""
    _lib.call("slu_wgrad_tc", _eptr(G, g_off), ldg, M, _eptr(X, x_off), ldx, N, B, T, taps, shift0, _eptr(out, o_off), s_m, s_n,
              s_tap, _lib.stream() if stream is None else stream)
    return out


def wgrad2_tc(G0, g0_off, ldg0, m_split, G1, g1_off, ldg1, M, X, x_off, ldx, N, B, T, out, o_off, s_m, shift0=0, stream=None):
    """Like wgrad_tc with the rows of G taken from two tensors (rows < m_split from G0, the rest from G1): slu_wgrad2_tc."""
    _lib.call("slu_wgrad2_tc", _eptr(G0, g0_off), ldg0, m_split, _eptr(G1, g1_off), ldg1, M, _eptr(X, x_off), ldx, N, B, T, 1, shift0,
              _eptr(out, o_off), s_m, 1, 0, _lib.stream() if stream is None else stream)
    return out


def linear_nt(x2, w, bias=None, img=None):
    """x2 [M,K] @ w[N,K]^T + bias -> [M,N]  (img: w's operand image if the caller already made it)."""
    M, K = x2.shape
    N = w.shape[0]
    out = torch.empty(M, N, device=x2.device, dtype=torch.float32)
    return gemm_tc(x2, K, presplit(w, *_form_nt(w)) if img is None else img, M, N, K, out, bias=bias)


def matmul_nn(a2, w, img=None):
    """a2 [M,K] @ w[K,N] -> [M,N]  (w row-major, i.e. the 'transposed weight' operand of an input gradient)."""
    M, K = a2.shape
    N = w.shape[1]
    out = torch.empty(M, N, device=a2.device, dtype=torch.float32)
    return gemm_tc(a2, K, presplit(w, *_form_nn(w)) if img is None else img, M, N, K, out)


def matmul_tn(g2, x2):
    """g2[R,M]^T @ x2[R,N] -> [M,N]: weight gradient, reduction over the R frames."""
    R, M = g2.shape
    N = x2.shape[1]
    out = torch.zeros(M, N, device=g2.device, dtype=torch.float32)
    return wgrad_tc(g2, 0, M, M, x2, 0, N, N, 1, R, out, 0, N)


class ConvBlock(torch.autograd.Function):
    """Conv1d(k odd, pad k//2) + bias + LeakyReLU on NLC activations as an accumulating tap-GEMM (models.py:200-220)."""

    @staticmethod
    def forward(ctx, x, weight, bias, slope, imgs=None):
        x = _f32(x)
        B, T, Cin = x.shape
        Cout, _, k = weight.shape
        w = weight.detach().contiguous()
        out = torch.empty(B, T, Cout, device=x.device, dtype=torch.float32)
        img_fwd, img_dx = imgs if imgs is not None else (presplit(w, *_form_conv_fwd(w)), None)
        gemm_tc(x, Cin, img_fwd, B * T, Cout, Cin, out, bias=bias.detach(), taps=k, tap_pad=k // 2, T=T, act=1, slope=slope)
        ctx.save_for_backward(x, w, out)
        ctx.slope = slope
        ctx.img_dx = img_dx
        ctx.slot = _reserve(ctx, Cout + Cout * Cin * k) if (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]) else None
        return out

    @staticmethod
    def backward(ctx, gy):
        x, w, out = ctx.saved_tensors
        B, T, Cin = x.shape
        Cout, _, k = w.shape
        gy = _f32(gy)
        if Cout % 4 != 0:
            raise NotImplementedError("slu_b200 CUDA path: conv blocks need a channel count that is a multiple of 4")
        # LeakyReLU backward + bias gradient in one pass
        dpre = torch.empty_like(out)
        zb = _zeros(ctx, ctx.slot, (Cout + Cout * Cin * k,))                               # db | dW (arena slot)
        db = zb[:Cout]
        _lib.call("slu_leaky_bwd_bias", _lib.ptr(out), _lib.ptr(gy), float(ctx.slope), _lib.ptr(dpre), _lib.ptr(db), B * T, Cout,
                  _lib.stream())
        dx = dw = fork = None
        if ctx.needs_input_grad[1]:      # all k taps in one launch, written straight into the [Cout][Cin][k] weight layout
            dw = zb[Cout:].view(Cout, Cin, k)
            fork = _Fork(1)
            wgrad_tc(dpre, 0, Cout, Cout, x, 0, Cin, Cin, B, T, dw, 0, Cin * k, k, 1, taps=k, shift0=-(k // 2), stream=fork.stream(0))
        if ctx.needs_input_grad[0]:
            dx = torch.empty(B, T, Cin, device=x.device, dtype=torch.float32)
            # dX[b,t,ci] = sum_d sum_co dpre[b,t-(d-k//2),co] W[co,ci,d]; tap' = k-1-d walks the kernel backwards
            img_dx = ctx.img_dx if ctx.img_dx is not None else presplit(w, *_form_conv_dx(w))
            gemm_tc(dpre, Cout, img_dx, B * T, Cin, Cout, dx, taps=k, tap_pad=k // 2, T=T)
        if not ctx.needs_input_grad[2]:
            db = None
        if fork is not None:
            fork.join()
        return dx, dw, db, None, None


def conv_block(x, weight, bias, slope, imgs=None):
    return ConvBlock.apply(x, weight, bias, slope, imgs)


def set_gru_precision(mode):
    """Operand format of the tcgen05 recurrence: 'bf16x3' (default, fp32-class bf16 hi/lo split; hi/lo rows stacked along N
    on the small-batch tiles), 'fp16' (single pass), 'bf16x3-separate' (three separate passes everywhere, for A/B checks)."""
    _lib.load().slu_set_gru_precision({"bf16x3": 0, "fp16": 1, "bf16x3-separate": 2}[mode])


class SincFrontend(torch.autograd.Function):
    """SincLayer conv (80 filters, 401 taps, stride 80, pad 200) + Abs + MaxPool1d(2, ceil).
    Reference: models.py:77-110, 163-168, 205.  Output is NLC [B, L1, 80]."""

    @staticmethod
    def forward(ctx, x, filt_b1, filt_band):
        x = _f32(x)
        B, T = x.shape
        L0 = (T - 1) // 80 + 1
        L1 = (L0 + 1) // 2
        b1 = filt_b1.detach().double().contiguous()
        band = filt_band.detach().double().contiguous()
        W = torch.empty(80, 401, device=x.device, dtype=torch.float32)
        _lib.call("slu_sinc_filters_fwd", _lib.ptr(b1), _lib.ptr(band), _lib.ptr(W), _lib.stream())
        out = torch.empty(B, L1, 80, device=x.device, dtype=torch.float32)
        need = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        route = torch.empty(B, L1, 80, device=x.device, dtype=torch.uint8) if need else None
        if SINC_IMPL == "tc":
            img = torch.empty(2 * 6 * 80 * 96, device=x.device, dtype=torch.bfloat16)
            _lib.call("slu_sincconv_fwd_tc", _lib.ptr(x), _lib.ptr(W), B, T, _lib.ptr(out), _lib.ptr(route), img.data_ptr(),
                      _lib.stream())
        else:
            _lib.call("slu_sincconv_fwd_simt", _lib.ptr(x), _lib.ptr(W), B, T, _lib.ptr(out), _lib.ptr(route), _lib.stream())
        if need:
            ctx.save_for_backward(x, b1, band, route)
            ctx.slot_f64 = _reserve(ctx, 160, f64=True)            # d_b1 | d_band (fp64, like the parameters)
        return out

    @staticmethod
    def backward(ctx, gy):
        x, b1, band, route = ctx.saved_tensors
        B, T = x.shape
        gy = _f32(gy)
        d = _zeros(ctx, ctx.slot_f64, (160,), f64=True)
        d_b1, d_band = d[:80], d[80:]
        if SINC_IMPL == "tc":
            # cut-off gradients = two more convolutions of the waveform (Jacobian banks) dotted with the routed gradient: no dW
            J = torch.empty(2, 80, 401, device=x.device, dtype=torch.float32)
            img = torch.empty(2 * 6 * 160 * 96, device=x.device, dtype=torch.bfloat16)
            _lib.call("slu_sinc_filters_jac", _lib.ptr(b1), _lib.ptr(band), _lib.ptr(J), _lib.stream())
            _lib.call("slu_sincconv_bwd_jac_tc", _lib.ptr(x), _lib.ptr(gy), _lib.ptr(route), _lib.ptr(J), B, T, d.data_ptr(),
                      img.data_ptr(), _lib.stream())
        else:
            dW = torch.empty(80, 401, device=x.device, dtype=torch.float32)
            _lib.call("slu_sincconv_bwd_simt", _lib.ptr(x), _lib.ptr(gy), _lib.ptr(route), B, T, _lib.ptr(dW), _lib.stream())
            _lib.call("slu_sinc_filters_bwd", _lib.ptr(b1), _lib.ptr(band), _lib.ptr(dW), d_b1.data_ptr(), d_band.data_ptr(),
                      _lib.stream())
        return None, d_b1, d_band


def sinc_filters(filt_b1, filt_band):
    """W[80,401] fp32 on the device (models.py:82-106), for tests / inspection."""
    b1 = filt_b1.detach().double().contiguous()
    band = filt_band.detach().double().contiguous()
    W = torch.empty(80, 401, device=b1.device, dtype=torch.float32)
    _lib.call("slu_sinc_filters_fwd", _lib.ptr(b1), _lib.ptr(band), _lib.ptr(W), _lib.stream())
    return W


class BiGRU(torch.autograd.Function):
    """Bidirectional single-layer GRU (H=128, h0=0) + Dropout + Downsample(avg 2 | none 1).
    Reference: nn.GRU at models.py:232/262/686, RNNSelect :138-149, Dropout :246, Downsample :26-46.
    x [B,T,I] -> [B, ceil(T/ds), 256].  `mask`: None (eval / p = 0), an explicit keep-mask tensor [B,T,256] (already scaled by
    1/(1-p)), or a (p, seed[, seed_word]) tuple = the kernels generate the canonical Philox mask in registers, forward and
    backward (seed_word: an int64 device tensor XOR-ed into the seed at run time, for CUDA-graph replays)."""

    @staticmethod
    def forward(ctx, x, w_ih, w_hh, b_ih, b_hh, w_ih_r, w_hh_r, b_ih_r, b_hh_r, mask, ds, packed=None, before_recurrence=None, imgs=None):
        x = _f32(x)
        B, T, I = x.shape
        dev = x.device
        if packed is not None:      # both directions already adjacent in memory (packed_params): no gather launches
            w_ih_cat, b_ih_cat, w_hh_cat, b_hh_cat = packed
        else:
            w_ih_cat = torch.cat([w_ih, w_ih_r], 0).detach()                   # [768, I]
            b_ih_cat = torch.cat([b_ih, b_ih_r], 0).detach()
            w_hh_cat = torch.stack([w_hh, w_hh_r], 0).detach().contiguous()   # [2,384,128]
            b_hh_cat = torch.stack([b_hh, b_hh_r], 0).detach().contiguous()
        img_nt, img_nn = imgs if imgs is not None else (None, None)
        gx = linear_nt(x.view(B * T, I), w_ih_cat, b_ih_cat, img_nt)            # x-projection, both directions
        T2 = (T + ds - 1) // ds
        drop_p, drop_seed, seed_dev = (float(mask[0]), int(mask[1]), mask[2] if len(mask) > 2 else None) if isinstance(mask, tuple) \
            else (0.0, 0, None)
        if isinstance(mask, tuple):
            mask = None
        y_full = torch.empty(B, T, 256, device=dev, dtype=torch.float32)
        y_out = torch.empty(B, T2, 256, device=dev, dtype=torch.float32) if (ds != 1 or mask is not None or drop_p > 0.0) else y_full
        need = any(ctx.needs_input_grad[:9])
        stash = torch.empty(B, T, 1024, device=dev, dtype=torch.float32) if need else None
        if before_recurrence is not None:       # e.g. join the side stream that wrote the dropout masks (the x-projection is queued)
            before_recurrence()
        _lib.call("slu_gru_fwd_" + GRU_IMPL, _lib.ptr(gx), _lib.ptr(w_hh_cat), _lib.ptr(b_hh_cat), _lib.ptr(mask), drop_p, drop_seed,
                  _lib.ptr(seed_dev), B, T, ds, _lib.ptr(y_full), _lib.ptr(y_out), _lib.ptr(stash), _lib.stream())
        if need:
            ctx.save_for_backward(x, w_ih_cat, w_hh_cat, y_full, stash, mask)
            ctx.ds = ds
            ctx.drop = (drop_p, drop_seed, seed_dev)
            ctx.img_nn = img_nn
            # the packed views alias the Parameters' storage without sharing their version counters: remember the versions so
            # that an in-place update between this forward and its backward is detected (stock autograd would raise too)
            ctx.param_versions = [(q, q._version) for q in (w_ih, w_hh, b_ih, b_hh, w_ih_r, w_hh_r, b_ih_r, b_hh_r)]
            # gradients in the layout of the packed parameter buffer: dW_ih [768,I] | dW_hh [2,384,H] | db_ih [768] | db_hh [768]
            ctx.slot = _reserve(ctx, 768 * I + 2 * 384 * H + 2 * 768) if any(ctx.needs_input_grad[1:9]) else None
        return y_out

    @staticmethod
    def backward(ctx, gy):
        x, w_ih_cat, w_hh_cat, y_full, stash, mask = ctx.saved_tensors
        ds = ctx.ds
        drop_p, drop_seed, seed_dev = ctx.drop
        for q, v in ctx.param_versions:
            if q._version != v:
                raise RuntimeError("slu_b200: a GRU parameter needed for gradient computation has been modified by an inplace "
                                   "operation between forward and backward")
        B, T, I = x.shape
        dev = x.device
        gy = _f32(gy)
        dgx = torch.empty(B, T, 768, device=dev, dtype=torch.float32)
        dhn = torch.empty(B, T, 256, device=dev, dtype=torch.float32)
        ni = ctx.needs_input_grad
        wg = any(ni[1:9])
        # one zero-filled slot in parameter order: dW_ih [768,I] | dW_hh [2,384,H] | db_ih [2,384] | db_hh [2,384]
        n_ih, n_hh = 768 * I, 2 * 384 * H
        grads = [None] * 8
        fork = None
        db_ih = db_hh = dw_ih = dw_hh = None
        if wg:
            zbuf = _zeros(ctx, ctx.slot, (n_ih + n_hh + 2 * 768,))
            dw_ih, dw_hh = zbuf[:n_ih].view(768, I), zbuf[n_ih:n_ih + n_hh].view(2, 384, H)
            db_ih, db_hh = zbuf[n_ih + n_hh:n_ih + n_hh + 768], zbuf[n_ih + n_hh + 768:]
        if GRU_IMPL == "tc" and FUSED_BWD and _lib._prof is None:
            # the whole launch sequence of the layer in one C-ABI call (csrc/bigru.cu); the per-launch path below is what the
            # profiling pass uses, and the only one for the CUDA-core GRU variant
            dx = torch.empty(B, T, I, device=dev, dtype=torch.float32) if ni[0] else None
            img = None
            if ni[0]:
                img = ctx.img_nn if ctx.img_nn is not None else presplit(w_ih_cat, *_form_nn(w_ih_cat))
            _lib.call("slu_bigru_bwd_tc", _lib.ptr(gy), _lib.ptr(mask), drop_p, drop_seed, _lib.ptr(seed_dev), _lib.ptr(y_full), _lib.ptr(stash),
                      _lib.ptr(w_hh_cat), _lib.ptr(x),
                      I, None if img is None else img.data_ptr(), B, T, ds, _lib.ptr(dgx), _lib.ptr(dhn),
                      db_ih.data_ptr() if wg else None, db_hh.data_ptr() if wg else None,
                      dw_ih.data_ptr() if wg else None, dw_hh.data_ptr() if wg else None, _lib.ptr(dx), 1 if OVERLAP else 0,
                      _lib.stream())
            _lib.stats["calls"] += (3 if wg else 0) + (1 if ni[0] else 0)          # kernels launched beyond the first
        else:
            _lib.call("slu_gru_bwd_" + GRU_IMPL, _lib.ptr(gy), _lib.ptr(mask), drop_p, drop_seed, _lib.ptr(seed_dev), _lib.ptr(y_full), _lib.ptr(stash),
                      _lib.ptr(w_hh_cat), B, T, ds, _lib.ptr(dgx), _lib.ptr(dhn), db_ih.data_ptr() if wg else None, db_hh.data_ptr() if wg else None,
                      _lib.stream())
            if wg:
                fork = _Fork(3)
                wgrad_tc(dgx, 0, 768, 768, x, 0, I, I, B, T, dw_ih, 0, I, stream=fork.stream(0))
                for d in range(2):      # dW_hh[d] = [dr,dz | dhn]^T . h_{t-1}  (h_{t+1} for the reverse direction), one launch
                    wgrad2_tc(dgx, d * 384, 768, 256, dhn, d * H, 256, 384, y_full, d * H, 256, H, B, T, dw_hh, d * 384 * H, H,
                              shift0=1 if d else -1, stream=fork.stream(1 + d))
            dx = matmul_nn(dgx.view(B * T, 768), w_ih_cat, ctx.img_nn).view(B, T, I) if ni[0] else None
        if wg:
            for d in range(2):
                grads[4 * d + 0] = dw_ih[d * 384:(d + 1) * 384]
                grads[4 * d + 1] = dw_hh[d]
                grads[4 * d + 2] = db_ih[d * 384:(d + 1) * 384]
                grads[4 * d + 3] = db_hh[d * 384:(d + 1) * 384]
            if fork is not None:
                fork.join()
        return (dx, *grads, None, None, None, None, None)


_PACK_ORDER = ("weight_ih_l0", "weight_ih_l0_reverse", "weight_hh_l0", "weight_hh_l0_reverse",
               "bias_ih_l0", "bias_ih_l0_reverse", "bias_hh_l0", "bias_hh_l0_reverse")


def packed_params(gru):
    """Keep the 8 parameter tensors of a bidirectional nn.GRU holder in ONE buffer laid out as the kernels read them
    (W_ih both directions [768,I] | W_hh [2,384,128] | b_ih [768] | b_hh [2,384]) and return those four views.
    The Parameters stay the same objects (only their storage moves, as nn.GRU.flatten_parameters does for cuDNN), so
    optimizers, state_dict and checkpoints are unaffected; if something re-homed them (`.cuda()`, `.to()`), re-pack."""
    ps = [getattr(gru, n) for n in _PACK_ORDER]
    flat = getattr(gru, "_slu_flat", None)
    ok = flat is not None and flat.device == ps[0].device
    if ok:
        off = flat.data_ptr()
        for p in ps:
            if p.data_ptr() != off:
                ok = False
                break
            off += 4 * p.numel()
    if not ok:
        if any(p.dtype != torch.float32 for p in ps):
            return None
        flat = torch.empty(sum(p.numel() for p in ps), device=ps[0].device, dtype=torch.float32)
        off = 0
        with torch.no_grad():
            for p in ps:
                v = flat[off:off + p.numel()].view(p.shape)
                v.copy_(p)
                p.data = v
                off += p.numel()
        I = gru.input_size
        n1, n2 = 768 * I, 768 * I + 2 * 384 * H
        gru._slu_flat = flat
        gru._slu_views = (flat[:n1].view(768, I), flat[n2:n2 + 768], flat[n1:n2].view(2, 384, H), flat[n2 + 768:].view(2, 384))
    return gru._slu_views


def bigru(x, gru, mask=None, ds=1, before_recurrence=None, imgs=None):
    """Run BiGRU on the parameters of an nn.GRU holder module (imgs: operand images of its W_ih from gru_weight_items)."""
    return BiGRU.apply(x, gru.weight_ih_l0, gru.weight_hh_l0, gru.bias_ih_l0, gru.bias_hh_l0,
                       gru.weight_ih_l0_reverse, gru.weight_hh_l0_reverse, gru.bias_ih_l0_reverse,
                       gru.bias_hh_l0_reverse, mask, ds, packed_params(gru), before_recurrence, imgs)


def gru_weight_items(gru, need_dx):
    """presplit_many items of one GRU: W_ih of both directions as the x-projection operand (+ the input-gradient operand)."""
    packed = packed_params(gru)
    if packed is None:
        return None
    w = packed[0]
    return [(w, _form_nt(w))] + ([(w, _form_nn(w))] if need_dx else [])


def conv_weight_items(conv, need_dx):
    w = conv.weight.detach()
    if not w.is_contiguous():
        return None
    return [(w, _form_conv_fwd(w))] + ([(w, _form_conv_dx(w))] if need_dx else [])


_tickets = {}


def _ticket(dev):
    if dev.index not in _tickets:
        _tickets[dev.index] = torch.zeros(1, device=dev, dtype=torch.int32)
    return _tickets[dev.index]


def intent_head_supported(weight, slots):
    return weight.shape[1] == 2 * H and weight.shape[0] <= 128 and 1 <= len(slots) <= 16 and sum(slots) == weight.shape[0]


def _head_fwd(feats, w, b, y, slots):
    B, T, _ = feats.shape
    C = w.shape[0]
    dev = feats.device
    fbuf = torch.empty(B * C + 2 * B + 2, device=dev, dtype=torch.float32)
    tstar = torch.empty(B, C, device=dev, dtype=torch.int32)
    logits = fbuf[:B * C].view(B, C)
    sl = (ctypes.c_int * len(slots))(*slots)
    _lib.call("slu_intent_head_fwd", _lib.ptr(feats), _lib.ptr(w), _lib.ptr(b), None if y is None else _lib.ptr(y), B, T, C, sl,
              len(slots), logits.data_ptr(), _lib.ptr(tstar), fbuf[B * C:].data_ptr(), fbuf[B * C + B:].data_ptr(),
              fbuf[B * C + 2 * B:].data_ptr(), _lib.ptr(_ticket(dev)), _lib.stream())
    return fbuf, logits, tstar


class IntentLogits(torch.autograd.Function):
    """feats [B,T,256] -> logits [B,C] = max over time of Linear(256->C): models.py:806-809 (predict path), differentiable like
    the reference's (the arg-max frame of each class receives the gradient)."""

    @staticmethod
    def forward(ctx, feats, weight, bias):
        feats = _f32(feats)
        w, b = weight.detach().contiguous(), bias.detach().contiguous()
        C = w.shape[0]
        _, logits, tstar = _head_fwd(feats, w, b, None, (C,))
        ctx.save_for_backward(feats, w, tstar)
        ctx.slot = _reserve(ctx, w.numel() + C) if (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]) else None
        return logits

    @staticmethod
    def backward(ctx, g_logits):
        feats, w, tstar = ctx.saved_tensors
        B, T, _ = feats.shape
        C = w.shape[0]
        dev = feats.device
        g = _f32(g_logits)
        dfeats = torch.empty(B, T, 2 * H, device=dev, dtype=torch.float32)
        zb = _zeros(ctx, ctx.slot, (C * 2 * H + C,))
        sl = (ctypes.c_int * 1)(C)
        _lib.call("slu_intent_head_bwd", _lib.ptr(g), _lib.ptr(feats), _lib.ptr(w), None, None, _lib.ptr(tstar),
                  B, T, C, sl, 1, _lib.ptr(dfeats), zb.data_ptr(), zb[C * 2 * H:].data_ptr(), _lib.stream())
        return dfeats, zb[:C * 2 * H].view(C, 2 * H), zb[C * 2 * H:]


def intent_head_logits(feats, weight, bias):
    """Linear + max over time -> logits [B,C] (no labels): models.py:806-809 on the predict path."""
    return IntentLogits.apply(feats, weight, bias)


class IntentHead(torch.autograd.Function):
    """feats [B,T,256] -> (loss, acc, logits): Linear(256->C), max over time, summed per-slot cross-entropy, accuracy.
    Reference: models.py:709, :112-123, :811-823."""

    @staticmethod
    def forward(ctx, feats, weight, bias, y, slots):
        feats = _f32(feats)
        w, b = weight.detach().contiguous(), bias.detach().contiguous()
        y = y.contiguous()
        assert y.dtype == torch.int64 and y.shape == (feats.shape[0], len(slots))
        fbuf, logits, tstar = _head_fwd(feats, w, b, y, slots)
        ctx.save_for_backward(feats, w, y, logits, tstar)
        ctx.slots = slots
        ctx.slot = _reserve(ctx, w.numel() + w.shape[0]) if (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]) else None
        loss, acc = fbuf[-2], fbuf[-1]
        ctx.mark_non_differentiable(acc, logits)
        return loss, acc, logits

    @staticmethod
    def backward(ctx, g_loss, g_acc, g_logits):
        feats, w, y, logits, tstar = ctx.saved_tensors
        slots = ctx.slots
        B, T, _ = feats.shape
        C = w.shape[0]
        dev = feats.device
        g = _f32(g_loss).reshape(1)
        dfeats = torch.empty(B, T, 2 * H, device=dev, dtype=torch.float32)
        zb = _zeros(ctx, ctx.slot, (C * 2 * H + C,))
        sl = (ctypes.c_int * len(slots))(*slots)
        _lib.call("slu_intent_head_bwd", _lib.ptr(g), _lib.ptr(feats), _lib.ptr(w), _lib.ptr(y), _lib.ptr(logits), _lib.ptr(tstar),
                  B, T, C, sl, len(slots), _lib.ptr(dfeats), zb.data_ptr(), zb[C * 2 * H:].data_ptr(), _lib.stream())
        return dfeats, zb[:C * 2 * H].view(C, 2 * H), zb[C * 2 * H:], None, None


def gru_rows_per_cta(B):
    """Batch rows one CTA of the tcgen05 recurrence carries (csrc/gru_tc.cu pick_rows, exported as slu_gru_rows_per_cta)."""
    return int(_lib.load().slu_gru_rows_per_cta(int(B)))


def gru_executed_flop_factor(B):
    """Executed / algorithmic tensor-core flops of the recurrence: the N=16 tile carries NR rows, hi/lo stacked, x (W_hi, W_lo)."""
    nr = gru_rows_per_cta(B)
    return 32 // nr if nr < 16 else 3


# ---- ASR heads ----------------------------------------------------------------------------------------
CE_CHUNK = int(os.environ.get("SLU_CE_CHUNK", "4096"))      # frames per logits tile (x V floats: 164 MB at V = 10 000)


def _pad4_rows(w, b, neg=-1e30):
    """Weight [V,K] / bias [V] with V padded to a multiple of 4 (zero rows, bias -1e30: the padded classes get probability 0)."""
    V, K = w.shape
    Vp = (V + 3) // 4 * 4
    if Vp == V:
        return w, b, V
    wp = torch.zeros(Vp, K, device=w.device, dtype=torch.float32)
    wp[:V].copy_(w)
    bp = torch.full((Vp,), neg, device=w.device, dtype=torch.float32)
    bp[:V].copy_(b)
    return wp, bp, Vp


class LinearCE(torch.autograd.Function):
    """feats [M,K] -> (loss, acc): Linear(K -> V) + cross_entropy(ignore_index=-1, mean) + masked arg-max accuracy, reference
    models.py:308-314 / 321-329, without ever holding the [M,V] logits: the frames are walked in chunks of CE_CHUNK rows whose
    logits tile is overwritten by its own gradient and consumed by the three gradient GEMMs before the next chunk (csrc/ce.cu).
    The gradients are therefore produced in forward (scaled for dL/dloss = 1) and multiplied by the incoming dL/dloss in backward."""

    @staticmethod
    def forward(ctx, feats, weight, bias, y):
        x = _f32(feats)
        M, K = x.shape
        dev = x.device
        w, b, Vp = _pad4_rows(weight.detach().contiguous(), bias.detach().contiguous())
        V = weight.shape[0]
        y = y.contiguous()
        assert y.dtype == torch.int64 and y.numel() == M and K % 4 == 0
        need_x = ctx.needs_input_grad[0]
        need_w = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        grad = need_x or need_w
        st = _lib.stream()
        sbuf = torch.empty(2 * M + 4, device=dev, dtype=torch.float32)          # row_loss | row_ok | n_valid, 1/n_valid | loss, acc
        row_loss, row_ok, nv, out = sbuf[:M], sbuf[M:2 * M], sbuf[2 * M:2 * M + 2], sbuf[2 * M + 2:]
        _lib.call("slu_ce_count", _lib.ptr(y), M, nv.data_ptr(), st)
        R = min(M, max(128, CE_CHUNK))
        tile = torch.empty(R, Vp, device=dev, dtype=torch.float32)
        img_nt = presplit(w, *_form_nt(w))
        img_nn = presplit(w, *_form_nn(w)) if need_x else None
        dx = torch.empty(M, K, device=dev, dtype=torch.float32) if need_x else None
        dwb = torch.zeros(Vp * K + Vp, device=dev, dtype=torch.float32) if need_w else None     # dW | db, accumulated over chunks
        for r0 in range(0, M, R):
            r = min(R, M - r0)
            gemm_tc(x[r0:r0 + r], K, img_nt, r, Vp, K, tile, bias=b)
            _lib.call("slu_ce_rows", tile.data_ptr(), Vp, Vp, y[r0:].data_ptr(), r, nv.data_ptr(), 1 if grad else 0,
                      row_loss[r0:].data_ptr(), row_ok[r0:].data_ptr(), st)
            if need_x:
                gemm_tc(tile, Vp, img_nn, r, K, Vp, dx[r0:r0 + r])
            if need_w:
                wgrad_tc(tile, 0, Vp, Vp, x, r0 * K, K, K, 1, r, dwb, 0, K)
                _lib.call("slu_colsum_acc", tile.data_ptr(), Vp, r, Vp, dwb[Vp * K:].data_ptr(), st)
        _lib.call("slu_ce_finish", row_loss.data_ptr(), row_ok.data_ptr(), M, nv.data_ptr(), out.data_ptr(), st)
        ctx.save_for_backward(dx, dwb)
        ctx.dims = (V, Vp, K)
        ctx.slot = _reserve(ctx, Vp * K + Vp) if need_w else None
        loss, acc = out[0], out[1]
        ctx.mark_non_differentiable(acc, row_loss)
        return loss, acc, row_loss              # row_loss [M]: per-row NLL (0 on ignored rows), values only

    @staticmethod
    def backward(ctx, g_loss, g_acc, g_rows):
        dx, dwb = ctx.saved_tensors
        V, Vp, K = ctx.dims
        g = _f32(g_loss).reshape(1)
        st = _lib.stream()
        gx = gw = gb = None
        if dx is not None:
            gx = torch.empty_like(dx)
            _lib.call("slu_scale", dx.data_ptr(), gx.data_ptr(), dx.numel(), g.data_ptr(), st)
        if dwb is not None:
            out = _zeros(ctx, ctx.slot, (Vp * K + Vp,))
            _lib.call("slu_scale", dwb.data_ptr(), out.data_ptr(), dwb.numel(), g.data_ptr(), st)
            gw, gb = out[:V * K].view(V, K), out[Vp * K:Vp * K + V]
        return gx, gw, gb, None


def linear_ce(feats, weight, bias, y):
    """(loss, acc) of a frame-wise classification head; feats [..., K], y [...] int64 with -1 = ignore."""
    return LinearCE.apply(feats.reshape(-1, feats.shape[-1]), weight, bias, y.reshape(-1))[:2]


class LinearNT(torch.autograd.Function):
    """x [..., K] @ W[V,K]^T + b on the tcgen05 GEMMs, differentiable (compute_posteriors, models.py:333-347)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x2 = _f32(x).reshape(-1, x.shape[-1])
        w, b, Vp = _pad4_rows(weight.detach().contiguous(), bias.detach().contiguous(), neg=0.0)
        out = linear_nt(x2, w, b)
        ctx.save_for_backward(x2, w)
        ctx.shape = tuple(x.shape)
        V = weight.shape[0]
        return (out if Vp == V else out[:, :V].contiguous()).view(*x.shape[:-1], V)

    @staticmethod
    def backward(ctx, gy):
        x2, w = ctx.saved_tensors
        Vp, K = w.shape
        M = x2.shape[0]
        V = gy.shape[-1]
        g = _f32(gy).reshape(M, V)
        if Vp != V:
            g = torch.nn.functional.pad(g, (0, Vp - V))
        gx = matmul_nn(g, w).view(ctx.shape) if ctx.needs_input_grad[0] else None
        gw = gb = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            dwb = torch.zeros(Vp * K + Vp, device=g.device, dtype=torch.float32)
            wgrad_tc(g, 0, Vp, Vp, x2, 0, K, K, 1, M, dwb, 0, K)
            _lib.call("slu_colsum_acc", g.data_ptr(), Vp, M, Vp, dwb[Vp * K:].data_ptr(), _lib.stream())
            gw, gb = dwb[:V * K].view(V, K), dwb[Vp * K:Vp * K + V]
        return gx, gw, gb
