"""Autograd wrappers over the C-ABI kernels (CUDA tensors only).

Layouts are the library's own (time-major NLC everywhere):
  waveform [B,T] -> sinc frames [B,L1,80] -> conv blocks [B,L1,60] -> GRU stacks [B,T_l,256].
"""
import os

import torch

from . import _lib

H = 128
# Which persistent-GRU kernel family runs the recurrence: "tc" = tcgen05 (weights stationary in TMEM),
# "simt" = fp32 CUDA-core variant.  Both are sm_100a kernels of this library with identical contracts.
GRU_IMPL = os.environ.get("SLU_GRU_IMPL", "simt")


def _f32(t):
    return t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()


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
        _lib.call("slu_sincconv_fwd_simt", _lib.ptr(x), _lib.ptr(W), B, T, _lib.ptr(out), _lib.ptr(route), _lib.stream())
        if need:
            ctx.save_for_backward(x, b1, band, route)
        return out

    @staticmethod
    def backward(ctx, gy):
        x, b1, band, route = ctx.saved_tensors
        B, T = x.shape
        gy = _f32(gy)
        dW = torch.empty(80, 401, device=x.device, dtype=torch.float32)
        _lib.call("slu_sincconv_bwd_simt", _lib.ptr(x), _lib.ptr(gy), _lib.ptr(route), B, T, _lib.ptr(dW), _lib.stream())
        d_b1 = torch.empty(80, device=x.device, dtype=torch.float64)
        d_band = torch.empty(80, device=x.device, dtype=torch.float64)
        _lib.call("slu_sinc_filters_bwd", _lib.ptr(b1), _lib.ptr(band), _lib.ptr(dW), _lib.ptr(d_b1), _lib.ptr(d_band),
                  _lib.stream())
        return None, d_b1, d_band


def sinc_filters(filt_b1, filt_band):
    """W[80,401] fp32 on the device (models.py:82-106), for tests / inspection."""
    b1 = filt_b1.detach().double().contiguous()
    band = filt_band.detach().double().contiguous()
    W = torch.empty(80, 401, device=b1.device, dtype=torch.float32)
    _lib.call("slu_sinc_filters_fwd", _lib.ptr(b1), _lib.ptr(band), _lib.ptr(W), _lib.stream())
    return W


def conv_block_nlc(x, weight, bias, negative_slope=0.2):
    """Conv1d(k, pad=k//2) + LeakyReLU on NLC input (models.py:200-220), as one dense GEMM over
    the k time-shifted views.  x [B,T,Cin], weight [Cout,Cin,k] (reference layout) -> [B,T,Cout]."""
    B, T, Cin = x.shape
    Cout, _, k = weight.shape
    xp = torch.nn.functional.pad(x, (0, 0, k // 2, k // 2))
    cols = torch.cat([xp[:, d:d + T, :] for d in range(k)], dim=2)            # [B,T,k*Cin]
    wm = weight.permute(2, 1, 0).reshape(k * Cin, Cout)
    out = torch.addmm(bias, cols.reshape(B * T, k * Cin), wm).view(B, T, Cout)
    return torch.nn.functional.leaky_relu(out, negative_slope)


class BiGRU(torch.autograd.Function):
    """Bidirectional single-layer GRU (H=128, h0=0) + Dropout(mask) + Downsample(avg 2 | none 1).
    Reference: nn.GRU at models.py:232/262/686, RNNSelect :138-149, Dropout :246, Downsample :26-46.
    x [B,T,I] -> [B, ceil(T/ds), 256]."""

    @staticmethod
    def forward(ctx, x, w_ih, w_hh, b_ih, b_hh, w_ih_r, w_hh_r, b_ih_r, b_hh_r, mask, ds):
        x = _f32(x)
        B, T, I = x.shape
        dev = x.device
        w_ih_cat = torch.cat([w_ih, w_ih_r], 0).detach()                       # [768, I]
        b_ih_cat = torch.cat([b_ih, b_ih_r], 0).detach()
        w_hh_cat = torch.stack([w_hh, w_hh_r], 0).detach().contiguous()       # [2,384,128]
        b_hh_cat = torch.stack([b_hh, b_hh_r], 0).detach().contiguous()
        gx = torch.addmm(b_ih_cat, x.view(B * T, I), w_ih_cat.t())            # x-projection, both directions
        T2 = (T + ds - 1) // ds
        y_full = torch.empty(B, T, 256, device=dev, dtype=torch.float32)
        y_out = torch.empty(B, T2, 256, device=dev, dtype=torch.float32) if (ds != 1 or mask is not None) else y_full
        need = any(ctx.needs_input_grad[:9])
        stash = torch.empty(B, T, 1024, device=dev, dtype=torch.float32) if need else None
        _lib.call("slu_gru_fwd_" + GRU_IMPL, _lib.ptr(gx), _lib.ptr(w_hh_cat), _lib.ptr(b_hh_cat), _lib.ptr(mask), B, T, ds,
                  _lib.ptr(y_full), _lib.ptr(y_out), _lib.ptr(stash), _lib.stream())
        if need:
            ctx.save_for_backward(x, w_ih_cat, w_hh_cat, y_full, stash, mask)
            ctx.ds = ds
        return y_out

    @staticmethod
    def backward(ctx, gy):
        x, w_ih_cat, w_hh_cat, y_full, stash, mask = ctx.saved_tensors
        ds = ctx.ds
        B, T, I = x.shape
        dev = x.device
        gy = _f32(gy)
        dgx = torch.empty(B, T, 768, device=dev, dtype=torch.float32)
        dhn = torch.empty(B, T, 256, device=dev, dtype=torch.float32)
        _lib.call("slu_gru_bwd_" + GRU_IMPL, _lib.ptr(gy), _lib.ptr(mask), _lib.ptr(y_full), _lib.ptr(stash), _lib.ptr(w_hh_cat),
                  B, T, ds, _lib.ptr(dgx), _lib.ptr(dhn), _lib.stream())
        ni = ctx.needs_input_grad
        dgx2 = dgx.view(B * T, 768)
        dx = (dgx2 @ w_ih_cat).view(B, T, I) if ni[0] else None
        grads = [None] * 8
        if any(ni[1:9]):
            x2 = x.view(B * T, I)
            dw_ih = dgx2.t() @ x2                                               # [768, I]
            db_ih = dgx2.sum(0)
            zero = torch.zeros(B, 1, H, device=dev, dtype=torch.float32)
            for d in range(2):
                hd = y_full[:, :, d * H:(d + 1) * H]
                hprev = torch.cat([zero, hd[:, :-1]], 1) if d == 0 else torch.cat([hd[:, 1:], zero], 1)
                gd = torch.cat([dgx[:, :, d * 384:d * 384 + 256], dhn[:, :, d * H:(d + 1) * H]], 2).reshape(B * T, 384)
                dw_hh = gd.t() @ hprev.reshape(B * T, H)
                db_hh = gd.sum(0)
                grads[4 * d + 0] = dw_ih[d * 384:(d + 1) * 384]
                grads[4 * d + 1] = dw_hh
                grads[4 * d + 2] = db_ih[d * 384:(d + 1) * 384]
                grads[4 * d + 3] = db_hh
        return (dx, *grads, None, None)


def bigru(x, gru, mask=None, ds=1):
    """Run BiGRU on the parameters of an nn.GRU holder module."""
    return BiGRU.apply(x, gru.weight_ih_l0, gru.weight_hh_l0, gru.bias_ih_l0, gru.bias_hh_l0,
                       gru.weight_ih_l0_reverse, gru.weight_hh_l0_reverse, gru.bias_ih_l0_reverse,
                       gru.bias_hh_l0_reverse, mask, ds)
