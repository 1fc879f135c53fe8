"""Design-time study (CPU): which tensor-core operand format keeps the intent logits within
the north-star tolerance (1e-3 relative, max-abs / max-abs)?  Emulates operand rounding of
every contraction on the path (sinc conv, CNN tail, x.W_ih, h.W_hh) with fp32 accumulation.
TEST/DESIGN INFRASTRUCTURE ONLY -- never imported by the product.  Results are recorded in DESIGN.md.

  python oracle/precision_study.py [ckpt.pth wav]   (defaults to synthetic params, 4 s synthetic batch)
"""
import sys, os
import torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import torch_ref as R


def q_bf16(t): return t.bfloat16().float()
def q_tf32_trunc(t): return (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)
def q_tf32_rn(t):
    i = t.contiguous().view(torch.int32)
    return ((i + 0xFFF + ((i >> 13) & 1)) & ~0x1FFF).view(torch.float32)
def q_none(t): return t


def mm(a, b, mode):
    """a @ b with emulated operand format."""
    if mode == "fp32": return a @ b
    if mode == "bf16x3":
        ah, bh = q_bf16(a), q_bf16(b); al, bl = q_bf16(a - ah), q_bf16(b - bh)
        return ah @ bh + (ah @ bl + al @ bh)
    if mode == "bf16x4":   # all four partial products: what the GRU's N-stacked hi/lo operand tile computes (W_hi, W_lo) x [h_hi; h_lo]
        ah, bh = q_bf16(a), q_bf16(b); al, bl = q_bf16(a - ah), q_bf16(b - bh)
        return (ah @ bh + al @ bh) + (ah @ bl + al @ bl)
    if mode == "fp16":     # single fp16 pass (the optional recurrence mode)
        return a.half().float() @ b.half().float()
    if mode == "bf16x2":   # activation hi/lo, weights bf16 only
        ah = q_bf16(a); al = q_bf16(a - ah); bh = q_bf16(b)
        return ah @ bh + al @ bh
    q = {"bf16": q_bf16, "tf32t": q_tf32_trunc, "tf32r": q_tf32_rn}[mode]
    return q(a) @ q(b)


def forward(x, p, mode_conv, mode_xw, mode_hu):
    P = R.P
    w = R.sinc_filters(p[P + "phoneme_layers.0.filt_b1"], p[P + "phoneme_layers.0.filt_band"])
    xp = F.pad(x, (200, 200))
    cols = xp.unfold(1, 401, 80)                       # [B, L0, 401]
    out = mm(cols, w.t(), mode_conv).transpose(1, 2)   # [B,80,L0]
    out = F.max_pool1d(out.abs(), 2, ceil_mode=True)
    for k in ("5", "9"):
        wt, b = p[P + "phoneme_layers.%s.weight" % k], p[P + "phoneme_layers.%s.bias" % k]
        cols = F.pad(out, (2, 2)).unfold(2, 5, 1)      # [B,C,T,5]
        cols = cols.permute(0, 2, 1, 3).reshape(out.shape[0], out.shape[2], -1)
        out = F.leaky_relu(mm(cols, wt.reshape(wt.shape[0], -1).t(), mode_conv) + b, 0.2).transpose(1, 2)
    out = out.transpose(1, 2)
    for li, key in enumerate(R.GRU_KEYS):
        ys = []
        for sfx, rev in (("", False), ("_reverse", True)):
            wi, wh = p[key + ".weight_ih_l0" + sfx], p[key + ".weight_hh_l0" + sfx]
            bi, bh = p[key + ".bias_ih_l0" + sfx], p[key + ".bias_hh_l0" + sfx]
            gx = mm(out, wi.t(), mode_xw) + bi
            B, T, _ = out.shape
            h = out.new_zeros(B, 128); o = [None] * T
            for t in (range(T - 1, -1, -1) if rev else range(T)):
                gh = mm(h, wh.t(), mode_hu) + bh
                r = torch.sigmoid(gx[:, t, :128] + gh[:, :128]); z = torch.sigmoid(gx[:, t, 128:256] + gh[:, 128:256])
                n = torch.tanh(gx[:, t, 256:] + r * gh[:, 256:]); h = (1 - z) * n + z * h; o[t] = h
            ys.append(torch.stack(o, 1))
        out = R.downsample(torch.cat(ys, 2), *R.GRU_DOWNSAMPLE[li])
    out = out @ p["intent_layers.4.weight"].t() + p["intent_layers.4.bias"]
    return out.max(1)[0]


if __name__ == "__main__":
    torch.set_grad_enabled(False)
    if len(sys.argv) > 2:
        import wave, numpy as np
        p = torch.load(sys.argv[1], map_location="cpu")
        w = wave.open(sys.argv[2]); x = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16).astype(np.float32) / 32768
        x = torch.tensor(x)[None]
    else:
        p = R.synthetic_params(seed=0); x, _ = R.synthetic_batch(4, 64000)
    ref = forward(x, p, "fp32", "fp32", "fp32")
    print("ref max|logit| %.3f" % ref.abs().max())
    for modes in [("bf16",) * 3, ("tf32t",) * 3, ("tf32r",) * 3, ("bf16x2",) * 3, ("bf16x3",) * 3,
                  ("bf16x3", "bf16x3", "bf16"), ("bf16x3", "bf16x3", "tf32r"), ("bf16x3", "tf32r", "tf32r"),
                  ("bf16x3", "bf16", "bf16x3"), ("bf16", "bf16x3", "bf16x3"), ("tf32r", "bf16x3", "bf16x3"),
                  ("bf16x3", "bf16x3", "bf16x4"), ("bf16x3", "bf16x3", "fp16")]:
        out = forward(x, p, *modes)
        print("conv=%-7s xw=%-7s hu=%-7s  rel err %.2e" % (*modes, ((out - ref).abs().max() / ref.abs().max()).item()))
