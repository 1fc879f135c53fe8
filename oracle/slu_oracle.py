"""CPU oracle (plain numpy, forward only) of the speech-encoder hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file; the
product path never does.  A second, independent restatement next to oracle/torch_ref.py (which
adds autograd): no torch ops here, explicit loops over time / taps, fp32 storage with fp32
accumulation (numpy matmul), fp64 only where the reference uses fp64 (the sinc cut-offs).
Parity is PINNED by tests/test_oracle_golden.py against fixtures produced by the real reference.

Reference lines restated: models.py:17-24, 77-110 (SincLayer), 163-168 (Abs), 205 (MaxPool1d
ceil), 200-220 (Conv1d + LeakyReLU), 26-46 (Downsample), 112-123 (FinalPool), 232/262/686
(nn.GRU; formula in SURVEY.md 8a4), 349-361 (compute_features), 830-846 (predict_intents).
"""
import math

import numpy as np

P = "pretrained_model."
GRU_KEYS = [P + "phoneme_layers.14", P + "phoneme_layers.18", P + "word_layers.0", P + "word_layers.4",
            "intent_layers.0"]
GRU_DOWNSAMPLE = ["avg", "avg", "avg", "avg", "none"]
f32 = np.float32


def sinc_filters(filt_b1, filt_band, fs=16000, n_taps=401):
    """models.py:82-106.  fp64 params -> per-filter fp32 cast (:99-100) -> fp32 arithmetic."""
    half = (n_taps - 1) // 2
    t_right = (np.linspace(1, half, half).astype(f32) / f32(fs)).astype(f32)
    beg = np.abs(filt_b1.astype(np.float64)) + 50.0 / fs
    end = beg + (np.abs(filt_band.astype(np.float64)) + 50.0 / fs)
    j = np.arange(n_taps, dtype=np.float64)
    n = (j * (n_taps / (n_taps - 1.0))).astype(f32)                  # torch.linspace(0, N, steps=N)
    window = (f32(0.54) - f32(0.46) * np.cos((f32(2 * math.pi) * n / f32(n_taps)).astype(f32))).astype(f32)
    W = np.empty((len(beg), n_taps), f32)
    for i in range(len(beg)):
        def low_pass(f):
            fl = f32(f)
            arg = (f32(2 * math.pi * float(fl * f32(fs))) * t_right).astype(f32)
            y_right = (np.sin(arg) / arg).astype(f32)
            return (f32(2) * fl * np.concatenate([y_right[::-1], np.ones(1, f32), y_right])).astype(f32)
        bp = low_pass(end[i]) - low_pass(beg[i])
        W[i] = (bp / bp.max()) * window
    return W


def sinc_frontend(x, filt_b1, filt_band, fs=16000):
    """x[B,T] -> [B,80,ceil(L0/2)]: strided correlation (:108), abs, max-pool(2, ceil)."""
    W = sinc_filters(filt_b1, filt_band, fs)
    B, T = x.shape
    L0 = (T - 1) // 80 + 1
    xp = np.zeros((B, T + 400), f32); xp[:, 200:200 + T] = x
    idx = (np.arange(L0) * 80)[:, None] + np.arange(401)[None, :]
    out = np.abs(xp[:, idx] @ W.T).transpose(0, 2, 1)               # [B,80,L0]
    L1 = (L0 + 1) // 2
    pad = np.full((B, 80, 2 * L1), -np.inf, f32); pad[:, :, :L0] = out
    return pad.reshape(B, 80, L1, 2).max(-1)


def conv_block(x, weight, bias):
    """Conv1d(k=5, p=2) + LeakyReLU(0.2), NCL layout (:200-220)."""
    B, C, T = x.shape
    K = weight.shape[2]
    xp = np.zeros((B, C, T + K - 1), f32); xp[:, :, K // 2:K // 2 + T] = x
    out = np.zeros((B, weight.shape[0], T), f32)
    for d in range(K):
        out += np.einsum("oc,bct->bot", weight[:, :, d], xp[:, :, d:d + T]).astype(f32)
    out += bias[None, :, None]
    return np.where(out > 0, out, f32(0.2) * out).astype(f32)


def _sigmoid(v):
    return (1.0 / (1.0 + np.exp(-v))).astype(f32)


def gru_direction(x, w_ih, w_hh, b_ih, b_hh, reverse):
    B, T, _ = x.shape
    H = w_hh.shape[1]
    gx = (x @ w_ih.T + b_ih).astype(f32)
    h = np.zeros((B, H), f32)
    out = np.empty((B, T, H), f32)
    for t in (range(T - 1, -1, -1) if reverse else range(T)):
        gh = (h @ w_hh.T + b_hh).astype(f32)
        r = _sigmoid(gx[:, t, :H] + gh[:, :H])
        z = _sigmoid(gx[:, t, H:2 * H] + gh[:, H:2 * H])
        n = np.tanh(gx[:, t, 2 * H:] + r * gh[:, 2 * H:]).astype(f32)
        h = ((1 - z) * n + z * h).astype(f32)
        out[:, t] = h
    return out


def bigru(x, p, key):
    f = gru_direction(x, p[key + ".weight_ih_l0"], p[key + ".weight_hh_l0"], p[key + ".bias_ih_l0"],
                      p[key + ".bias_hh_l0"], False)
    b = gru_direction(x, p[key + ".weight_ih_l0_reverse"], p[key + ".weight_hh_l0_reverse"],
                      p[key + ".bias_ih_l0_reverse"], p[key + ".bias_hh_l0_reverse"], True)
    return np.concatenate([f, b], axis=2)


def downsample_avg2(x):
    """avg_pool1d(k=2, ceil_mode=True) over time: an odd tail frame is divided by 1 (:44)."""
    B, T, C = x.shape
    T2 = (T + 1) // 2
    out = np.empty((B, T2, C), f32)
    out[:, :T // 2] = (x[:, 0:2 * (T // 2):2] + x[:, 1:2 * (T // 2):2]) * f32(0.5)
    if T % 2:
        out[:, -1] = x[:, -1]
    return out


def intent_logits(x, p, return_all=False):
    acts = {}
    out = sinc_frontend(x, p[P + "phoneme_layers.0.filt_b1"], p[P + "phoneme_layers.0.filt_band"]); acts["sinc"] = out
    out = conv_block(out, p[P + "phoneme_layers.5.weight"], p[P + "phoneme_layers.5.bias"]); acts["conv1"] = out
    out = conv_block(out, p[P + "phoneme_layers.9.weight"], p[P + "phoneme_layers.9.bias"]); acts["conv2"] = out
    out = out.transpose(0, 2, 1)
    for li, key in enumerate(GRU_KEYS):
        out = bigru(out, p, key)
        acts["gru%d_raw" % li] = out
        if GRU_DOWNSAMPLE[li] == "avg":
            out = downsample_avg2(out)
        acts["gru%d" % li] = out
        if li == 3:
            acts["features"] = out
    out = out @ p["intent_layers.4.weight"].T + p["intent_layers.4.bias"]
    logits = out.max(axis=1)
    return (logits, acts) if return_all else logits


def predict(logits, values_per_slot=(6, 14, 4)):
    pred, s = [], 0
    for n in values_per_slot:
        pred.append(logits[:, s:s + n].argmax(1)); s += n
    return np.stack(pred, 1)
