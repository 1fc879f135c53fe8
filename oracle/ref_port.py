"""Timing port of the reference's execution structure.  TEST/BENCH INFRASTRUCTURE ONLY
(bench.py's cpu_baseline / reference_gpu / --impl reference legs; never on the product path).

Same math as oracle/torch_ref.py, but built from the library modules the reference itself uses
(nn.GRU, F.conv1d, nn.Dropout, nn.Linear, F.cross_entropy ...) and -- with loop80=True -- bug-compatible
with the reference's SincLayer.forward, which re-runs the full conv1d inside its 80-iteration filter
loop (models.py:98-108: 80 convolutions + ~25 small ops + 4 host->device copies per filter per
forward: `torch.arange(..).cuda()` twice in flip() :12-13 and `torch.ones(1).cuda()` twice in sinc()
:21 -- all four are present below as `.to(dev)` of CPU-made tensors).  That is what "the reference's
own implementation" costs, on CPU (mkldnn + aten::gru) or on GPU (cuDNN conv + cuDNN RNN + cuBLAS).

kinds (BASELINE.json configs):
  "slu"     Model.forward, everything trainable                       (config 3; models.py:797-823)
  "frozen"  Model.forward after freeze_all_layers(): intent module only (config 2; models.py:738-742)
  "asr"     PretrainedModel.forward, pretraining_type 2: frame-wise CE phoneme + word heads
                                                                        (config 4; models.py:291-331)
  "seq2seq" Model.forward with the attention decoder, teacher forced   (config 5; models.py:500-556)
What the port omits relative to the reference: nothing on the timed path for "slu"/"frozen"/"asr";
for "seq2seq" the decoder is this repo's seq2seq.py (same per-step library calls as models.py:413-556).
"""
import math
import os
import sys

import torch
import torch.nn.functional as F

from . import torch_ref as R

GRU_IN = (60, 256, 256, 256, 256)


class RefPort(torch.nn.Module):
    def __init__(self, params, loop80=True, drop_p=0.5, kind="slu", n_labels=102):
        super().__init__()
        P = R.P
        self.kind = kind
        self.filt_b1 = torch.nn.Parameter(params[P + "phoneme_layers.0.filt_b1"].clone())
        self.filt_band = torch.nn.Parameter(params[P + "phoneme_layers.0.filt_band"].clone())
        self.conv1 = torch.nn.Conv1d(80, 60, 5, padding=2)
        self.conv2 = torch.nn.Conv1d(60, 60, 5, padding=2)
        n_enc = 4 if kind in ("asr", "seq2seq") else 5
        self.grus = torch.nn.ModuleList([torch.nn.GRU(I, 128, batch_first=True, bidirectional=True) for I in GRU_IN[:n_enc]])
        self.drop = torch.nn.Dropout(drop_p)
        self.loop80 = loop80
        with torch.no_grad():
            self.conv1.weight.copy_(params[P + "phoneme_layers.5.weight"]); self.conv1.bias.copy_(params[P + "phoneme_layers.5.bias"])
            self.conv2.weight.copy_(params[P + "phoneme_layers.9.weight"]); self.conv2.bias.copy_(params[P + "phoneme_layers.9.bias"])
            for gru, key in zip(self.grus, R.GRU_KEYS):
                for n, q in gru.named_parameters():
                    q.copy_(params[key + "." + n])
        if kind in ("slu", "frozen"):
            self.final = torch.nn.Linear(256, 24)
            with torch.no_grad():
                self.final.weight.copy_(params["intent_layers.4.weight"]); self.final.bias.copy_(params["intent_layers.4.bias"])
        if kind == "asr":
            self.phoneme_linear = torch.nn.Linear(256, 42)
            self.word_linear = torch.nn.Linear(256, 10000)
        if kind == "seq2seq":
            root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
            if root not in sys.path:
                sys.path.insert(0, root)
            import seq2seq as S
            self.s2s_encoder = torch.nn.GRU(256, 128, batch_first=True, bidirectional=True)
            self.s2s_decoder = S.Seq2SeqDecoder(n_labels, 2, 128, 256, 100, 200, SOS=0)
        if kind == "frozen":       # freeze_all_layers(): phoneme + word modules (models.py:738-742)
            for m in [self.conv1, self.conv2] + list(self.grus[:4]):
                for q in m.parameters():
                    q.requires_grad = False
            self.filt_b1.requires_grad = False
            self.filt_band.requires_grad = False

    def sinc(self, x):
        dev = x.device
        fs, N = 16000, 401
        if not self.loop80:
            w = R.sinc_filters(self.filt_b1.cpu(), self.filt_band.cpu()).to(dev) if dev.type != "cpu" else R.sinc_filters(self.filt_b1, self.filt_band)
            return F.conv1d(x, w.unsqueeze(1), stride=80, padding=200)
        # structure of models.py:77-110, per-filter python loop with the conv inside it
        filters = torch.zeros((80, N)).to(dev)
        t_right = (torch.linspace(1, (N - 1) / 2, steps=int((N - 1) / 2)) / fs).to(dev)
        beg = torch.abs(self.filt_b1) + 50.0 / fs
        end = beg + (torch.abs(self.filt_band) + 50.0 / fs)
        n = torch.linspace(0, N, steps=N)
        window = (0.54 - 0.46 * torch.cos(2 * math.pi * n / N)).float().to(dev)
        out = None
        for i in range(80):
            def lp(f):
                f = f.float()
                y_right = torch.sin(2 * math.pi * (f * fs) * t_right) / (2 * math.pi * (f * fs) * t_right)
                idx = torch.arange(y_right.shape[0] - 1, -1, -1).to(dev)          # flip(): H2D copy of the index vector
                return 2 * f * torch.cat([y_right[idx], torch.ones(1).to(dev), y_right])   # sinc(): H2D copy of the centre tap
            bp = lp(end[i]) - lp(beg[i])
            bp = bp / torch.max(bp)
            filters[i, :] = bp * window
            out = F.conv1d(x, filters.view(80, 1, N), stride=80, padding=200)
        return out

    def features(self, x, n):
        out = self.sinc(x.unsqueeze(1))
        out = F.leaky_relu(F.max_pool1d(torch.abs(out), 2, ceil_mode=True), 0.2)
        out = F.leaky_relu(self.conv1(out), 0.2)
        out = F.leaky_relu(self.conv2(out), 0.2)
        out = out.transpose(1, 2)
        mids = []
        for li in range(n):
            out = self.drop(self.grus[li](out)[0])
            out = R.downsample(out, *R.GRU_DOWNSAMPLE[li])
            mids.append(out)
        return out, mids

    def forward(self, x, *ys):
        if self.kind in ("slu", "frozen"):
            out, _ = self.features(x, 5)
            logits = self.final(out).max(dim=1)[0]
            loss, acc, _ = R.intent_loss_acc(logits, ys[0])
            return loss
        if self.kind == "asr":
            y_phoneme, y_word = ys
            out, mids = self.features(x, 4)
            lp = self.phoneme_linear(mids[1])
            lp = lp.view(lp.shape[0] * lp.shape[1], -1)
            phoneme_loss = F.cross_entropy(lp, y_phoneme.view(-1), ignore_index=-1)
            lw = self.word_linear(out)
            lw = lw.view(lw.shape[0] * lw.shape[1], -1)
            word_loss = F.cross_entropy(lw, y_word.view(-1), ignore_index=-1)
            valid = y_word.view(-1) != -1
            _ = (lw.max(1)[1][valid] == y_word.view(-1)[valid]).float().mean()
            return phoneme_loss + word_loss
        out, _ = self.features(x, 4)
        out = self.drop(self.s2s_encoder(out)[0])
        return -self.s2s_decoder(out, ys[0]).mean()


def synthetic_labels(kind, B, T, seed, n_labels=102, U=40):
    """Labels of the shapes SURVEY.md 8(d) lists: intents [B,3]; ASR frame labels with -1 padding; one-hot strings."""
    import numpy as np
    rs = np.random.RandomState(seed + 77)
    if kind in ("slu", "frozen"):
        return (torch.from_numpy(np.stack([rs.randint(0, v, size=B) for v in R.VALUES_PER_SLOT], axis=1).astype(np.int64)),)
    if kind == "asr":
        tp, tw = -(-T // 640), -(-T // 2560)
        return (torch.from_numpy(rs.randint(-1, 42, size=(B, tp)).astype(np.int64)),
                torch.from_numpy(rs.randint(-1, 10000, size=(B, tw)).astype(np.int64)))
    idx = rs.randint(1, n_labels - 1, size=(B, U))
    idx[:, 0] = 0
    idx[:, -1] = n_labels - 1
    return (F.one_hot(torch.from_numpy(idx.astype(np.int64)), n_labels).float(),)


def train_steps(params, B, T, steps, warmup, device="cpu", loop80=True, seed=1234, kind="slu"):
    """Adam train steps of the port on `device`; returns seconds per step.  CPU: wall clock.  GPU: CUDA events around the
    whole timed run (synchronised on both sides), which includes the reference's per-step host work and result read."""
    import time
    from torch.optim.adam import Adam as StockAdam      # the library class itself, whatever `torch.optim.Adam` is bound to
    model = RefPort(params, loop80=loop80, kind=kind).to(device).train()
    opt = StockAdam(model.parameters(), lr=1e-3)
    x, _ = R.synthetic_batch(B, T, seed=seed)
    ys = synthetic_labels(kind, B, T, seed)
    x = x.to(device)
    ys = tuple(y.to(device) for y in ys)

    def one():
        loss = model(x, *ys)
        opt.zero_grad()
        loss.backward()
        opt.step()
        loss.item()

    for _ in range(warmup):
        one()
    if device == "cpu":
        t0 = time.perf_counter()
        for _ in range(steps):
            one()
        return (time.perf_counter() - t0) / steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        one()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / steps
