"""Timing port of the reference's execution structure.  TEST/BENCH INFRASTRUCTURE ONLY
(bench.py's cpu_baseline / --impl reference legs; never on the product path).

Same math as oracle/torch_ref.py, but built from the library modules the reference itself uses
(nn.GRU, F.conv1d, nn.Dropout, ...) and -- with loop80=True -- bug-compatible with the reference's
SincLayer.forward, which re-runs the full conv1d inside its 80-iteration filter loop
(models.py:98-108: 80 convolutions + ~25 small ops per filter per forward).  That is what
"the reference's own implementation" costs, on CPU (mkldnn + aten::gru) or on GPU (cuDNN).
"""
import math

import torch
import torch.nn.functional as F

from . import torch_ref as R


class RefPort(torch.nn.Module):
    def __init__(self, params, loop80=True, drop_p=0.5):
        super().__init__()
        P = R.P
        self.filt_b1 = torch.nn.Parameter(params[P + "phoneme_layers.0.filt_b1"].clone())
        self.filt_band = torch.nn.Parameter(params[P + "phoneme_layers.0.filt_band"].clone())
        self.conv1 = torch.nn.Conv1d(80, 60, 5, padding=2)
        self.conv2 = torch.nn.Conv1d(60, 60, 5, padding=2)
        self.grus = torch.nn.ModuleList([torch.nn.GRU(I, 128, batch_first=True, bidirectional=True)
                                         for I in (60, 256, 256, 256, 256)])
        self.final = torch.nn.Linear(256, 24)
        self.drop = torch.nn.Dropout(drop_p)
        self.loop80 = loop80
        with torch.no_grad():
            self.conv1.weight.copy_(params[P + "phoneme_layers.5.weight"]); self.conv1.bias.copy_(params[P + "phoneme_layers.5.bias"])
            self.conv2.weight.copy_(params[P + "phoneme_layers.9.weight"]); self.conv2.bias.copy_(params[P + "phoneme_layers.9.bias"])
            for gru, key in zip(self.grus, R.GRU_KEYS):
                for n, q in gru.named_parameters():
                    q.copy_(params[key + "." + n])
            self.final.weight.copy_(params["intent_layers.4.weight"]); self.final.bias.copy_(params["intent_layers.4.bias"])

    def sinc(self, x):
        dev = x.device
        fs, N = 16000, 401
        if not self.loop80:
            w = R.sinc_filters(self.filt_b1.cpu(), self.filt_band.cpu()).to(dev) if dev.type != "cpu" else R.sinc_filters(self.filt_b1, self.filt_band)
            return F.conv1d(x, w.unsqueeze(1), stride=80, padding=200)
        # structure of models.py:77-110, per-filter python loop with the conv inside it
        filters = torch.zeros((80, N), device=dev)
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
                idx = torch.arange(y_right.shape[0] - 1, -1, -1).to(dev)
                return 2 * f * torch.cat([y_right[idx], torch.ones(1).to(dev), y_right])
            bp = lp(end[i]) - lp(beg[i])
            bp = bp / torch.max(bp)
            filters[i, :] = bp * window
            out = F.conv1d(x, filters.view(80, 1, N), stride=80, padding=200)
        return out

    def forward(self, x, y):
        out = self.sinc(x.unsqueeze(1))
        out = F.leaky_relu(F.max_pool1d(torch.abs(out), 2, ceil_mode=True), 0.2)
        out = F.leaky_relu(self.conv1(out), 0.2)
        out = F.leaky_relu(self.conv2(out), 0.2)
        out = out.transpose(1, 2)
        for li, gru in enumerate(self.grus):
            out = self.drop(gru(out)[0])
            out = R.downsample(out, *R.GRU_DOWNSAMPLE[li])
        logits = self.final(out).max(dim=1)[0]
        loss, acc, _ = R.intent_loss_acc(logits, y)
        return loss, acc, logits


def train_steps(params, B, T, steps, warmup, device="cpu", loop80=True, seed=1234):
    """Adam train steps of the port on `device`; returns seconds per step (wall clock, synchronised)."""
    import time
    model = RefPort(params, loop80=loop80).to(device).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    x, y = R.synthetic_batch(B, T, seed=seed)
    x, y = x.to(device), y.to(device)
    times = []
    for i in range(warmup + steps):
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        loss, _, _ = model(x, y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        loss.item()
        if device != "cpu":
            torch.cuda.synchronize()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)
