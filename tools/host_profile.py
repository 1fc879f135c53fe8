"""Host-side cost of one train step: wall time to ISSUE a step (no sync inside) vs its device time, and a cProfile of the
issue path.  Run on the GPU box:  python tools/host_profile.py [B]"""
import cProfile
import importlib
import io
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import models  # noqa: E402

pkg = importlib.import_module("end-to-end-slu_b200")
cfg = pkg.config.read_config(os.path.join(ROOT, "configs", "unfreeze_all_layers.cfg"))
cfg.pretraining_type = 0
cfg.num_phonemes = 42
cfg.Sy_intent, cfg.values_per_slot = pkg.config.fsc_intent_table()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
m = models.Model(cfg).train()
opt = torch.optim.Adam(m.parameters(), lr=1e-3)
dev = "cuda" if torch.cuda.is_available() else "cpu"
T = 64000 if dev == "cuda" else 4000
x = (0.1 * torch.randn(B, T)).to(dev)
y = torch.stack([torch.randint(0, v, (B,)) for v in (6, 14, 4)], 1).to(dev)
sync = torch.cuda.synchronize if dev == "cuda" else (lambda: None)


wall = {}


def timed(name, fn):
    def w(*a, **k):
        t = time.perf_counter()
        r = fn(*a, **k)
        e = wall.setdefault(name, [0, 0.0])
        e[0] += 1
        e[1] += time.perf_counter() - t
        return r
    return w


if dev == "cuda":
    for cls in (pkg.ops.BiGRU, pkg.ops.ConvBlock, pkg.ops.SincFrontend):
        cls.backward = staticmethod(timed(cls.__name__ + ".backward", cls.backward))
        cls.forward = staticmethod(timed(cls.__name__ + ".forward", cls.forward))
    pkg._lib.call = timed("_lib.call", pkg._lib.call)


def step():
    loss, acc = m(x, y)
    opt.zero_grad()
    loss.backward()
    opt.step()
    return loss


def phases():
    t = [time.perf_counter()]
    loss, acc = m(x, y); t.append(time.perf_counter())
    opt.zero_grad(); t.append(time.perf_counter())
    loss.backward(); t.append(time.perf_counter())
    opt.step(); t.append(time.perf_counter())
    return [1e3 * (b - a) for a, b in zip(t, t[1:])]


for _ in range(5):
    step()
for overlap in (True, False):
    pkg.ops.OVERLAP = overlap
    for _ in range(3):
        step()
    sync()
    acc = [0.0] * 4
    n = 4
    t0 = time.perf_counter()
    for _ in range(n):
        acc = [a + b for a, b in zip(acc, phases())]
    t1 = time.perf_counter()
    sync()
    t2 = time.perf_counter()
    print({k: (v[0] // (n + 3), round(1e3 * v[1] / (n + 3), 3)) for k, v in wall.items()})
    wall.clear()
    print("overlap=%s  issue %.2f ms/step (fwd %.2f, zero_grad %.2f, bwd %.2f, adam %.2f); issue+drain %.2f ms/step"
          % (overlap, 1e3 * (t1 - t0) / n, *[a / n for a in acc], 1e3 * (t2 - t0) / n))
pkg.ops.OVERLAP = False
sync()
pr = cProfile.Profile()
pr.enable()
for _ in range(4):
    step()
pr.disable()
sync()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue())

# ---- where does the end-to-end step lose time against the device-timed one? --------------------------------------------------
if dev == "cuda":
    pkg.ops.OVERLAP = True
    xs = [(0.1 * torch.randn(B, T)).pin_memory() for _ in range(4)]
    ys = [torch.stack([torch.randint(0, v, (B,)) for v in (6, 14, 4)], 1).pin_memory() for _ in range(4)]

    def run(n, host, read):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync()
        e0.record()
        if host:
            for xb, yb in pkg.loader.DevicePrefetcher(((xs[i % 4], ys[i % 4]) for i in range(n))):
                loss, _ = m(xb, yb); opt.zero_grad(); loss.backward(); opt.step()
                if read:
                    loss.item()
        else:
            for i in range(n):
                loss, _ = m(x, y); opt.zero_grad(); loss.backward(); opt.step()
                if read:
                    loss.item()
        e1.record()
        sync()
        return e0.elapsed_time(e1) / n

    for host, read in ((False, False), (False, True), (True, False), (True, True)):
        run(3, host, read)
        print("inputs %-6s loss.item() per step %-5s : %.3f ms/step" % ("host" if host else "device", read, run(20, host, read)))
