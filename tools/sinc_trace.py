"""Developer tool: hand-off timeline of CTA 0 of the persistent SincConv kernel (cycles since its first event).
Needs a library built with the debug switches: SLU_KERNEL_DEBUG=1 python __graft_entry__.py (the default build has none)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("end-to-end-slu_b200")
ops, lib = pkg.ops, pkg._lib
B, T = 256, 64000
x = 0.1 * torch.randn(B, T, device="cuda")
b1 = torch.rand(80, dtype=torch.float64, device="cuda") * 0.2 + 0.01
band = torch.rand(80, dtype=torch.float64, device="cuda") * 0.05 + 0.005
for _ in range(3):
    ops.SincFrontend.apply(x, b1, band)
buf = torch.zeros(16, 8, dtype=torch.int64, device="cuda")
lib.call("slu_debug_sinc_trace", buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.SincFrontend.apply(x, b1, band); e1.record()
torch.cuda.synchronize()
lib.call("slu_debug_sinc_trace", None)
t = buf.cpu()
t0 = int(t[t > 0].min())
names = ["stg:slot free", "stg:committed", "mma:img ready", "mma:acc free", "mma:issued", "epi:acc full", "epi:drained", "bank:tap5"]
print("front end %.1f us" % (e0.elapsed_time(e1) * 1e3))
print("tile " + " ".join("%14s" % n for n in names))
for i in range(16):
    if t[i].max() > 0:
        print("%4d " % i + " ".join("%14d" % (int(v) - t0 if v > 0 else -1) for v in t[i]))
