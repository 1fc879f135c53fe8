"""Developer tool: warm per-kernel timings at the shapes of the B=256 train step (CUDA events, median of N repetitions, inputs
rotated over buffers larger than L2 where it matters).  One short GPU call answers "did this kernel change help?" without a
full bench.py run.    python tools/kbench.py [reps]"""
import importlib, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

pkg = importlib.import_module("end-to-end-slu_b200")
ops = pkg.ops
dev = "cuda"
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B = 256
torch.manual_seed(0)


def timeit(name, fn, bytes_moved=None):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    med = statistics.median(ts)
    extra = "  %.2f TB/s" % (bytes_moved / med / 1e6) if bytes_moved else ""
    print("%-44s %8.1f us (min %.1f)%s" % (name, med, min(ts), extra))


# ---- SincConv ---------------------------------------------------------------------------------------------------------------
x = 0.1 * torch.randn(B, 64000, device=dev)
b1 = torch.rand(80, dtype=torch.float64, device=dev) * 0.2 + 0.01
band = torch.rand(80, dtype=torch.float64, device=dev) * 0.05 + 0.005
timeit("sinc front end forward (filters+bank+conv)", lambda: ops.SincFrontend.apply(x, b1, band), B * (4 * 64000 + 4 * 80 * 400))

# ---- tap-GEMM ---------------------------------------------------------------------------------------------------------------
T_l, I_l = (400, 200, 100, 50, 25), (60, 256, 256, 256, 256)
for T, I in zip(T_l, I_l):
    xa = torch.randn(B * T, I, device=dev)
    w = torch.randn(768, I, device=dev) * 0.05
    bias = torch.zeros(768, device=dev)
    img = ops.presplit(w, *ops._form_nt(w))
    timeit("x-projection  M=%6d K=%3d N=768" % (B * T, I), lambda: ops.linear_nt(xa, w, bias, img), 4 * B * T * (I + 768))
    dgx = torch.randn(B * T, 768, device=dev)
    img2 = ops.presplit(w, *ops._form_nn(w))
    timeit("dX            M=%6d K=768 N=%3d" % (B * T, I), lambda: ops.matmul_nn(dgx, w, img2), 4 * B * T * (I + 768))
for cin, cout in ((80, 60), (60, 60)):
    xa = torch.randn(B, 400, cin, device=dev)
    w = torch.randn(cout, cin, 5, device=dev) * 0.05
    bias = torch.zeros(cout, device=dev)
    timeit("conv block forward %d->%d, 5 taps" % (cin, cout), lambda: ops.conv_block(xa, w, bias, 0.2), 4 * B * 400 * (cin + cout))

# ---- weight gradients ---------------------------------------------------------------------------------------------------------
for T, I in zip(T_l, I_l):
    dgx = torch.randn(B, T, 768, device=dev)
    dhn = torch.randn(B, T, 256, device=dev)
    xa = torch.randn(B, T, I, device=dev)
    y = torch.randn(B, T, 256, device=dev)
    o1 = torch.zeros(768, I, device=dev)
    o2 = torch.zeros(2, 384, 128, device=dev)
    timeit("dW_ih  T=%3d I=%3d" % (T, I), lambda: ops.wgrad_tc(dgx, 0, 768, 768, xa, 0, I, I, B, T, o1, 0, I), 4 * B * T * (768 + I))
    timeit("dW_hh  T=%3d (one direction)" % T,
           lambda: ops.wgrad2_tc(dgx, 0, 768, 256, dhn, 0, 256, 384, y, 0, 256, 128, B, T, o2, 0, 128, shift0=-1), 4 * B * T * 512)

# ---- recurrence ---------------------------------------------------------------------------------------------------------------
for T, I in zip(T_l, I_l):
    gru = torch.nn.GRU(I, 128, batch_first=True, bidirectional=True).cuda()
    xa = torch.randn(B, T, I, device=dev, requires_grad=True)
    mask = (torch.rand(B, T, 256, device=dev) > 0.5).float() * 2
    out = {}

    def fwd():
        out["y"] = ops.bigru(xa, gru, mask, 2 if T > 25 else 1)
    timeit("bi-GRU layer forward (x-proj + recurrence) T=%3d" % T, fwd)
    gy = torch.randn_like(out["y"])
    timeit("bi-GRU layer backward (all launches)      T=%3d" % T, lambda: torch.autograd.grad(out["y"], [xa] + list(gru.parameters()), gy, retain_graph=True))
print("done")
