"""Developer tool: sustained time per weight-gradient launch, measured over `reps` back-to-back launches (no host gaps, data sets
larger than L2), per pipeline-ablation mode (slu_debug_wgrad_mode).   python tools/wgrad_rate.py [modes] [reps]
Needs a library built with the debug switches: SLU_KERNEL_DEBUG=1 python __graft_entry__.py (the default build has none)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("end-to-end-slu_b200")
ops, _lib = pkg.ops, pkg._lib
lib = _lib.load()
modes = [int(m) for m in (sys.argv[1] if len(sys.argv) > 1 else "0,7").split(",")]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
H, dev = 128, "cuda"
r = lambda *s: torch.randn(*s, device=dev)
z = lambda *s: torch.zeros(*s, device=dev)
cases = []
for (name, B, T, I) in [("L0", 256, 400, 60), ("L1", 512, 200, 256), ("L3", 2048, 50, 256)]:
    dgx, dhn, y, x = r(B, T, 768), r(B, T, 256), r(B, T, 256), r(B, T, I)
    dwh, dwi = z(2, 384, H), z(768, I)
    g0, g1, yd = r(B, T, 256), r(B, T, 128), r(B, T, 128)
    M1 = 384 if I == 60 else 256
    gq, dwq = r(B, T, M1), z(M1, I)
    cases += [
        (f"{name} dW_hh strided  B={B} T={T}", (384 + 128) * 4 * B * T,
         lambda dgx=dgx, dhn=dhn, y=y, dwh=dwh, B=B, T=T: ops.wgrad2_tc(dgx, 0, 768, 256, dhn, 0, 256, 384, y, 0, 256, H, B, T, dwh, 0, H, shift0=-1)),
        (f"{name} dW_hh dense    B={B} T={T}", (384 + 128) * 4 * B * T,
         lambda g0=g0, g1=g1, yd=yd, dwh=dwh, B=B, T=T: ops.wgrad2_tc(g0, 0, 256, 256, g1, 0, 128, 384, yd, 0, 128, H, B, T, dwh, 0, H, shift0=-1)),
        (f"{name} dW_ih 768x{I}   B={B} T={T}", (768 + I * (2 if I == 60 else 3)) * 4 * B * T,
         lambda dgx=dgx, x=x, dwi=dwi, B=B, T=T, I=I: ops.wgrad_tc(dgx, 0, 768, 768, x, 0, I, I, B, T, dwi, 0, I)),
        (f"{name} dW_ih 1 group {M1}x{I}", (M1 + I) * 4 * B * T,
         lambda gq=gq, x=x, dwq=dwq, B=B, T=T, I=I, M1=M1: ops.wgrad_tc(gq, 0, M1, M1, x, 0, I, I, B, T, dwq, 0, I)),
    ]
B = 256
for (Cout, Cin, T) in [(60, 80, 400), (60, 60, 200)]:
    Bc = 256 * (400 // T) * 4
    g, x, dw = r(Bc, T, Cout), r(Bc, T, Cin), z(Cout, Cin * 5)
    cases.append((f"conv dW {Cout}x{Cin}x5 B={Bc} T={T}", (Cout + Cin) * 4 * Bc * T,
                  lambda g=g, x=x, dw=dw, Cout=Cout, Cin=Cin, T=T, Bc=Bc: ops.wgrad_tc(g, 0, Cout, Cout, x, 0, Cin, Cin, Bc, T, dw, 0, Cin * 5, 5, 1, taps=5, shift0=-2)))
print(f"{'launch':40s} {'MB':>7s} " + " ".join(f"mode{m} us  TB/s |" for m in modes))
for name, nbytes, fn in cases:
    row = f"{name:40s} {nbytes / 1e6:7.1f} "
    for m in modes:
        lib.slu_debug_wgrad_mode(m)
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        row += f"{us:8.1f} {nbytes / us / 1e6:5.2f} |"
    lib.slu_debug_wgrad_mode(0)
    print(row)
