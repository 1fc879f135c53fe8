"""Developer tool: the weight-gradient launches of one train step (config 3, B=256 x 4 s) one by one, CUDA-event timed with an
L2 flush in between, optionally with pipeline stages switched off (slu_debug_wgrad_mode) to see which stage bounds the kernel.
  python tools/wgrad_only.py [modes, e.g. 0,1,3,7] [reps]
Needs a library built with the debug switches: SLU_KERNEL_DEBUG=1 python __graft_entry__.py (the default build has none)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("end-to-end-slu_b200")
ops, _lib = pkg.ops, pkg._lib
lib = _lib.load()
modes = [int(m) for m in (sys.argv[1] if len(sys.argv) > 1 else "0").split(",")]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
B, H = 256, 128
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
shapes = []          # (name, fn)
for li, (T, I) in enumerate([(400, 60), (200, 256), (100, 256), (50, 256), (25, 256)]):
    dgx = torch.randn(B, T, 768, device=dev); dhn = torch.randn(B, T, 256, device=dev)
    y = torch.randn(B, T, 256, device=dev); x = torch.randn(B, T, I, device=dev)
    dw_ih = torch.zeros(768, I, device=dev); dw_hh = torch.zeros(2, 384, H, device=dev)
    shapes.append((f"L{li} dW_ih  768x{I:3d} T={T}", (768 * 4 + I * 4) * B * T,
                   lambda dgx=dgx, x=x, I=I, T=T, dw_ih=dw_ih: ops.wgrad_tc(dgx, 0, 768, 768, x, 0, I, I, B, T, dw_ih, 0, I)))
    shapes.append((f"L{li} dW_hh0 384x128 T={T}", (384 + 128) * 4 * B * T,
                   lambda dgx=dgx, dhn=dhn, y=y, T=T, dw_hh=dw_hh: ops.wgrad2_tc(dgx, 0, 768, 256, dhn, 0, 256, 384, y, 0, 256, H, B, T, dw_hh, 0, H, shift0=-1)))
if os.environ.get("WGRAD_EXTRA"):
    T = 400
    gd = torch.randn(B, T, 384, device=dev); xd = torch.randn(B, T, 128, device=dev); dwd = torch.zeros(384, 128, device=dev)
    shapes.append(("dense 384x128 T=400 (ld=width)", (384 + 128) * 4 * B * T,
                   lambda: ops.wgrad_tc(gd, 0, 384, 384, xd, 0, 128, 128, B, T, dwd, 0, 128, shift0=-1)))
    g0d = torch.randn(B, T, 256, device=dev); g1d = torch.randn(B, T, 128, device=dev)
    shapes.append(("L0 dW_hh, dense sources (SoA emu)", (384 + 128) * 4 * B * T,
                   lambda: ops.wgrad2_tc(g0d, 0, 256, 256, g1d, 0, 128, 384, xd, 0, 128, H, B, T, dwd, 0, H, shift0=-1)))
    for (TT, II) in [(400, 60), (200, 256)]:
        gq = torch.randn(B, TT, 384 if II == 60 else 256, device=dev); xq = torch.randn(B, TT, II, device=dev)
        Mq = gq.shape[2]; dwq = torch.zeros(Mq, II, device=dev)
        shapes.append((f"dW_ih one m-group dense {Mq}x{II} T={TT}", (Mq + II) * 4 * B * TT,
                       lambda gq=gq, xq=xq, Mq=Mq, II=II, TT=TT, dwq=dwq: ops.wgrad_tc(gq, 0, Mq, Mq, xq, 0, II, II, B, TT, dwq, 0, II)))
    dgx2 = torch.randn(B, T, 768, device=dev); dhn2 = torch.randn(B, T, 256, device=dev); y2 = torch.randn(B, T, 256, device=dev)
    dw2 = torch.zeros(2, 384, H, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def both():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur); s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            ops.wgrad2_tc(dgx2, 0, 768, 256, dhn2, 0, 256, 384, y2, 0, 256, H, B, T, dw2, 0, H, shift0=-1, stream=s1.cuda_stream)
        with torch.cuda.stream(s2):
            ops.wgrad2_tc(dgx2, 384, 768, 256, dhn2, H, 256, 384, y2, H, 256, H, B, T, dw2, 384 * H, H, shift0=1, stream=s2.cuda_stream)
        cur.wait_stream(s1); cur.wait_stream(s2)
    shapes.append(("L0 dW_hh both dirs, 2 streams", 2 * (384 + 128) * 4 * B * T, both))
for (Cout, Cin, T) in [(60, 80, 400), (60, 60, 200)]:
    k = 5
    dpre = torch.randn(B, T, Cout, device=dev); x = torch.randn(B, T, Cin, device=dev); dw = torch.zeros(Cout, Cin * k, device=dev)
    shapes.append((f"conv dW {Cout}x{Cin}x5 T={T}", (Cout + Cin) * 4 * B * T,
                   lambda dpre=dpre, x=x, Cout=Cout, Cin=Cin, T=T, dw=dw: ops.wgrad_tc(dpre, 0, Cout, Cout, x, 0, Cin, Cin, B, T, dw, 0, Cin * k, k, 1, taps=k, shift0=-2)))
res = {}
for mode in modes:
    lib.slu_debug_wgrad_mode(mode)
    for name, nbytes, fn in shapes:
        fn(); torch.cuda.synchronize()
        ts = []
        for r in range(reps):
            flush.fill_(r)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        res[(mode, name)] = (sorted(ts)[len(ts) // 2], nbytes)
lib.slu_debug_wgrad_mode(0)
print(f"{'launch':34s} {'MB':>7s} {'floor us':>8s} " + " ".join(f"mode{m:>2d} us" for m in modes))
tot = {m: 0.0 for m in modes}
for name, nbytes, fn in shapes:
    row = f"{name:34s} {nbytes / 1e6:7.1f} {nbytes / 6.56e6:8.1f} "
    for m in modes:
        row += f"{res[(m, name)][0]:9.1f} "
        tot[m] += res[(m, name)][0] * (2 if "dW_hh0" in name else 1)
    print(row)
print("step total (dW_hh x2): " + "  ".join(f"mode {m}: {tot[m]:.1f} us" for m in modes))
