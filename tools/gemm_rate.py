"""Developer tool: sustained time per slu_gemm_tc launch for the x-projection / input-gradient shapes of the config-3 step, over
`reps` back-to-back launches, per pipeline-ablation mode (slu_debug_gemm_mode).   python tools/gemm_rate.py [modes] [reps]
Needs a library built with the debug switches: SLU_KERNEL_DEBUG=1 python __graft_entry__.py (the default build has none)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("end-to-end-slu_b200")
ops, _lib = pkg.ops, pkg._lib
lib = _lib.load()
modes = [int(m) for m in (sys.argv[1] if len(sys.argv) > 1 else "0,4,8,1,2").split(",")]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = "cuda"
r = lambda *s: torch.randn(*s, device=dev)
cases = []
for (name, M, K, N) in [("x-proj L0", 102400, 60, 768), ("x-proj L1", 2 * 51200, 256, 768), ("dX L0 (768->60)", 102400, 768, 60),
                        ("dX L1 (768->256)", 2 * 51200, 768, 256)]:
    x = r(M, K); w = r(N, K) / 8; b = r(N); out = torch.empty(M, N, device=dev)
    img = ops.presplit(w, *ops._form_nt(w))
    cases.append((f"{name:18s} M={M} K={K} N={N}", (M * K + M * N) * 4, lambda x=x, img=img, M=M, N=N, K=K, out=out, b=b: ops.gemm_tc(x, K, img, M, N, K, out, bias=b)))
print(f"{'launch':46s} {'MB':>7s} " + " ".join(f"mode{m:<2d} us  TB/s |" for m in modes))
for name, nbytes, fn in cases:
    row = f"{name:46s} {nbytes / 1e6:7.1f} "
    for m in modes:
        lib.slu_debug_gemm_mode(m)
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        row += f"{us:9.1f} {nbytes / us / 1e6:5.2f} |"
    lib.slu_debug_gemm_mode(0)
    print(row)
