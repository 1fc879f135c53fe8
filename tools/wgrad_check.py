"""Developer tool: slu_wgrad_tc against torch for a few shapes (debug modes via argv[1]).
Needs a library built with the debug switches: SLU_KERNEL_DEBUG=1 python __graft_entry__.py (the default build has none)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("end-to-end-slu_b200")
ops, _lib = pkg.ops, pkg._lib
lib = _lib.load()
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
torch.manual_seed(0)
for (M, N, T) in [(256, 256, 2048), (1024, 256, 2048), (10000, 256, 2048), (10000, 256, 904), (768, 256, 5000), (384, 128, 6000)]:
    G = torch.randn(T, M, device="cuda"); X = torch.randn(T, N, device="cuda")
    ref = (G.double().t() @ X.double())
    for rep in range(3):
        out = torch.zeros(M, N, device="cuda")
        lib.slu_debug_wgrad_mode(mode)
        ops.wgrad_tc(G, 0, M, M, X, 0, N, N, 1, T, out, 0, N)
        torch.cuda.synchronize()
        lib.slu_debug_wgrad_mode(0)
        err = (out.double() - ref).abs()
        rowbad = (err.max(1)[0] > 1e-2 * ref.abs().max()).nonzero().flatten()
        print(f"mode {mode} M={M} N={N} T={T} rep {rep}: rel {float(err.norm() / ref.norm()):.2e}  bad rows {rowbad.numel()}"
              + (f" first {rowbad[:6].tolist()} last {rowbad[-3:].tolist()}" if rowbad.numel() else ""))
