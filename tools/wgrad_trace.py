"""Developer tool: clock64() hand-off trace of CTA 0 of one weight-gradient launch (slu_debug_wgrad_trace).
  python tools/wgrad_trace.py kind[,kind..] [layer 0..4] [mode] [-v]
kinds: hh (dW_hh as BiGRU.backward launches it), hh_dense (same from dense sources), ih (dW_ih 768 x I), ih_1g (one dense m-group),
       conv0 / conv1 (Conv1d weight gradients)
Needs a library built with the debug switches: SLU_KERNEL_DEBUG=1 python __graft_entry__.py (the default build has none)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("end-to-end-slu_b200")
ops, _lib = pkg.ops, pkg._lib
lib = _lib.load()
kinds = sys.argv[1].split(",") if len(sys.argv) > 1 else ["hh"]
li = int(sys.argv[2]) if len(sys.argv) > 2 else 0
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
verbose = "-v" in sys.argv
B, H = 256, 128
T, I = [(400, 60), (200, 256), (100, 256), (50, 256), (25, 256)][li]
dev = "cuda"
r = lambda *s: torch.randn(*s, device=dev)
z = lambda *s: torch.zeros(*s, device=dev)


def make(kind):
    if kind == "hh":
        dgx, dhn, y, dw = r(B, T, 768), r(B, T, 256), r(B, T, 256), z(2, 384, H)
        return lambda: ops.wgrad2_tc(dgx, 0, 768, 256, dhn, 0, 256, 384, y, 0, 256, H, B, T, dw, 0, H, shift0=-1), (384 + 128) * 4 * B * T
    if kind == "hh_dense":
        g0, g1, y, dw = r(B, T, 256), r(B, T, 128), r(B, T, 128), z(384, H)
        return lambda: ops.wgrad2_tc(g0, 0, 256, 256, g1, 0, 128, 384, y, 0, 128, H, B, T, dw, 0, H, shift0=-1), (384 + 128) * 4 * B * T
    if kind == "ih":
        dgx, x, dw = r(B, T, 768), r(B, T, I), z(768, I)
        return lambda: ops.wgrad_tc(dgx, 0, 768, 768, x, 0, I, I, B, T, dw, 0, I), (768 + I * (2 if I == 60 else 3)) * 4 * B * T
    if kind == "ih_1g":
        M = 384 if I == 60 else 256
        g, x, dw = r(B, T, M), r(B, T, I), z(M, I)
        return lambda: ops.wgrad_tc(g, 0, M, M, x, 0, I, I, B, T, dw, 0, I), (M + I) * 4 * B * T
    if kind in ("conv0", "conv1"):
        Cout, Cin, Tc = (60, 80, 400) if kind == "conv0" else (60, 60, 200)
        g, x, dw = r(B, Tc, Cout), r(B, Tc, Cin), z(Cout, Cin * 5)
        return lambda: ops.wgrad_tc(g, 0, Cout, Cout, x, 0, Cin, Cin, B, Tc, dw, 0, Cin * 5, 5, 1, taps=5, shift0=-2), (Cout + Cin) * 4 * B * Tc
    raise SystemExit("unknown kind " + kind)


flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for kind in kinds:
    fn, nbytes = make(kind)
    fn(); torch.cuda.synchronize()
    buf = torch.zeros(64, 8, dtype=torch.int64, device=dev)
    flush.fill_(1)
    lib.slu_debug_wgrad_mode(mode); lib.slu_debug_wgrad_trace(buf.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    lib.slu_debug_wgrad_trace(None); lib.slu_debug_wgrad_mode(0)
    t = buf.cpu()
    n = int((t[:, 0] != 0).sum())
    t0 = int(t[0, 0])
    rows = [[int(v) - t0 for v in t[i, :6]] for i in range(n)]
    if verbose:
        print("tile  issue0  issue1  full  conv_done  mma_see  mma_issued | load latency (full - issue1)")
        for i, rr in enumerate(rows):
            print(f"{i:4d} " + " ".join(f"{v:8d}" for v in rr) + f" | {rr[2] - rr[1]:6d}")
    a = min(8, n // 3)
    per_tile = (rows[n - 1][2] - rows[a][2]) / max(1, n - 1 - a)
    lat = sum(rr[2] - rr[1] for rr in rows[a:]) / max(1, n - a)
    conv = sum(rr[3] - rr[2] for rr in rows[a:]) / max(1, n - a)
    mma = sum(rr[5] - rr[4] for rr in rows[a:]) / max(1, n - a)
    print(f"{kind:9s} L{li} mode {mode}: {e0.elapsed_time(e1) * 1e3:6.1f} us for {nbytes / 1e6:6.1f} MB | CTA 0, tiles {a}..{n - 1}: {per_tile:6.0f} cycles/tile, "
          f"first full at {rows[0][2]}, load latency {lat:5.0f}, convert {conv:5.0f}, mma issue {mma:4.0f} cycles")
