"""Developer tool: the layer-0 x-projection GEMM (M=102400, K=60, N=768) a few times, for ncu."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("end-to-end-slu_b200")
ops = pkg.ops
M, K, N = 102400, 60, 768
x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") / 8; b = torch.randn(N, device="cuda")
out = torch.empty(M, N, device="cuda")
img = ops.presplit(w, *ops._form_nt(w))
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    ops.gemm_tc(x, K, img, M, N, K, out, bias=b)
torch.cuda.synchronize()
print("ok")
