"""Developer tool: the SincConv forward + Jacobian backward at the bench shape (256 x 4 s), for `ncu` captures and quick timings."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("end-to-end-slu_b200")
ops = pkg.ops
B, T = 256, 64000
x = 0.1 * torch.randn(B, T, device="cuda")
b1 = (torch.rand(80, dtype=torch.float64, device="cuda") * 0.2 + 0.01).requires_grad_(True)
band = (torch.rand(80, dtype=torch.float64, device="cuda") * 0.05 + 0.005).requires_grad_(True)
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    out = ops.SincFrontend.apply(x, b1, band)
    out.backward(torch.randn_like(out))
torch.cuda.synchronize()
print("ok")
