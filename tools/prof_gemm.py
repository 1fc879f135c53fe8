"""Developer tool: a few representative tap-GEMM / weight-gradient launches of the B=256 train step, for `ncu --set full`.
  ncu --set full --import-source on --clock-control none -k regex:"gemm_tc_kernel|wgrad_tc_kernel" -o gpurun_out/gemm python tools/prof_gemm.py"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("end-to-end-slu_b200")
ops = pkg.ops
dev = "cuda"
torch.manual_seed(0)
B = 256
x80 = torch.randn(B, 400, 80, device=dev)
w1 = torch.randn(60, 80, 5, device=dev) * 0.05
b1 = torch.zeros(60, device=dev)
x256 = torch.randn(B * 25, 256, device=dev)
w_ih = torch.randn(768, 256, device=dev) * 0.05
bias = torch.zeros(768, device=dev)
dgx0 = torch.randn(B * 400, 768, device=dev)
w_ih0 = torch.randn(768, 60, device=dev) * 0.05
dgx1 = torch.randn(B, 200, 768, device=dev)
x1 = torch.randn(B, 200, 256, device=dev)
out = torch.zeros(768, 256, device=dev)
for _ in range(2):
    ops.conv_block(x80, w1, b1, 0.2)                      # conv1 forward: K=80, 5 taps, N=60
    ops.linear_nt(x256, w_ih, bias)                       # x-projection of the intent layer: M=6400, K=256, N=768
    ops.matmul_nn(dgx0, w_ih0)                            # dX of GRU layer 0: M=102400, K=768, N=60
    ops.wgrad_tc(dgx1, 0, 768, 768, x1, 0, 256, 256, B, 200, out, 0, 256)      # dW_ih of GRU layer 1
torch.cuda.synchronize()
print("done")
