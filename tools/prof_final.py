"""Developer tool: the dominant launches of the B=256 train step, one each, for `ncu --set full`:
  ncu --set full --import-source on --clock-control none -k regex:"gru_fwd_tc|sincconv_fwd_tc|wgrad_tc_kernel|gemm_tc_kernel" \
      --launch-skip <warm-up launches> -o gpurun_out/final python tools/prof_final.py"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("end-to-end-slu_b200")
ops = pkg.ops
dev = "cuda"
torch.manual_seed(0)
B = 256
x = 0.1 * torch.randn(B, 64000, device=dev)
b1 = torch.rand(80, dtype=torch.float64, device=dev) * 0.2 + 0.01
band = torch.rand(80, dtype=torch.float64, device=dev) * 0.05 + 0.005
gru0 = torch.nn.GRU(60, 128, batch_first=True, bidirectional=True).cuda()
x0 = torch.randn(B, 400, 60, device=dev)
mask = (torch.rand(B, 400, 256, device=dev) > 0.5).float() * 2
dgx1 = torch.randn(B, 200, 768, device=dev)
dhn1 = torch.randn(B, 200, 256, device=dev)
x1 = torch.randn(B, 200, 256, device=dev)
y1 = torch.randn(B, 200, 256, device=dev)
out_ih = torch.zeros(768, 256, device=dev)
out_hh = torch.zeros(2, 384, 128, device=dev)
for _ in range(2):      # launch order per round: sinc fwd, gemm (x-proj L0), gru fwd L0, wgrad dW_ih L1, wgrad2 dW_hh L1
    ops.SincFrontend.apply(x, b1, band)
    with torch.no_grad():
        xx = x0.clone().requires_grad_(True)
    y = ops.bigru(xx, gru0, mask, 2)
    ops.wgrad_tc(dgx1, 0, 768, 768, x1, 0, 256, 256, B, 200, out_ih, 0, 256)
    ops.wgrad2_tc(dgx1, 0, 768, 256, dhn1, 0, 256, 384, y1, 0, 256, 128, B, 200, out_hh, 0, 128, shift0=-1)
torch.cuda.synchronize()
print("done")
