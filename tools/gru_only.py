"""Developer tool: the layer-0 persistent-GRU forward + backward at the bench shape (B=256, T=400, I=60, dropout 0.5) for ncu."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("end-to-end-slu_b200")
ops = pkg.ops
B, T, I = 256, 400, 60
gru = torch.nn.GRU(I, 128, batch_first=True, bidirectional=True).cuda()
x = torch.randn(B, T, I, device="cuda", requires_grad=True)
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    y = ops.bigru(x, gru, (0.5, 1234 + it), 2)
    y.backward(torch.randn_like(y))
torch.cuda.synchronize()
print("ok")
