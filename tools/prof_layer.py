"""One forward+backward of a single BiGRU layer (and optionally the conv block) at benchmark size,
for `ncu -k regex:...` captures.  Usage: python tools/prof_layer.py [B] [T] [I]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("end-to-end-slu_b200")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 400
I = int(sys.argv[3]) if len(sys.argv) > 3 else 60
torch.manual_seed(0)
gru = torch.nn.GRU(I, 128, batch_first=True, bidirectional=True).cuda()
x = torch.randn(B, T, I, device="cuda", requires_grad=True)
mask = (torch.rand(B, T, 256, device="cuda") > 0.5).float() * 2
for it in range(2):
    y = pkg.ops.bigru(x, gru, mask, 2)
    y.sum().backward()
torch.cuda.synchronize()
print("ok", y.shape, pkg.ops.GRU_IMPL)
