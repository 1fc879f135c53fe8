"""One small SLU train step (B=3, 0.3 s) + the ASR and seq2seq paths through every kernel family, for compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitizer_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitizer_smoke.py
(persistent kernels with mbarriers / TMEM are race-prone; the summaries are kept under profiles/.)"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("SLU_STEP_GRAPH", "0")
import models  # noqa: E402
from oracle import torch_ref as R  # noqa: E402
from util import make_config  # noqa: E402

pkg = importlib.import_module("end-to-end-slu_b200")
torch.manual_seed(0)
m = models.Model(make_config()).train()
for q in m.parameters():
    q.requires_grad = True
opt = torch.optim.Adam(m.parameters(), lr=1e-3)
x, y = R.synthetic_batch(3, 4800, seed=1)
for _ in range(2):
    loss, acc = m(x, y); opt.zero_grad(); loss.backward(); opt.step()
print("slu step ok", loss.item())
pm = models.PretrainedModel(make_config(pretraining_type=2)).train()
xa, _ = R.synthetic_batch(2, 5120, seed=2)
yp = torch.randint(-1, 42, (2, 8)); yw = torch.randint(-1, 10000, (2, 2))
pl, wl, _, _ = pm(xa, yp, yw); (pl + wl).backward()
print("asr step ok", pl.item(), wl.item())
cfg = make_config("seq2seq"); cfg.Sy_intent = ["<sos>"] + list("abcdefghij") + ["<eos>"]
s = models.Model(cfg).train()
S = len(cfg.Sy_intent)
idx = torch.randint(1, S - 1, (3, 5)); idx[:, 0] = 0; idx[:, -1] = S - 1
ls, _ = s(x, torch.nn.functional.one_hot(idx, S).float()); ls.backward()
print("seq2seq step ok", ls.item())
torch.cuda.synchronize()
