"""Per-parameter gradient error of the CUDA path against oracle autograd at full utterance length (B=4 x 4 s), for the kernel
variants selectable at run time.  Usage: python tools/grad_error_table.py [T] [B]   (prints a table; developer tool)."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import models  # noqa: E402
from oracle import torch_ref as R  # noqa: E402
from util import make_config, rel_err  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 64000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
pkg = importlib.import_module("end-to-end-slu_b200")
p = R.synthetic_params(seed=4)
x, y = R.synthetic_batch(256, T, seed=5)
x, y = x[:B], y[:B]
pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
l_ref, _, lg_ref = R.slu_forward(x, y, pr)
l_ref.backward()


def run(tag):
    m = models.Model(make_config()).eval()
    sd = m.state_dict(); sd.update({k: v for k, v in p.items() if k in sd}); m.load_state_dict(sd)
    for q in m.parameters():
        q.requires_grad = True
    loss, _ = m(x, y)
    loss.backward()
    errs = {k: rel_err(q.grad.cpu(), pr[k].grad) for k, q in m.named_parameters() if q.grad is not None}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
    print("%-18s loss err %.2e | " % (tag, abs(loss.item() - l_ref.item()) / abs(l_ref.item())) +
          "  ".join("%s %.1e" % (k.replace("pretrained_model.", "").replace("_layers", ""), e) for k, e in worst))
    return errs


run("default")
pkg.ops.set_gru_precision("bf16x3-separate"); run("gru 3-pass"); pkg.ops.set_gru_precision("bf16x3")
pkg.ops.GRU_IMPL = "simt"; run("gru simt"); pkg.ops.GRU_IMPL = "tc"
pkg.ops.SINC_IMPL = "simt"; run("sinc simt")
pkg.ops.GRU_IMPL = "simt"; run("gru+sinc simt")
