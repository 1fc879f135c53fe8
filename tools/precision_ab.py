"""Precision A/B of the tcgen05 recurrence: for every operand mode of slu_gru_{fwd,bwd}_tc --
  bf16x3 (default: bf16 hi/lo split, hi/lo rows stacked along N on the small tiles), bf16x3-separate (three passes), fp16 (one pass) --
report the errors against ALL reference goldens (test.wav known answer with the trained checkpoint; three synthetic cases with
loss, logits and all 48 gradients) next to the measured train-step throughput (config 3, 256 x 4 s, device-timed).
The bar is 1e-3 relative on the intent logits against the CPU fp32 reference; the reference's own cuDNN path allows TF32.
    python tools/precision_ab.py > profiles/r2_precision_ab.json
"""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import models  # noqa: E402
from oracle import torch_ref as R  # noqa: E402
from util import ckpt_params, golden, load_test_wav, make_config, rel_err  # noqa: E402

pkg = importlib.import_module("end-to-end-slu_b200")


def model_with(params, train=False):
    m = models.Model(make_config())
    sd = m.state_dict(); sd.update({k: v for k, v in params.items() if k in sd}); m.load_state_dict(sd)
    return m.train() if train else m.eval()


def goldens():
    out = {}
    g = golden("golden_testwav.npz")
    m = model_with(ckpt_params())
    with torch.no_grad():
        logits, pred = m.predict_intents(load_test_wav())
    out["testwav_logits_rel_err"] = rel_err(logits.cpu(), g["logits"])
    out["testwav_pred_ok"] = pred.tolist() == [[1, 2, 1]]
    worst_l, worst_g = 0.0, 0.0
    for tag in ("small", "ragged", "odd"):
        g = golden("golden_synth_%s.npz" % tag)
        p = R.synthetic_params(seed=int(g["pseed"]))
        m = model_with(p)
        for q in m.parameters():
            q.requires_grad = True
        x, y = R.synthetic_batch(int(g["B"]), int(g["T"]), seed=int(g["bseed"]))
        loss, _ = m(x, y)
        loss.backward()
        with torch.no_grad():
            logits, _ = m.predict_intents(x)
        worst_l = max(worst_l, rel_err(logits.cpu(), g["logits"]))
        named = dict(m.named_parameters())
        for k in p:
            if "g/" + k in g.files and "filt_" not in k:      # (the two cut-off vectors carry max-pool routing noise, see tests)
                flat = named[k].grad.flatten().cpu()
                sub = flat if flat.numel() <= 30000 else flat[::7]
                ref = torch.from_numpy(g["g/" + k])
                if ref.abs().max() > 0:
                    worst_g = max(worst_g, rel_err(sub, ref))
    out["synthetic_logits_rel_err_max"] = worst_l
    out["synthetic_grad_rel_err_max"] = worst_g
    return out


def throughput(B=256, T=64000, steps=12):
    cfgmod = importlib.import_module("end-to-end-slu_b200.config")
    cfg = cfgmod.read_config(os.path.join(ROOT, "configs", "unfreeze_all_layers.cfg"))
    cfg.pretraining_type = 0
    cfg.Sy_intent, cfg.values_per_slot = cfgmod.fsc_intent_table()
    cfg.num_phonemes = 42
    torch.manual_seed(cfg.seed)
    m = models.Model(cfg).train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    gen = torch.Generator().manual_seed(1)
    xs = [(0.1 * torch.randn(B, T, generator=gen)).cuda() for _ in range(4)]
    ys = [torch.stack([torch.randint(0, v, (B,), generator=gen) for v in (6, 14, 4)], 1).cuda() for _ in range(4)]

    def step(i):
        loss, _ = m(xs[i % 4], ys[i % 4]); opt.zero_grad(); loss.backward(); opt.step()
    for i in range(6):
        step(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(steps):
        step(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"utt_per_s": B / ms * 1e3, "ms_per_step": ms}


res = {}
for mode in ("bf16x3", "bf16x3-separate", "fp16"):
    pkg.ops.set_gru_precision(mode)
    res[mode] = goldens()
    res[mode].update(throughput())
pkg.ops.set_gru_precision("bf16x3")
print(json.dumps({"what": "GRU recurrence operand formats: errors against the reference goldens vs train-step throughput (config 3)",
                  "logit_bar": 1e-3, "modes": res}, indent=1))
