import os
import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def ckpt_params():
    z = golden("ckpt_no_unfreezing_slu.npz")
    return {k: torch.from_numpy(z[k]) for k in z.files}


def load_test_wav():
    z = golden("test_wav.npz")
    return torch.from_numpy(z["pcm"].astype(np.float32) / 32768).unsqueeze(0)


def rel_err(a, b):
    a = torch.as_tensor(a).double(); b = torch.as_tensor(b).double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def make_config(name="no_unfreezing", pretraining_type=0, **over):
    """Config for a random-init model (no checkpoint on disk): the shipped cfg with
    pretraining_type overridden and the FSC slot table filled in."""
    import importlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfgmod = importlib.import_module("end-to-end-slu_b200.config")
    cfg = cfgmod.read_config(os.path.join(root, "configs", name + ".cfg"))
    cfg.pretraining_type = pretraining_type
    cfg.Sy_intent, cfg.values_per_slot = cfgmod.fsc_intent_table()
    cfg.num_phonemes = 42
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg
