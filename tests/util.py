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
