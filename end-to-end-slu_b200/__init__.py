"""slu_b200: Blackwell-native (sm_100a) speech-encoder hot path for end-to-end SLU.

The directory name contains a hyphen (task layout), so import it with
    importlib.import_module("end-to-end-slu_b200")
or through the repo-root `models.py`, which is the drop-in surface.
Only what the hot path needs lives here: csrc/ (CUDA kernels + C-ABI), _lib (ctypes binding to the
C-ABI shared library), ops (autograd wrappers), engine (the encoder/SLU forward on CUDA).
"""
from . import _lib, grads, ops, engine, dp, config, loader, optim, decoder  # noqa: F401
