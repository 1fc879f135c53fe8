"""Flat gradient arena: every weight-gradient kernel of one backward pass writes into ONE zero-filled fp32 buffer and the
`.grad` tensors autograd hands to the parameters are views of it.

Why: (a) one memset per step instead of one `torch.zeros` per layer (the split-K / atomic kernels accumulate), (b) the
data-parallel exchange all-reduces the buffer as it is -- no pack / unpack copies (dp.py), (c) the fused Adam step walks one
address range.  The reference's Trainer calls `optimizer.zero_grad()` (set_to_none) every step (training.py:64/96), which drops the
views before the next backward; the next arena then gets the same block back from the caching allocator.

Life cycle: `begin(device)` at the start of a training forward (engine.phoneme_features) makes a fresh Arena "current"; each
autograd Function reserves its slot in forward (`reserve`) and asks for the view in backward (`view`); the buffer itself is
allocated (and zero-filled by one memset) at the first `view` of the backward pass.  A slot that is asked for twice (a second
backward through a retained graph) gets a private zero tensor instead, so accumulation semantics stay those of stock autograd.

fp64 slots (the two SincNet cut-off vectors) live at the front of the buffer, 8-byte aligned; `hilo` is a float region of the
same element count x 2 behind them that dp.py uses to carry them through the fp32 collective as (hi, lo) pairs.
"""
import weakref

import torch

_current = None
_recent = []          # weak references to the arenas of the last few forwards (dp.py looks grads up in them)


class Arena:
    def __init__(self, device):
        self.device = torch.device(device)
        self.n64 = 0                    # doubles reserved at the front
        self.slots64 = {}               # key -> (offset in doubles, numel)
        self.n32 = 0
        self.slots32 = {}               # key -> (offset in floats, counted after the fp64 + hilo regions), numel
        self.flat = None
        self.taken = set()
        self.frozen = False

    # ---- forward: reserve --------------------------------------------------------------------------------------------------
    def reserve(self, numel, f64=False):
        """-> slot key for view(), or None (the layout is fixed once the buffer exists: a late op uses its own zeros)."""
        if self.frozen:
            return None
        if f64:
            key = len(self.slots64)
            self.slots64[key] = (self.n64, numel)
            self.n64 += numel
        else:
            key = len(self.slots32)
            self.slots32[key] = (self.n32, numel)
            self.n32 += (numel + 3) // 4 * 4          # 16-byte aligned slots (vector reductions / float4 Adam traffic)
        return key

    # ---- backward: views ---------------------------------------------------------------------------------------------------
    @property
    def head(self):                     # floats in front of the fp32 slots: fp64 region (2 floats each) + hi/lo staging (2 floats each)
        return (4 * self.n64 + 3) // 4 * 4

    def buffer(self):
        if self.flat is None:
            self.frozen = True
            self.flat = torch.zeros(self.head + self.n32, device=self.device, dtype=torch.float32)     # ONE memset per backward pass
        return self.flat

    def view(self, key, shape, f64=False):
        """Zero-initialised gradient buffer for `key` ([*shape]); None if the slot was never reserved."""
        table = self.slots64 if f64 else self.slots32
        if key is None or key not in table:
            return None
        off, n = table[key]
        if (key, f64) in self.taken:    # second backward through the same graph: do not alias the first one's .grad
            return torch.zeros(shape, device=self.device, dtype=torch.float64 if f64 else torch.float32)
        self.taken.add((key, f64))
        flat = self.buffer()
        if f64:
            return flat[:2 * self.n64].view(torch.float64)[off:off + n].view(shape)
        return flat[self.head + off:self.head + off + n].view(shape)

    def f64_region(self):
        """(doubles [n64], hi floats [n64], lo floats [n64]) views of the front of the buffer."""
        flat = self.buffer()
        n = self.n64
        return flat[:2 * n].view(torch.float64), flat[2 * n:3 * n], flat[3 * n:4 * n]

    def owns(self, t):
        return self.flat is not None and t.untyped_storage().data_ptr() == self.flat.untyped_storage().data_ptr()


def begin(device):
    """Start a new arena for the forward pass that begins now (no-op outside grad mode)."""
    global _current
    if not torch.is_grad_enabled():
        _current = None
        return None
    _current = Arena(device)
    _recent.append(weakref.ref(_current))
    del _recent[:-4]
    return _current


def current():
    return _current


_pinned = []          # arenas of captured CUDA-graph steps (engine._StepGraph): static buffers that outlive their forward pass


def pin(arena):
    if arena is not None and arena not in _pinned:
        _pinned.append(arena)
        del _pinned[:-16]
    return arena


def find(t):
    """The live arena whose buffer `t` is a view of, or None."""
    for ref in reversed(_recent):
        a = ref()
        if a is not None and a.owns(t):
            return a
    for a in reversed(_pinned):
        if a.owns(t):
            return a
    return None
