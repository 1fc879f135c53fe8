"""Data-parallel gradient exchange: ONE all-reduce per optimizer step over a flat fp32 bucket.

The reference has no distributed code.  The utterance minibatch shards across
ranks (one process per GPU); parameters are replicated; after backward every rank holds grads of
its shard's mean loss, and the mean over ranks is the global-batch gradient (equal shard sizes).

`install()` registers a global optimizer pre-step hook (torch.optim.optimizer.register_optimizer_step_pre_hook),
so the reference's Trainer (`loss.backward(); optimizer.step()`, training.py:65-66/97-98) needs no change:
on the CUDA path every `.grad` is a view of ONE flat arena (grads.py) that the weight-gradient kernels wrote into, so the hook
all-reduces that buffer in place (NCCL AVG) -- no pack / unpack; the fp64 sinc grads cross the fp32 collective as (hi, lo) float
pairs.  Gradients that are not arena views (CPU / gloo runs, modules executed by torch ops such as the seq2seq decoder) take the
generic path: pack into a cached flat bucket, one all-reduce, unpack.
Parameters whose `.grad` is None (frozen layers, the unused ASR heads) are skipped; the set is re-derived
each step, which follows `unfreeze_one_layer()` for free.  All ranks must agree on that set.
"""
import torch
import torch.distributed as dist

from . import _lib, grads as grads_mod

_installed = False
stats = {"allreduce_calls": 0, "bucket_bytes": 0}


class _Bucket:
    """Flat fp32 bucket + per-gradient views, built once per set of gradient shapes (the host cost of slicing / concatenating
    ~50 tensors every step is what makes a multi-GPU step host-bound, not the collective)."""

    def __init__(self, grads):
        self.key = _key(grads)
        n = sum(g.numel() * (2 if g.dtype == torch.float64 else 1) for g in grads)
        self.flat = torch.empty(n, device=grads[0].device, dtype=torch.float32)
        self.views32, self.idx32, self.views64, self.idx64 = [], [], [], []
        off = 0
        for i, g in enumerate(grads):
            k = g.numel()
            if g.dtype == torch.float64:                      # carried as an fp32 (hi, lo) pair inside the single bucket
                self.views64.append((self.flat[off:off + k].view_as(g), self.flat[off + k:off + 2 * k].view_as(g)))
                self.idx64.append(i)
                off += 2 * k
            else:
                self.views32.append(self.flat[off:off + k].view_as(g))
                self.idx32.append(i)
                off += k

    def pack(self, grads):
        if self.idx32:
            torch._foreach_copy_(self.views32, [grads[i] for i in self.idx32])
        for (hi, lo), i in zip(self.views64, self.idx64):
            g = grads[i]
            hi.copy_(g)
            lo.copy_(g - hi.double())

    def unpack(self, grads):
        if self.idx32:
            torch._foreach_copy_([grads[i] for i in self.idx32], self.views32)
        for (hi, lo), i in zip(self.views64, self.idx64):
            grads[i].copy_(hi.double() + lo.double())


def _key(grads):
    return tuple((g.shape, g.dtype, g.device) for g in grads)


_bucket = None


def allreduce_grads(params, group=None):
    """Average `.grad` of `params` across ranks with a single collective.  Returns bytes exchanged."""
    global _bucket
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    arena = grads_mod.find(grads[0]) if grads[0].is_cuda else None
    if arena is not None and all(arena.owns(g) for g in grads):
        # Every gradient is a view of this backward pass's arena (grads.py): the buffer IS the bucket -- no pack, no unpack.
        # The fp64 SincNet gradients sit at its front; they cross the fp32 collective as (hi, lo) float pairs staged behind
        # them (2 tiny launches), and the merge overwrites whatever the collective made of the raw fp64 words.
        flat = arena.flat
        if arena.n64:
            d, hi, lo = arena.f64_region()
            _lib.call("slu_f64_hilo_split", d.data_ptr(), hi.data_ptr(), lo.data_ptr(), arena.n64, _lib.stream())
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
        if arena.n64:
            _lib.call("slu_f64_hilo_merge", d.data_ptr(), hi.data_ptr(), lo.data_ptr(), arena.n64, _lib.stream())
        stats["allreduce_calls"] += 1
        stats["bucket_bytes"] = flat.numel() * 4
        stats["arena"] = stats.get("arena", 0) + 1
        return stats["bucket_bytes"]
    if _bucket is None or _bucket.key != _key(grads):
        _bucket = _Bucket(grads)
    _bucket.pack(grads)
    flat = _bucket.flat
    if flat.is_cuda:                                   # NCCL averages inside the collective
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.mul_(1.0 / dist.get_world_size(group))
    _bucket.unpack(grads)
    stats["allreduce_calls"] += 1
    stats["bucket_bytes"] = flat.numel() * 4
    return stats["bucket_bytes"]


def _pre_step_hook(optimizer, args, kwargs):
    params = [p for grp in optimizer.param_groups for p in grp["params"]]
    allreduce_grads(params)


def install():
    """Idempotently register the global optimizer pre-step hook (no-op for single-process runs)."""
    global _installed
    if not _installed:
        from torch.optim.optimizer import register_optimizer_step_pre_hook
        register_optimizer_step_pre_hook(_pre_step_hook)
        _installed = True
