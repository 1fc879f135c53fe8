"""Data-parallel gradient exchange: ONE all-reduce per optimizer step over a flat fp32 bucket.

The reference has no distributed code.  The utterance minibatch shards across
ranks (one process per GPU); parameters are replicated; after backward every rank holds grads of
its shard's mean loss, and the mean over ranks is the global-batch gradient (equal shard sizes).

`install()` registers a global optimizer pre-step hook (torch.optim.optimizer.register_optimizer_step_pre_hook),
so the reference's Trainer (`loss.backward(); optimizer.step()`, training.py:65-66/97-98) needs no change:
the hook packs every existing `.grad` (fp64 sinc grads are carried as fp32 pairs hi/lo to stay inside the
single fp32 bucket), issues one NCCL all-reduce on the current stream, and unpacks scaled by 1/world.
Parameters whose `.grad` is None (frozen layers, the unused ASR heads) are skipped; the set is re-derived
each step, which follows `unfreeze_one_layer()` for free.  All ranks must agree on that set.
"""
import torch
import torch.distributed as dist

_installed = False
stats = {"allreduce_calls": 0, "bucket_bytes": 0}


def _flatten(grads):
    parts = []
    for g in grads:
        if g.dtype == torch.float64:
            hi = g.float()
            parts.append(hi.reshape(-1))
            parts.append((g - hi.double()).float().reshape(-1))
        else:
            parts.append(g.reshape(-1))
    return torch.cat(parts)


def _unflatten(flat, grads):
    """Write the (already averaged) bucket back: one multi-tensor copy for the fp32 gradients."""
    off = 0
    dst, src = [], []
    for g in grads:
        n = g.numel()
        if g.dtype == torch.float64:
            g.copy_((flat[off:off + n].double() + flat[off + n:off + 2 * n].double()).view_as(g))
            off += 2 * n
        else:
            dst.append(g)
            src.append(flat[off:off + n].view_as(g))
            off += n
    if dst:
        torch._foreach_copy_(dst, src)


def allreduce_grads(params, group=None):
    """Average `.grad` of `params` across ranks with a single collective.  Returns bytes exchanged."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    flat = _flatten(grads)
    if flat.is_cuda:                                   # NCCL averages inside the collective
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.mul_(1.0 / dist.get_world_size(group))
    _unflatten(flat, grads)
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
