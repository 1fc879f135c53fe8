"""CUDA execution of the encoder / SLU forward over the kernels in ops.py.

The nn.Module tree built by the repo-root models.py only HOLDS parameters (so state_dict,
freeze/unfreeze, Adam, .cpu()/.cuda() behave exactly like the reference's); this file is what
runs when those parameters live on a CUDA device.  Unsupported architectures raise -- there is
no silent fallback to library kernels or to the CPU.
"""
import os

import torch

from . import _lib, grads, ops


# One weight-image launch per forward (slu_presplit_multi) instead of one per GEMM; SLU_MULTI_PRESPLIT=0 restores the latter (A/B).
MULTI_PRESPLIT = os.environ.get("SLU_MULTI_PRESPLIT", "1") != "0"


class Plan:
    """Static description of one PretrainedModel's layer lists, extracted once at construction
    (reference layer order: models.py:180-286)."""

    def __init__(self, phoneme_layers, word_layers):
        self.sinc = None
        self.convs = []          # (conv module, negative_slope)
        self.phone = []          # (gru, dropout p, ds)
        self.word = []
        self._scan(list(phoneme_layers), self.phone, cnn=True)
        self._scan(list(word_layers), self.word, cnn=False)

    def _scan(self, layers, rnn_out, cnn):
        i = 0
        n = len(layers)
        while i < n:
            l = layers[i]
            name = getattr(l, "name", "")
            if name.startswith("sinc"):
                _require(cnn and l.N_filt == 80 and l.Filt_dim == 401 and l.stride == 80 and l.padding == 200
                         and int(l.fs) == 16000, "SincLayer must be 80 filters x 401 taps, stride 80, pad 200, fs 16000")
                self.sinc = l
                # abs0, pool0 (2, ceil), act0, dropout0 follow
                pool = layers[i + 2]
                _require(_pool_len(pool) == 2, "cnn_max_pool_len[0] must be 2")
                _require(layers[i + 4].p == 0.0, "cnn_drop must be 0")
                i += 5
            elif isinstance(l, torch.nn.Conv1d):
                _require(cnn and l.in_channels != 1, "use_sincnet=False first layer is not supported on CUDA")
                _require(l.kernel_size[0] % 2 == 1 and l.stride[0] == 1 and l.padding[0] == l.kernel_size[0] // 2,
                         "conv layers must be odd-k, stride 1, same padding")
                _require(_pool_len(layers[i + 1]) == 1, "cnn_max_pool_len[1:] must be 1")
                _require(l.out_channels % 4 == 0 and l.in_channels % 4 == 0, "conv channel counts must be multiples of 4")
                act = layers[i + 2]
                slope = act.negative_slope if isinstance(act, torch.nn.LeakyReLU) else 0.0
                _require(layers[i + 3].p == 0.0, "cnn_drop must be 0")
                self.convs.append((l, slope))
                i += 4
            elif isinstance(l, torch.nn.GRU):
                _require(l.hidden_size == 128 and l.bidirectional and l.num_layers == 1 and l.batch_first,
                         "GRU layers must be single-layer bidirectional with 128 hidden units")
                drop, down = layers[i + 2], layers[i + 3]
                rnn_out.append((l, drop.p, downsample_factor(down)))
                i += 4
            else:           # ncl2nlc and friends carry no parameters / are folded into the layouts
                i += 1


def _pool_len(pool):
    k = pool.kernel_size
    return k if isinstance(k, int) else k[0]


def _require(cond, msg):
    if not cond:
        raise NotImplementedError("slu_b200 CUDA path: " + msg)


def downsample_factor(down):
    if down.method == "avg" and down.factor == 2:
        return 2
    if down.method == "none" and down.factor == 1:
        return 1
    if down.method == "avg" and down.factor == 1:
        return 1
    raise NotImplementedError("slu_b200 CUDA path: Downsample(%s, %d) is not supported" % (down.method, down.factor))


# SLU_DROPOUT_MASKS=1: materialise the dropout masks as tensors (slu_dropout_mask_gru writes the same canonical Philox mask the
# kernels otherwise generate in registers) -- for A/B checks; the default keeps them out of HBM altogether.
MASK_TENSORS = os.environ.get("SLU_DROPOUT_MASKS", "0") != "0"


def _drop_mask(shape, p, training, device):
    """Dropout of one GRU layer output [B,T,256] -> None (eval / p = 0) or (p, seed): the persistent-GRU kernels regenerate
    the canonical Philox keep-mask of that pair in registers, forward and backward (csrc/philox.cuh; reference nn.Dropout at
    models.py:246/276/700).  The seed is a host draw from torch's CPU generator, so `torch.manual_seed` reproduces a run.
    Tests replace this function to supply explicit mask tensors (the reference-order golden)."""
    if not training or p <= 0.0:
        return None
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    if _graph_seed_word is not None:                 # the step is being captured as a CUDA graph: see graphed_train_step
        return (float(p), seed, _graph_seed_word)
    if MASK_TENSORS:
        mask = torch.empty(shape, device=device, dtype=torch.float32)
        _lib.call("slu_dropout_mask_gru", _lib.ptr(mask), shape[0], shape[1], float(p), seed, _lib.stream())
        return mask
    return (float(p), seed)


_default_drop_mask = _drop_mask


def _premask(stacks, B, T, training, device):
    """Dropout arguments of several GRU stacks ahead of time ([per stack [per layer]], same draw order as layer by layer)."""
    out = []
    for rnns in stacks:
        cur = []
        for _, p, ds in rnns:
            cur.append(_drop_mask((B, T, 256), p, training, device))
            T = (T + ds - 1) // ds
        out.append(cur)
    return out, None


def _weight_images(convs, rnns):
    """Operand images (bf16 hi/lo, k-chunk-major) of every GEMM weight of the given layers, made by ONE launch
    (slu_presplit_multi): x-projection / conv forward forms and, when gradients are on, the input-gradient forms that the
    backward pass will use.  Returns ([per conv (fwd, dx)], [per GRU (nt, nn)]); None entries = let the op make its own."""
    if not MULTI_PRESPLIT:
        return [None] * len(convs), [None] * len(rnns)
    need_dx = torch.is_grad_enabled()
    items, owners = [], []
    for li, (conv, _) in enumerate(convs):
        it = ops.conv_weight_items(conv, need_dx)
        if it is not None:
            owners += [("c", li, k) for k in range(len(it))]
            items += it
    for li, (gru, _, _) in enumerate(rnns):
        it = ops.gru_weight_items(gru, need_dx)
        if it is not None:
            owners += [("g", li, k) for k in range(len(it))]
            items += it
    imgs = ops.presplit_many(items) if items else []
    conv_imgs, gru_imgs = [None] * len(convs), [None] * len(rnns)
    for (kind, li, k), img in zip(owners, imgs):
        dst = conv_imgs if kind == "c" else gru_imgs
        pair = list(dst[li]) if dst[li] is not None else [None, None]
        pair[k] = img
        dst[li] = tuple(pair)
    return conv_imgs, gru_imgs


def _run_rnns(out, rnns, training, masks=None, join=None, imgs=None):
    if imgs is None and rnns:
        imgs = _weight_images([], rnns)[1]
    for i, (gru, p, ds) in enumerate(rnns):
        B, T, _ = out.shape
        mask = masks[i] if masks is not None else _drop_mask((B, T, 256), p, training, out.device)
        if torch.is_tensor(mask) and tuple(mask.shape) != (B, T, 256):
            raise RuntimeError("slu_b200: pre-generated dropout mask does not match the layer input")
        out = ops.bigru(out, gru, mask, ds, join if i == 0 else None, imgs[i])   # masks are joined right before the first recurrence
    if join is not None and not rnns:
        join()
    return out


def phoneme_features(pm, x, with_word=True):
    """[B,T] waveform -> ([B, ceil(T/640), 256] output of the phoneme module, carry).  `carry` = what was prepared in the same
    pass for the word module that follows (its dropout masks and weight operand images), handed to word_features() explicitly;
    with_word=False (phoneme-only pre-training, models.py:243) prepares nothing for it."""
    plan = pm._plan
    _require(plan.sinc is not None, "use_sincnet must be True")
    grads.begin(x.device)            # a new forward pass: its weight gradients will share one flat arena (grads.py)
    out = ops.SincFrontend.apply(x, plan.sinc.filt_b1, plan.sinc.filt_band)      # [B, L1, 80] (LeakyReLU is identity on >=0)
    for conv, slope in plan.convs:
        out = ops.conv_block(out, conv.weight, conv.bias, slope)
    # W_ih operand images of both GRU stacks in one launch, queued while the GPU is busy with the front end (an A/B on the same
    # box showed that making them -- and the conv images -- BEFORE the front end costs more host time in the idle gap at the
    # start of a step than the saved launches are worth)
    word = plan.word if with_word else []
    _, gru_imgs = _weight_images([], plan.phone + word)
    # Dropout of the phoneme AND word stacks: one (p, seed) pair per layer, drawn now so the word module reuses this pass's draws.
    (m_phone, m_word), join = _premask([plan.phone, word], out.shape[0], out.shape[1], pm.training, out.device)
    out = _run_rnns(out, plan.phone, pm.training, m_phone, join, gru_imgs[:len(plan.phone)])
    carry = (m_word if pm.training else None, gru_imgs[len(plan.phone):]) if with_word else None
    return out, carry


def word_features(pm, ph, carry=None):
    """Word module on the phoneme module's output; `carry` comes from the phoneme_features() call that produced `ph`."""
    masks, imgs = carry if carry is not None else (None, None)
    return _run_rnns(ph, pm._plan.word, pm.training, masks, None, imgs)


def compute_features(pm, x):
    ph, carry = phoneme_features(pm, x)
    return word_features(pm, ph, carry)


def intent_logits(model, feats):
    """intent GRU(s) -> Linear -> max over time (models.py:806-809)."""
    out = _run_rnns(feats, model._intent_rnns, model.training)
    lin = model._final_classifier
    _require(ops.intent_head_supported(lin.weight, (lin.weight.shape[0],)),
             "the intent head kernel takes 256 features and at most 128 values in total")
    return ops.intent_head_logits(out, lin.weight, lin.bias)


def intent_loss_acc(model, x, y_intent):
    """Training tail of Model.forward (models.py:806-823) as one kernel per direction: intent GRU(s) -> Linear -> max over
    time -> summed per-slot cross-entropy and all-slots-right accuracy.  Returns None when the labels do not fit the fused loss
    (more than 16 slots, unusual label layout): the caller then takes intent_logits() -- the same head kernel, differentiable --
    and composes the loss on the [B, C] logits."""
    lin = model._final_classifier
    slots = tuple(int(n) for n in model.values_per_slot)
    if not (ops.intent_head_supported(lin.weight, slots) and y_intent.dtype == torch.int64 and y_intent.dim() == 2):
        return None
    out = _run_rnns(model.pretrained_model.compute_features(x), model._intent_rnns, model.training)
    loss, acc, _ = ops.IntentHead.apply(out, lin.weight, lin.bias, y_intent, slots)
    return loss, acc


# ---- the SLU train step as two CUDA graphs (forward, backward) ------------------------------------------------------------------
# A train step at these sizes is ~55 kernel launches for ~2.8 ms of GPU work (600+ launches for the seq2seq decoder): close to
# host-bound, and the reference's Trainer reads the loss back every step (training.py:99-100), so the host cannot run ahead.
# When the same training forward (shapes, trainable set, parameter addresses) has been seen twice, its third occurrence is
# CAPTURED: graph F = seed word advance + forward (-> static loss / acc), graph B = the backward pass of that loss given a
# static dL/dloss (torch.autograd.grad inside the capture: the gradients are the arena views the kernels write).  From then
# on `Model.forward` = copy the batch into the static input buffers + replay F; `loss.backward()` = copy dL/dloss + replay B +
# hand the static gradient tensors to the parameters.  The Trainer is unchanged; the optimizer step and the data-parallel
# all-reduce stay outside the graphs.  Dropout seeds are frozen into a graph as by-value arguments, so the kernels XOR them with
# a device word that graph F advances on every replay (slu_seed_advance).  SLU_STEP_GRAPH=0 disables all of this.
STEP_GRAPH = os.environ.get("SLU_STEP_GRAPH", "1") != "0"
GRAPH_WARMUP = 2                    # eager occurrences of a key before it is captured
GRAPH_CACHE = 3                     # captured steps kept per model (least recently used goes)
_graph_seed_word = None             # set while a step is being captured: _drop_mask hands it to the kernels


def graph_seed_word():
    return _graph_seed_word


class _GraphedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dummy, runner):
        ctx.runner, ctx.epoch = runner, runner.epoch
        acc = runner.acc.clone()
        ctx.mark_non_differentiable(acc)
        return runner.loss.clone(), acc

    @staticmethod
    def backward(ctx, g_loss, g_acc):
        r = ctx.runner
        if ctx.epoch != r.epoch:
            raise RuntimeError("slu_b200: backward through a graphed train step whose buffers a later forward has already reused "
                               "(call backward before the next forward, or set SLU_STEP_GRAPH=0)")
        r.g_in.copy_(g_loss.reshape(r.g_in.shape))
        r.bwd.replay()
        _lib.stats["calls"] += r.n_bwd
        for p, g in r.grads:
            p.grad = g if p.grad is None else p.grad + g
        return None, None


def grads_arena_of(pairs):
    for _, g in pairs:
        a = grads.find(g)
        if a is not None:
            return grads.pin(a)
    return None


class _StepGraph:
    def __init__(self, model, x, y, eager):
        dev = next(model.parameters()).device
        self.sx = torch.empty(tuple(x.shape), device=dev, dtype=torch.float32)
        self.sy = torch.empty(tuple(y.shape), device=dev, dtype=y.dtype)
        self.sx.copy_(x); self.sy.copy_(y)
        self.seed_word = torch.randint(0, 2 ** 62, (1,)).to(dev)
        self.g_in = torch.ones((), device=dev, dtype=torch.float32)
        self.dummy = torch.zeros((), device=dev, requires_grad=True)
        self.epoch = 0
        # The capture runs on torch's capture stream, but a Parameter's AccumulateGrad node remembers the stream it was created on
        # (the default stream, if anything of an earlier eager step -- e.g. the loss the Trainer still holds -- keeps it alive), and
        # the autograd engine would then try to make that stream wait for the capturing one.  So the captured forward runs on fresh
        # leaf ALIASES of the parameters (same storage, new autograd identity) swapped into the module tree for the duration.
        from torch.nn.utils.stateless import _reparametrize_module
        named = list(model.named_parameters())
        alias = {n: p.detach().requires_grad_(p.requires_grad) for n, p in named}
        params = [p for _, p in named if p.requires_grad]
        leaves = [alias[n] for n, p in named if p.requires_grad]
        saved = [(p, p.grad) for p in params]
        global _graph_seed_word
        torch.cuda.synchronize()
        self.fwd, self.bwd = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        c0 = _lib.stats["calls"]
        try:
            _graph_seed_word = self.seed_word
            with torch.cuda.graph(self.fwd, capture_error_mode="thread_local"):
                _lib.call("slu_seed_advance", self.seed_word.data_ptr(), _lib.stream())
                with _reparametrize_module(model, alias):
                    loss, acc = eager(self.sx, self.sy)
            c1 = _lib.stats["calls"]
            with torch.cuda.graph(self.bwd, pool=self.fwd.pool(), capture_error_mode="thread_local"):
                grads = torch.autograd.grad((loss,), leaves, (self.g_in.reshape(loss.shape),), allow_unused=True)
        finally:
            _graph_seed_word = None
        self.n_fwd, self.n_bwd = c1 - c0, _lib.stats["calls"] - c1
        _lib.stats["calls"] = c0                     # nothing ran during capture
        self.loss, self.acc = loss.detach(), acc.detach().to(dev) if not acc.is_cuda else acc.detach()
        self.grads = [(p, g) for p, g in zip(params, grads) if g is not None]
        # the arena the captured backward writes into lives as long as this graph: keep it findable for dp.py's in-place all-reduce
        self.arena = grads_arena_of(self.grads)
        for p, g in saved:
            p.grad = g

    def run(self, x, y):
        self.sx.copy_(x, non_blocking=True)
        self.sy.copy_(y, non_blocking=True)
        self.epoch += 1
        self.fwd.replay()
        _lib.stats["calls"] += self.n_fwd
        return _GraphedLoss.apply(self.dummy, self)


def graphed_train_step(model, x, y, eager):
    """Model.forward's training path on CUDA: replay the captured step when there is one for this (shapes, trainable set,
    parameter addresses), count occurrences / capture otherwise.  Returns (loss, acc) or None = run `eager` yourself."""
    if not STEP_GRAPH or _lib._prof is not None or torch.cuda.is_current_stream_capturing():
        return None
    params = list(model.parameters())
    if any(p.requires_grad and p.grad is not None for p in params):
        return None                                   # gradient accumulation across steps: stock semantics, eager
    key = (tuple(x.shape), tuple(y.shape), y.dtype, tuple(p.requires_grad for p in params), tuple(p.data_ptr() for p in params),
           ops.GRU_IMPL, ops.SINC_IMPL, _drop_mask is _default_drop_mask)
    cache = model.__dict__.setdefault("_step_graphs", {})
    entry = cache.get(key)
    if isinstance(entry, _StepGraph):
        cache[key] = cache.pop(key)                   # most recently used last
        return entry.run(x, y)
    if not key[-1]:
        return None                                   # a test supplies explicit mask tensors: eager only
    seen = (entry or 0) + 1
    if seen <= GRAPH_WARMUP:
        cache[key] = seen
        for k in [k for k, v in cache.items() if not isinstance(v, _StepGraph)][:-8]:
            del cache[k]                              # bound the bookkeeping of one-off shapes
        return None
    graphs = [k for k, v in cache.items() if isinstance(v, _StepGraph)]
    for k in graphs[:max(0, len(graphs) - GRAPH_CACHE + 1)]:
        del cache[k]
    runner = cache[key] = _StepGraph(model, x, y, eager)
    return runner.run(x, y)
