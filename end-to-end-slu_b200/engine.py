"""CUDA execution of the encoder / SLU forward over the kernels in ops.py.

The nn.Module tree built by the repo-root models.py only HOLDS parameters (so state_dict,
freeze/unfreeze, Adam, .cpu()/.cuda() behave exactly like the reference's); this file is what
runs when those parameters live on a CUDA device.  Unsupported architectures raise -- there is
no silent fallback to library kernels or to the CPU.
"""
import torch

from . import _lib, ops


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


def _drop_mask(shape, p, training, device):
    if not training or p <= 0.0:
        return None
    mask = torch.empty(shape, device=device, dtype=torch.float32)
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())        # host draw from torch's CPU generator: follows torch.manual_seed
    _lib.call("slu_dropout_mask", _lib.ptr(mask), mask.numel(), float(p), seed, _lib.stream())
    return mask


def _run_rnns(out, rnns, training):
    for gru, p, ds in rnns:
        B, T, _ = out.shape
        out = ops.bigru(out, gru, _drop_mask((B, T, 256), p, training, out.device), ds)
    return out


def phoneme_features(pm, x):
    """[B,T] waveform -> [B, ceil(T/640), 256] (output of the phoneme module)."""
    plan = pm._plan
    _require(plan.sinc is not None, "use_sincnet must be True")
    out = ops.SincFrontend.apply(x, plan.sinc.filt_b1, plan.sinc.filt_band)      # [B, L1, 80] (LeakyReLU is identity on >=0)
    for conv, slope in plan.convs:
        out = ops.conv_block(out, conv.weight, conv.bias, slope)
    return _run_rnns(out, plan.phone, pm.training)


def word_features(pm, ph):
    return _run_rnns(ph, pm._plan.word, pm.training)


def compute_features(pm, x):
    return word_features(pm, phoneme_features(pm, x))


def intent_logits(model, feats):
    """intent GRU(s) -> Linear -> max over time (models.py:806-809)."""
    out = _run_rnns(feats, model._intent_rnns, model.training)
    lin = model._final_classifier
    if ops.intent_head_supported(lin.weight, (lin.weight.shape[0],)):
        return ops.intent_head_logits(out, lin.weight, lin.bias)
    B, T, C = out.shape
    logits = torch.addmm(lin.bias, out.reshape(B * T, C), lin.weight.t()).view(B, T, -1)
    return logits.max(dim=1)[0]


def intent_loss_acc(model, x, y_intent):
    """Training tail of Model.forward (models.py:806-823) as one kernel per direction: intent GRU(s) -> Linear -> max over
    time -> summed per-slot cross-entropy and all-slots-right accuracy.  Returns None when the head does not fit the kernel
    (more than 128 values / 16 slots), and the caller composes it from library ops."""
    lin = model._final_classifier
    slots = tuple(int(n) for n in model.values_per_slot)
    if not (ops.intent_head_supported(lin.weight, slots) and y_intent.dtype == torch.int64 and y_intent.dim() == 2):
        return None
    out = _run_rnns(model.pretrained_model.compute_features(x), model._intent_rnns, model.training)
    loss, acc, _ = ops.IntentHead.apply(out, lin.weight, lin.bias, y_intent, slots)
    return loss, acc
