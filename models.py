"""Drop-in replacement for the reference's `models.py` (lorenlugosch/end-to-end-SLU).

`main.py` / `training.py` / the README snippet of the reference import `PretrainedModel` and
`Model` from a module called `models`; put this repo first on PYTHONPATH and they get these
classes instead.  Same constructor (`data.Config` from `read_config`), same
methods and return values, same `state_dict` keys/shapes/dtypes (shipped `.pth` files load
strictly), same freeze / unfreeze schedule.

What differs is execution.  The module tree below only HOLDS the parameters.  When they live on
a CUDA device, forward work goes to the sm_100a kernels of `end-to-end-slu_b200/` (fused
SincConv+abs+pool, persistent bidirectional GRU with fused gates/dropout/downsample, and their
backward kernels) through a C-ABI library; if that library is missing the call raises.  When the
parameters are on the CPU (the reference's Trainer.test() moves the model there for validation,
training.py:150) the same math runs as plain torch ops.

Reference behaviour cited as models.py:<line> refers to /root/reference/models.py.
"""
import importlib
import math
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

_pkg = None


def _engine():
    """The CUDA engine package (directory name has a hyphen, hence importlib)."""
    global _pkg
    if _pkg is None:
        here = os.path.dirname(os.path.abspath(__file__))
        if here not in sys.path:
            sys.path.insert(0, here)
        _pkg = importlib.import_module("end-to-end-slu_b200")
        if torch.cuda.is_available():
            _pkg.optim.install()         # torch.optim.Adam -> the fused-step subclass the unchanged Trainer then constructs
    return _pkg.engine


def _ops():
    _engine()
    return _pkg.ops


# ---------------------------------------------------------------------------------------------
# small parameter-free layers (names kept: they appear in pickles / isinstance checks)
# ---------------------------------------------------------------------------------------------
class Downsample(torch.nn.Module):
    """Time-axis downsampling: "none" = stride, "avg"/"max" = ceil-mode pooling (models.py:26-46)."""

    METHODS = ("none", "avg", "max")

    def __init__(self, method="none", factor=1, axis=1):
        super().__init__()
        if method not in self.METHODS:
            print("Error: downsampling method must be one of the following: \"none\", \"avg\", \"max\"")
            sys.exit()
        self.method, self.factor, self.axis = method, factor, axis

    def forward(self, x):
        if self.method == "none":
            return x.transpose(self.axis, 0)[::self.factor].transpose(self.axis, 0)
        pool = F.avg_pool1d if self.method == "avg" else F.max_pool1d
        return pool(x.transpose(self.axis, 2), kernel_size=self.factor, ceil_mode=True).transpose(self.axis, 2)


class FinalPool(torch.nn.Module):
    """(B, T, C) -> (B, C): max over time (models.py:112-123)."""

    def forward(self, input):
        return input.max(dim=1)[0]


class NCL2NLC(torch.nn.Module):
    """(B, C, T) -> (B, T, C) (models.py:125-136)."""

    def forward(self, input):
        return input.transpose(1, 2)


class RNNSelect(torch.nn.Module):
    """Keep the per-timestep outputs of an nn.GRU call (models.py:138-149)."""

    def forward(self, input):
        return input[0]


class Abs(torch.nn.Module):
    def forward(self, input):
        return torch.abs(input)


def mel_cutoffs(n_filt, fs):
    """Mel-spaced initial (low edge, bandwidth) in Hz (models.py:56-64)."""
    mel = np.linspace(80, 2595 * np.log10(1 + (fs / 2) / 700), n_filt)
    hz = 700 * (10 ** (mel / 2595) - 1)
    lo, hi = np.roll(hz, 1), np.roll(hz, -1)
    lo[0], hi[-1] = 30, (fs / 2) - 100
    return lo, hi - lo


def sinc_filter_bank(filt_b1, filt_band, n_taps, fs):
    """[N_filt, n_taps] fp32 band-pass bank from the fp64 cut-offs (models.py:82-106), vectorised
    over filters; differentiable (CPU path and tests).  The CUDA path synthesises the same bank
    in a kernel (csrc/sinc.cu)."""
    dev = filt_b1.device
    half = (n_taps - 1) // 2
    t = torch.linspace(1, half, steps=half, device=dev) / fs
    lo = torch.abs(filt_b1) + 50.0 / fs
    hi = lo + (torch.abs(filt_band) + 50.0 / fs)
    n = torch.linspace(0, n_taps, steps=n_taps, device=dev)
    window = (0.54 - 0.46 * torch.cos(2 * math.pi * n / n_taps)).float()

    def low_pass(f):
        f = f.float()
        arg = (2 * math.pi * (f * float(fs)))[:, None] * t[None, :]
        side = torch.sin(arg) / arg
        return 2 * f[:, None] * torch.cat([side.flip(1), torch.ones(f.shape[0], 1, device=dev), side], dim=1)

    band = low_pass(hi) - low_pass(lo)
    return band / band.amax(dim=1, keepdim=True) * window[None, :]


class SincLayer(torch.nn.Module):
    """Parametric band-pass front end (SincNet).  Parameters `filt_b1`, `filt_band` are fp64 and
    normalised by fs exactly as in the reference (models.py:53-75) so checkpoints interchange.
    Unlike the reference (which re-runs conv1d inside its 80-iteration filter loop, models.py:98-108,
    with an identical final result) the bank is built once and convolved once."""

    def __init__(self, N_filt, Filt_dim, fs, stride=1, padding=0, is_cuda=False):
        super().__init__()
        lo, bw = mel_cutoffs(N_filt, fs)
        self.freq_scale = fs * 1.0
        self.filt_b1 = torch.nn.Parameter(torch.from_numpy(lo / self.freq_scale))
        self.filt_band = torch.nn.Parameter(torch.from_numpy(bw / self.freq_scale))
        self.N_filt, self.Filt_dim, self.fs = N_filt, Filt_dim, fs
        self.stride, self.padding, self.is_cuda = stride, padding, is_cuda

    def filters(self):
        return sinc_filter_bank(self.filt_b1, self.filt_band, self.Filt_dim, self.fs)

    def forward(self, x):
        self.is_cuda = self.filt_b1.is_cuda
        w = self.filters().view(self.N_filt, 1, self.Filt_dim)
        return F.conv1d(x, w, stride=self.stride, padding=self.padding)


def _named(layer, name):
    layer.name = name
    return layer


def _rnn_stack(prefix, in_dim, hidden, bidirectional, drops, ds_types, ds_lens):
    """[GRU, RNNSelect, Dropout, Downsample] per entry of `hidden` (models.py:227-253 / 258-283 / 683-707)."""
    layers = []
    for idx, h in enumerate(hidden):
        layers.append(_named(torch.nn.GRU(input_size=in_dim, hidden_size=h, batch_first=True, bidirectional=bidirectional),
                             "%s_rnn%d" % (prefix, idx)))
        in_dim = h * (2 if bidirectional else 1)
        layers.append(_named(RNNSelect(), "%s_rnn_select%d" % (prefix, idx)))
        layers.append(_named(torch.nn.Dropout(p=drops[idx]), "%s_dropout%d" % (prefix, idx)))
        layers.append(_named(Downsample(method=ds_types[idx], factor=ds_lens[idx], axis=1), "%s_downsample%d" % (prefix, idx)))
    return layers, in_dim


def _run(layers, out):
    for layer in layers:
        out = layer(out)
    return out


def _masked_acc(logits, target):
    valid = target != -1
    return (logits.max(1)[1][valid] == target[valid]).float().mean()


class PretrainedModel(torch.nn.Module):
    """Phoneme + word encoders with their ASR heads (models.py:170-361)."""

    def __init__(self, config):
        super().__init__()
        self.is_cuda = torch.cuda.is_available()
        layers = []
        n_conv = len(config.cnn_N_filt)
        for idx in range(n_conv):
            k, stride = config.cnn_len_filt[idx], config.cnn_stride[idx]
            if idx == 0:
                if config.use_sincnet:
                    layers.append(_named(SincLayer(config.cnn_N_filt[0], k, config.fs, stride=stride, padding=k // 2,
                                                   is_cuda=self.is_cuda), "sinc0"))
                else:
                    layers.append(_named(torch.nn.Conv1d(1, config.cnn_N_filt[0], k, stride=stride, padding=k // 2), "conv0"))
                layers.append(_named(Abs(), "abs0"))
            else:
                layers.append(_named(torch.nn.Conv1d(config.cnn_N_filt[idx - 1], config.cnn_N_filt[idx], k, stride=stride,
                                                     padding=k // 2), "conv%d" % idx))
            layers.append(_named(torch.nn.MaxPool1d(config.cnn_max_pool_len[idx], ceil_mode=True), "pool%d" % idx))
            act = torch.nn.LeakyReLU(0.2) if config.cnn_act[idx] == "leaky_relu" else torch.nn.ReLU()
            layers.append(_named(act, "act%d" % idx))
            layers.append(_named(torch.nn.Dropout(p=config.cnn_drop[idx]), "dropout%d" % idx))
        layers.append(_named(NCL2NLC(), "ncl2nlc"))
        rnn, out_dim = _rnn_stack("phone", config.cnn_N_filt[-1], config.phone_rnn_num_hidden, config.phone_rnn_bidirectional,
                                  config.phone_rnn_drop, config.phone_downsample_type, config.phone_downsample_len)
        self.phoneme_layers = torch.nn.ModuleList(layers + rnn)
        self.phoneme_linear = torch.nn.Linear(out_dim, config.num_phonemes)
        rnn, out_dim = _rnn_stack("word", out_dim, config.word_rnn_num_hidden, config.word_rnn_bidirectional,
                                  config.word_rnn_drop, config.word_downsample_type, config.word_downsample_len)
        self.word_layers = torch.nn.ModuleList(rnn)
        self.word_linear = torch.nn.Linear(out_dim, config.vocabulary_size)
        self.pretraining_type = config.pretraining_type
        self._plan_cache = None
        if self.is_cuda:
            self.cuda()
            _engine()          # load the CUDA engine now: the Trainer builds its optimizer right after the model (training.py:19)

    # -- execution ------------------------------------------------------------------------------
    @property
    def _plan(self):
        if self._plan_cache is None:
            self._plan_cache = _engine().Plan(self.phoneme_layers, self.word_layers)
        return self._plan_cache

    def _on_gpu(self):
        self.is_cuda = next(self.parameters()).is_cuda
        return self.is_cuda

    def _phoneme_features(self, x, with_word=True):
        """-> (phoneme-module output, carry for _word_features: what the CUDA engine prepared for the word module in the same pass)"""
        if self._on_gpu():
            return _engine().phoneme_features(self, x.cuda(), with_word)
        return _run(self.phoneme_layers, x.unsqueeze(1)), None

    def _word_features(self, ph, carry=None):
        if self.is_cuda:
            return _engine().word_features(self, ph, carry)
        return _run(self.word_layers, ph)

    def forward(self, x, y_phoneme, y_word):
        """x (B,T) float; y_phoneme (B,T') long; y_word (B,T'') long, -1 = ignore.
        Returns (phoneme_loss, word_loss, phoneme_acc, word_acc) (models.py:291-331)."""
        if self._on_gpu():
            y_phoneme, y_word = y_phoneme.cuda(), y_word.cuda()
        out, carry = self._phoneme_features(x, with_word=self.pretraining_type != 1)
        phoneme_loss, phoneme_acc = self._frame_ce(self.phoneme_linear, out, y_phoneme)
        if self.pretraining_type == 1:          # phoneme-only pre-training: skip the word module
            return phoneme_loss, torch.tensor([0.]), phoneme_acc, torch.tensor([0.])
        out = self._word_features(out, carry)
        word_loss, word_acc = self._frame_ce(self.word_linear, out, y_word)
        return phoneme_loss, word_loss, phoneme_acc, word_acc

    def _frame_ce(self, linear, feats, y):
        """Frame-wise classification head: Linear -> cross_entropy(ignore_index=-1) + masked accuracy (models.py:308-314,
        321-329).  On CUDA: the chunked tcgen05 head that never holds the [B*T', V] logits (ops.LinearCE); on the CPU: torch ops."""
        if self.is_cuda:
            return _ops().linear_ce(feats, linear.weight, linear.bias, y)
        logits = linear(feats)
        logits = logits.reshape(-1, logits.shape[-1])
        y = y.reshape(-1)
        return F.cross_entropy(logits, y, ignore_index=-1), _masked_acc(logits, y)

    def _logits(self, linear, feats):
        if self.is_cuda:
            return _ops().LinearNT.apply(feats, linear.weight, linear.bias)
        return linear(feats)

    def compute_posteriors(self, x):
        """(phoneme_logits [B,T/640,P], word_logits [B,T/2560,V]) (models.py:333-347)."""
        ph, carry = self._phoneme_features(x)
        return self._logits(self.phoneme_linear, ph), self._logits(self.word_linear, self._word_features(ph, carry))

    def compute_features(self, x):
        """[B,T] -> [B, T/2560, 256] word-module features (models.py:349-361)."""
        ph, carry = self._phoneme_features(x)
        return self._word_features(ph, carry)


def freeze_layer(layer):
    for param in layer.parameters():
        param.requires_grad = False


def unfreeze_layer(layer):
    for param in layer.parameters():
        param.requires_grad = True


def has_params(layer):
    return sum(p.numel() for p in layer.parameters()) > 0


def is_frozen(layer):
    return not any(p.requires_grad for p in layer.parameters())


class Model(torch.nn.Module):
    """End-to-end SLU model: pretrained encoder + intent module (models.py:653-875)."""

    def __init__(self, config):
        super().__init__()
        self.is_cuda = torch.cuda.is_available()
        self.Sy_intent = config.Sy_intent
        pretrained_model = PretrainedModel(config)
        if config.pretraining_type != 0:
            path = os.path.join(config.folder, "pretraining", "model_state.pth")
            pretrained_model.load_state_dict(torch.load(path, map_location=None if self.is_cuda else "cpu"))
        self.pretrained_model = pretrained_model
        self.unfreezing_type = config.unfreezing_type
        self.unfreezing_index = config.starting_unfreezing_index
        self.intent_layers = []
        if config.pretraining_type != 0:
            self.freeze_all_layers()
        self.seq2seq = config.seq2seq
        out_dim = config.word_rnn_num_hidden[-1] * (2 if config.word_rnn_bidirectional else 1)
        if not self.seq2seq:
            self.values_per_slot = config.values_per_slot
            self.num_values_total = sum(self.values_per_slot)
            rnn, out_dim = _rnn_stack("intent", out_dim, config.intent_rnn_num_hidden, config.intent_rnn_bidirectional,
                                      config.intent_rnn_drop, config.intent_downsample_type, config.intent_downsample_len)
            rnn.append(_named(torch.nn.Linear(out_dim, self.num_values_total), "final_classifier"))
            rnn.append(_named(FinalPool(), "final_pool"))
            self.intent_layers = torch.nn.ModuleList(rnn)
        else:
            from seq2seq import Seq2SeqDecoder, Seq2SeqEncoder
            self.SOS = config.Sy_intent.index("<sos>")
            self.num_labels = len(config.Sy_intent)
            self.encoder = Seq2SeqEncoder(out_dim, config.num_intent_encoder_layers, config.intent_encoder_dim)
            self.decoder = Seq2SeqDecoder(self.num_labels, config.num_intent_decoder_layers, config.intent_encoder_dim,
                                          config.intent_decoder_dim, config.intent_decoder_key_dim,
                                          config.intent_decoder_value_dim, self.SOS)
        if self.is_cuda:
            self.cuda()

    # -- engine hooks -----------------------------------------------------------------------------
    @property
    def _intent_rnns(self):
        eng = _engine()
        layers = list(self.intent_layers)
        return [(l, layers[i + 2].p, eng.downsample_factor(layers[i + 3])) for i, l in enumerate(layers)
                if isinstance(l, torch.nn.GRU)]

    @property
    def _final_classifier(self):
        return self.intent_layers[-2]

    def _intent_logits(self, x):
        """(B,T) waveform -> (B, num_values_total) logits."""
        feats = self.pretrained_model.compute_features(x)
        if self.pretrained_model.is_cuda:
            return _engine().intent_logits(self, feats)
        return _run(self.intent_layers, feats)

    def _slot_argmax(self, logits):
        pred, start = [], 0
        for n in self.values_per_slot:
            pred.append(logits[:, start:start + n].max(1)[1])
            start += n
        return torch.stack(pred, dim=1)

    # -- reference surface --------------------------------------------------------------------------
    def one_hot_to_string(self, input, S):
        """input (T, |S|) one-hot rows -> string over alphabet S (models.py:730-736)."""
        return "".join([S[c] for c in input.max(dim=1)[1]]).lstrip("<sos>").rstrip("<eos>")

    def freeze_all_layers(self):
        for layer in list(self.pretrained_model.phoneme_layers) + list(self.pretrained_model.word_layers):
            freeze_layer(layer)

    def print_frozen(self):
        for layer in list(self.pretrained_model.phoneme_layers) + list(self.pretrained_model.word_layers):
            if has_params(layer):
                print(layer.name + ": " + ("frozen" if is_frozen(layer) else "unfrozen"))

    def unfreeze_one_layer(self):
        """ULMFiT-style gradual unfreezing (models.py:754-795): type 1 walks the word module from the
        top, type 2 continues into the phoneme module; every call un-freezes up to and including
        the `unfreezing_index`-th parameterised layer counted from the top."""
        if self.unfreezing_type == 0:
            return
        stacks = [self.pretrained_model.word_layers]
        if self.unfreezing_type == 2:
            stacks.append(self.pretrained_model.phoneme_layers)
        trainable = 0
        for stack in stacks:
            for layer in reversed(list(stack)):
                unfreeze_layer(layer)
                if has_params(layer):
                    trainable += 1
                if trainable == self.unfreezing_index:
                    self.unfreezing_index += 1
                    return

    def forward(self, x, y_intent):
        """x (B,T); y_intent (B, n_slots) long [or (B,U,|S|) one-hot for seq2seq].  Returns (loss, acc)."""
        on_gpu = self.pretrained_model._on_gpu()
        if on_gpu and self.training and torch.is_grad_enabled():
            # the training step as captured CUDA graphs once its shape has been seen a few times (engine.graphed_train_step)
            out = _engine().graphed_train_step(self, x, y_intent, self._forward_eager)
            if out is not None:
                return out
        return self._forward_eager(x, y_intent)

    def _forward_eager(self, x, y_intent):
        on_gpu = self.pretrained_model._on_gpu()
        if on_gpu:
            y_intent = y_intent.cuda()
        if self.seq2seq:
            out = self.encoder(self.pretrained_model.compute_features(x))
            log_probs = self.decoder(out, y_intent)
            return -log_probs.mean(), torch.tensor([0.])
        if on_gpu:                                  # fused head: Linear + max over time + slot CE + accuracy, one kernel
            fused = _engine().intent_loss_acc(self, x, y_intent)
            if fused is not None:
                return fused
        logits = self._intent_logits(x)
        loss, start = 0., 0
        for slot, n in enumerate(self.values_per_slot):
            loss = loss + F.cross_entropy(logits[:, start:start + n], y_intent[:, slot])
            start += n
        acc = (self._slot_argmax(logits) == y_intent).prod(1).float().mean()      # all slots must be right
        return loss, acc

    def predict_intents(self, x):
        if self.seq2seq:
            out = self.encoder(self.pretrained_model.compute_features(x))
            return self.decoder.infer(out, self.Sy_intent, B=4)
        logits = self._intent_logits(x)
        return logits, self._slot_argmax(logits)

    def decode_intents(self, x):
        _, predicted = self.predict_intents(x)
        if self.seq2seq:   # predicted: (beam, batch, U, num_labels); best hypothesis is beam 0
            return [self.one_hot_to_string(predicted[0, i], self.Sy_intent) for i in range(predicted.shape[1])]
        intents = []
        for row in predicted:
            intents.append([value for idx, slot in enumerate(self.Sy_intent)
                            for value, code in self.Sy_intent[slot].items() if row[idx].item() == code])
        return intents
