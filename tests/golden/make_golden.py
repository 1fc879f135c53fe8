"""Generate the golden fixtures in this directory by running the REAL reference
(/root/reference, imported unmodified; only `soundfile`/`textgrid` are stubbed because they are
not installed and are not used on this path).  Run in the build container only:

    python tests/golden/make_golden.py

The GPU box has no /root/reference; tests there read the committed .npz files.
Inputs/weights for the synthetic cases come from oracle.torch_ref.synthetic_params /
synthetic_batch (numpy legacy RandomState => regenerated bit-identically anywhere), so only
the reference's OUTPUTS are stored.  The trained-checkpoint case stores the SLU-path weights of
experiments/no_unfreezing/training/model_state.pth (everything except the unused 10 MB word head).
"""
import os, sys, types, wave, tempfile
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SLU_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
for m in ("soundfile", "textgrid"):
    sys.modules[m] = types.ModuleType(m)
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)
work = tempfile.mkdtemp()
os.symlink(os.path.join(REF, "experiments"), os.path.join(work, "experiments"))
os.chdir(work)
import data as ref_data            # noqa: E402  (reference data.py)
import models as ref_models        # noqa: E402  (reference models.py)
from oracle import torch_ref as R  # noqa: E402

assert ref_models.__file__.startswith(REF)
torch.manual_seed(0)
cfg = ref_data.read_config("experiments/no_unfreezing.cfg")
cfg.Sy_intent = {"action": {}, "object": {}, "location": {}}
cfg.values_per_slot = [6, 14, 4]
cfg.num_phonemes = 42


ONLY = set(filter(None, os.environ.get("GOLDEN_ONLY", "").split(",")))     # e.g. GOLDEN_ONLY=seq2seq regenerates one fixture


def want(tag):
    return not ONLY or tag in ONLY


def save(name, **arrs):
    if not want(name.split(".")[0].replace("golden_", "").replace("synth_", "").split("_")[0]) and not want(name):
        return
    out = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()}
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, {k: v.shape for k, v in out.items() if v.size > 1 and len(out) < 20})


def layer_acts(model, x):
    """Run the reference layer loops (models.py:354-359, 807-808) recording named activations."""
    acts = {}
    out = x.unsqueeze(1)
    for layer in list(model.pretrained_model.phoneme_layers) + list(model.pretrained_model.word_layers) + list(model.intent_layers):
        out = layer(out)
        if torch.is_tensor(out):
            acts[layer.name] = out
    return acts


# ---- 1. trained checkpoint + test.wav (README.md:26-42 known answer) -------------------------
model = ref_models.Model(cfg).eval()
sd = torch.load("experiments/no_unfreezing/training/model_state.pth", map_location="cpu")
print(model.load_state_dict(sd))
w = wave.open(os.path.join(REF, "test.wav"))
pcm = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16)
x = torch.tensor(pcm.astype(np.float32) / 32768).unsqueeze(0)
logits, pred = model.predict_intents(x)
acts = layer_acts(model, x)
save("ckpt_no_unfreezing_slu.npz", **{k: v for k, v in sd.items() if "word_linear" not in k})
save("test_wav.npz", pcm=pcm, fs=16000)
save("golden_testwav.npz", logits=logits, pred=pred, features=model.pretrained_model.compute_features(x),
     sinc=acts["dropout0"], conv1=acts["dropout1"], conv2=acts["dropout2"],
     gru0=acts["phone_downsample0"], gru1=acts["phone_downsample1"], gru2=acts["word_downsample0"],
     gru3=acts["word_downsample1"], gru4=acts["intent_downsample0"], gru0_raw=acts["phone_rnn_select0"])

# ---- 2. sinc filter banks -----------------------------------------------------------------------
def ref_filters(layer):
    captured = {}
    orig = torch.nn.functional.conv1d
    def spy(inp, weight, **kw):
        captured["w"] = weight.detach().clone()
        return orig(inp, weight, **kw)
    torch.nn.functional.conv1d = spy
    try:
        layer(torch.zeros(1, 1, 800))
    finally:
        torch.nn.functional.conv1d = orig
    return captured["w"].view(80, 401)
fresh = ref_models.SincLayer(80, 401, 16000, stride=80, padding=200)
save("golden_sinc_filters.npz", mel_init=ref_filters(fresh), ckpt=ref_filters(model.pretrained_model.phoneme_layers[0]),
     mel_b1=fresh.filt_b1, mel_band=fresh.filt_band)

# ---- 3. synthetic params: loss / logits / all gradients (train graph, dropout off) --------------
def load_synth(model, seed):
    p = R.synthetic_params(seed=seed)
    sd2 = model.state_dict()
    for k, v in p.items():
        sd2[k] = v.clone()
    model.load_state_dict(sd2)
    for q in model.parameters():
        q.requires_grad = True
    return p

BIG = 30000
for tag, (B, T, pseed, bseed) in {"small": (3, 8000, 3, 5), "ragged": (2, 1234, 4, 6), "odd": (2, 12345, 7, 8)}.items():
    p = load_synth(model, pseed)
    xb, yb = R.synthetic_batch(B, T, seed=bseed)
    model.zero_grad()
    loss, acc = model(xb, yb)
    loss.backward()
    lg, pr = model.predict_intents(xb)
    out = {"B": B, "T": T, "pseed": pseed, "bseed": bseed, "loss": loss, "acc": acc, "logits": lg, "pred": pr,
           "features": model.pretrained_model.compute_features(xb)}
    named = dict(model.named_parameters())
    for k in p:
        g = named[k].grad
        if g is None:
            continue
        g = g.flatten()
        out["gsum/" + k] = g.double().sum(); out["gl2/" + k] = g.double().norm()
        out["g/" + k] = g if g.numel() <= BIG else g[::7]
    save("golden_synth_%s.npz" % tag, **out)

# ---- 4. train-mode dropout with explicit masks: reference nn.Dropout replaced by mask multiply ---
#      (checks the mask plumbing / ordering dropout -> downsample, models.py:246-253)
p = load_synth(model, 9)
xb, yb = R.synthetic_batch(2, 6400, seed=10)
rs = np.random.RandomState(11)
masks = []
class MaskDrop(torch.nn.Module):
    def __init__(self, name): super().__init__(); self.name = name; self.mask = None
    def forward(self, x):
        self.mask = torch.from_numpy((rs.uniform(size=tuple(x.shape)) >= 0.5).astype(np.float32) * 2.0); masks.append(self.mask)
        return x * self.mask
def swap(layers):
    for i, l in enumerate(layers):
        if isinstance(l, torch.nn.Dropout) and l.p > 0:
            layers[i] = MaskDrop(l.name)
swap(model.pretrained_model.phoneme_layers); swap(model.pretrained_model.word_layers); swap(model.intent_layers)
model.train(); model.zero_grad()
loss, acc = model(xb, yb); loss.backward()
named = dict(model.named_parameters())
save("golden_synth_dropout.npz", loss=loss, mask_seed=11, pseed=9, bseed=10, B=2, T=6400,
     **{"gl2/" + k: named[k].grad.double().norm() for k in p if named[k].grad is not None},
     **{"gsum/" + k: named[k].grad.double().sum() for k in p if named[k].grad is not None})

# ---- 5. ASR pre-training forward (PretrainedModel.forward, models.py:291-331) -------------------
cfg.pretraining_type = 2
pm = ref_models.PretrainedModel(cfg).eval()
p = R.synthetic_params(seed=12, asr=True)
sd3 = pm.state_dict()
for k, v in p.items():
    kk = k[len(R.P):] if k.startswith(R.P) else None
    if kk in sd3:
        sd3[kk] = v.clone()
pm.load_state_dict(sd3)
B, T = 2, 5120
xb, _ = R.synthetic_batch(B, T, seed=13)
rs = np.random.RandomState(14)
yp = torch.from_numpy(rs.randint(-1, 42, size=(B, T // 640)).astype(np.int64))
yw = torch.from_numpy(rs.randint(-1, 10000, size=(B, T // 2560)).astype(np.int64))
pl, wl, pa, wa = pm(xb, yp, yw)
(pl + wl).backward()
ph_post, w_post = pm.compute_posteriors(xb)
named = dict(pm.named_parameters())
save("golden_asr.npz", phoneme_loss=pl, word_loss=wl, phoneme_acc=pa, word_acc=wa, y_phoneme=yp, y_word=yw,
     phoneme_logits=ph_post, word_logits_sub=w_post[..., ::50], B=B, T=T,
     **{"gl2/" + k: v.grad.double().norm() for k, v in named.items() if v.grad is not None})

# ---- 6. seq2seq intent module (Model.forward / Seq2SeqDecoder.forward + infer, models.py:381-651) -------------
#      Weights: torch.manual_seed(21) default init of the reference Model itself -- the same seed reproduces them in
#      models.Model on any box with this torch build (tests/test_models_cpu.py checks the init order is identical), so
#      only outputs and a parameter checksum are stored.
if want("seq2seq"):
    import importlib
    cfgmod = importlib.import_module("end-to-end-slu_b200.config")
    scfg = cfgmod.read_config(os.path.join(ROOT, "configs", "seq2seq.cfg"))
    scfg.pretraining_type = 0
    scfg.num_phonemes = 42
    scfg.Sy_intent = ["<sos>"] + list("abcdefghij {}:'\",") + ["<eos>"]
    torch.manual_seed(21)
    sm = ref_models.Model(scfg); sm.cpu(); sm.is_cuda = False
    sm.eval()
    S, U, Bs, Ts = len(scfg.Sy_intent), 7, 3, 8000
    xb, _ = R.synthetic_batch(Bs, Ts, seed=31)
    rs = np.random.RandomState(32)
    idx = rs.randint(1, S - 1, size=(Bs, U)); idx[:, 0] = 0; idx[:, -1] = S - 1
    yb = torch.nn.functional.one_hot(torch.from_numpy(idx.astype(np.int64)), S).float()
    for q in sm.parameters():
        q.requires_grad = True
    loss, _ = sm(xb, yb)
    loss.backward()
    enc = sm.encoder(sm.pretrained_model.compute_features(xb))
    log_p = sm.decoder(enc, yb)
    scores, beam = sm.decoder.infer(enc, scfg.Sy_intent, B=4, y_lengths=[6])
    named = dict(sm.named_parameters())
    save("golden_seq2seq.npz", loss=loss, log_p=log_p, enc=enc, beam_scores=scores, beam_ids=beam.argmax(-1), idx=idx, B=Bs, T=Ts, U=U,
         seed=21, bseed=31, param_abs_sum=sum(v.detach().double().abs().sum() for v in sm.state_dict().values()),
         **{"gl2/" + k: v.grad.double().norm() for k, v in named.items() if v.grad is not None})
print("done")
