"""Pin both CPU oracles (oracle/torch_ref.py, oracle/slu_oracle.py) against golden vectors
produced by the real reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import slu_oracle as N
from oracle import torch_ref as R
from util import ckpt_params, golden, rel_err, load_test_wav

TOL = 2e-5   # fp32 re-association noise only


def test_sinc_filters_match_reference():
    g = golden("golden_sinc_filters.npz")
    p = ckpt_params()
    w_mel = R.sinc_filters(torch.from_numpy(g["mel_b1"]), torch.from_numpy(g["mel_band"]))
    w_ck = R.sinc_filters(p[R.P + "phoneme_layers.0.filt_b1"], p[R.P + "phoneme_layers.0.filt_band"])
    assert rel_err(w_mel, g["mel_init"]) < 1e-6
    assert rel_err(w_ck, g["ckpt"]) < 1e-6
    assert rel_err(N.sinc_filters(g["mel_b1"], g["mel_band"]), g["mel_init"]) < 2e-5
    # mel init parameters themselves (models.py:56-68)
    sp = R.synthetic_params()
    assert np.array_equal(sp[R.P + "phoneme_layers.0.filt_b1"].numpy(), g["mel_b1"])
    assert np.array_equal(sp[R.P + "phoneme_layers.0.filt_band"].numpy(), g["mel_band"])


def test_testwav_known_answer_torch_oracle():
    g = golden("golden_testwav.npz")
    with torch.no_grad():
        logits, acts = R.intent_logits(load_test_wav(), ckpt_params(), return_all=True)
    assert rel_err(logits, g["logits"]) < TOL
    _, _, pred = R.intent_loss_acc(logits, torch.from_numpy(g["pred"]))
    assert pred.tolist() == [[1, 2, 1]] == g["pred"].tolist()          # README.md:42 {activate, lights, kitchen}
    for k in ("sinc", "conv1", "conv2", "gru0", "gru1", "gru2", "gru3", "gru4", "gru0_raw"):
        assert acts[k].shape == g[k].shape, k
        assert rel_err(acts[k], g[k]) < TOL, k
    assert [acts["gru%d" % i].shape[1] for i in range(5)] == [180, 90, 45, 23, 23]


def test_testwav_known_answer_numpy_oracle():
    g = golden("golden_testwav.npz")
    p = {k: v.numpy() for k, v in ckpt_params().items()}
    logits, acts = N.intent_logits(load_test_wav().numpy(), p, return_all=True)
    assert rel_err(logits, g["logits"]) < 1e-4
    assert N.predict(logits).tolist() == [[1, 2, 1]]
    for k in ("sinc", "conv1", "conv2", "gru0", "gru3", "features"):
        assert rel_err(acts[k], g[k]) < 1e-4, k


@pytest.mark.parametrize("tag", ["small", "ragged", "odd"])
def test_synthetic_loss_logits_grads(tag):
    g = golden("golden_synth_%s.npz" % tag)
    p = {k: v.clone().requires_grad_(True) for k, v in R.synthetic_params(seed=int(g["pseed"])).items()}
    x, y = R.synthetic_batch(int(g["B"]), int(g["T"]), seed=int(g["bseed"]))
    loss, acc, logits = R.slu_forward(x, y, p)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    assert acc.item() == float(g["acc"])
    assert rel_err(logits, g["logits"]) < TOL
    n_checked = 0
    for k in p:
        if "g/" + k not in g.files:
            continue
        grad = p[k].grad.flatten()
        sub = grad if grad.numel() <= 30000 else grad[::7]
        assert rel_err(sub, g["g/" + k]) < 1e-4, k
        assert abs(grad.double().norm().item() - float(g["gl2/" + k])) <= 1e-4 * float(g["gl2/" + k]) + 1e-12, k
        n_checked += 1
    assert n_checked == 48          # 2 sinc + 4 conv + 5*8 GRU + 2 classifier
    # numpy oracle, forward
    pn = {k: v.detach().numpy() for k, v in p.items()}
    assert rel_err(N.intent_logits(x.numpy(), pn), g["logits"]) < 1e-4


def test_explicit_dropout_masks():
    g = golden("golden_synth_dropout.npz")
    p = {k: v.clone().requires_grad_(True) for k, v in R.synthetic_params(seed=int(g["pseed"])).items()}
    x, y = R.synthetic_batch(int(g["B"]), int(g["T"]), seed=int(g["bseed"]))
    rs = np.random.RandomState(int(g["mask_seed"]))
    T = (int(g["T"]) - 1) // 80 + 1
    T = (T + 1) // 2
    masks = []
    for li in range(5):
        masks.append(torch.from_numpy((rs.uniform(size=(int(g["B"]), T, 256)) >= 0.5).astype(np.float32) * 2.0))
        if li < 4:
            T = (T + 1) // 2
    loss, _, _ = R.slu_forward(x, y, p, drop_masks=masks)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    for k in p:
        if "gl2/" + k in g.files:
            assert abs(p[k].grad.double().norm().item() - float(g["gl2/" + k])) < 2e-4 * float(g["gl2/" + k]) + 1e-9, k


def test_asr_heads():
    g = golden("golden_asr.npz")
    p = R.synthetic_params(seed=12, asr=True)
    pa = {k[len(R.P):]: v.clone().requires_grad_(True) for k, v in p.items() if k.startswith(R.P)}
    x, _ = R.synthetic_batch(int(g["B"]), int(g["T"]), seed=13)
    pl, wl, pacc, wacc = R.asr_forward(x, torch.from_numpy(g["y_phoneme"]), torch.from_numpy(g["y_word"]), pa)
    assert abs(pl.item() - float(g["phoneme_loss"])) < 1e-5 * float(g["phoneme_loss"])
    assert abs(wl.item() - float(g["word_loss"])) < 1e-5 * float(g["word_loss"])
    assert abs(pacc.item() - float(g["phoneme_acc"])) < 1e-6 and abs(wacc.item() - float(g["word_acc"])) < 1e-6
    (pl + wl).backward()
    for k, v in pa.items():
        if "gl2/" + k in g.files:
            assert abs(v.grad.double().norm().item() - float(g["gl2/" + k])) < 2e-4 * float(g["gl2/" + k]) + 1e-9, k


def test_downsample_odd_tail():
    x = torch.arange(5.).view(1, 5, 1)
    assert R.downsample(x, "avg", 2).flatten().tolist() == [0.5, 2.5, 4.0]      # SURVEY.md K8 [probed]
    assert N.downsample_avg2(x.numpy()).flatten().tolist() == [0.5, 2.5, 4.0]
