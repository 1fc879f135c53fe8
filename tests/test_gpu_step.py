"""GPU tests of the train-step glue: the fused Adam kernel behind torch.optim.Adam, the flat gradient
arena the weight-gradient kernels write into, and the drop-in details around them."""
import importlib

import numpy as np
import pytest
import torch

import models
from oracle import torch_ref as R
from util import make_config, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    p = importlib.import_module("end-to-end-slu_b200")
    p._lib.load()
    return p


def test_fused_adam_matches_torch_adam(pkg):
    """slu_adam_multi against the stock torch.optim.Adam: fp32 + fp64 tensors, ragged sizes (vector and tail paths), more than
    64 tensors (two launches), a parameter that starts receiving gradients later (its own bias-correction step count), weight decay."""
    from torch.optim.adam import Adam as StockAdam
    rs = np.random.RandomState(0)
    shapes = [(384, 256), (768,), (5, 7, 3), (1,), (4096,), (4097,), (60, 80, 5), (10000, 16)] + [(33,)] * 70
    mk = lambda: [torch.nn.Parameter(torch.from_numpy(rs.standard_normal(s).astype(np.float32)).cuda()) for s in shapes] + \
                 [torch.nn.Parameter(torch.from_numpy(rs.standard_normal(80) * 1e-2).cuda())]          # fp64, like filt_b1
    rs = np.random.RandomState(0); pa = mk()
    rs = np.random.RandomState(0); pb = mk()
    for wd in (0.0, 0.01):
        oa = pkg.optim.FusedAdam(pa, lr=1e-3, weight_decay=wd)
        ob = StockAdam(pb, lr=1e-3, weight_decay=wd)
        calls0 = pkg._lib.stats["calls"]
        for it in range(5):
            g = np.random.RandomState(100 + it)
            for i, (a, b) in enumerate(zip(pa, pb)):
                if i == 2 and it < 2:           # "frozen" for the first two steps
                    a.grad = b.grad = None
                    continue
                gr = torch.from_numpy(g.standard_normal(tuple(a.shape))).to(a.dtype).cuda()
                a.grad, b.grad = gr.clone(), gr.clone()
            oa.step(); ob.step()
        assert pkg._lib.stats["calls"] - calls0 == 5              # one C-ABI call per step (79 tensors -> 2 launches inside it)
        for a, b in zip(pa, pb):
            assert a.dtype == b.dtype
            assert rel_err(a.detach().cpu(), b.detach().cpu()) < 2e-6
        sa, sb = oa.state_dict()["state"], ob.state_dict()["state"]
        assert float(sa[2]["step"]) == float(sb[2]["step"]) == 3.0 and float(sa[0]["step"]) == 5.0
        assert rel_err(sa[0]["exp_avg_sq"].cpu(), sb[0]["exp_avg_sq"].cpu()) < 2e-6


def test_trainer_constructs_the_fused_adam_and_survives_the_cpu_hop(pkg):
    m = models.Model(make_config())                       # construction on a CUDA box installs the subclass
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)       # what training.py:19 executes
    from torch.optim.adam import Adam as StockAdam
    assert isinstance(opt, pkg.optim.FusedAdam) and isinstance(opt, StockAdam)
    x, y = R.synthetic_batch(2, 8000, seed=1)
    for _ in range(2):
        loss, _ = m(x, y); opt.zero_grad(); loss.backward(); opt.step()
    m.cpu(); m.is_cuda = False                            # Trainer.test() hop (training.py:150/166): validation forward on the CPU
    with torch.no_grad():
        m(x, y)
    m.cuda(); m.is_cuda = True                            # the parameters come back at new addresses: the cached table must follow
    loss2, _ = m(x, y); opt.zero_grad(); loss2.backward(); opt.step()
    assert torch.isfinite(loss2) and float(opt.state_dict()["state"][len(list(m.parameters())) - 1]["step"]) == 3.0
    ref = StockAdam([torch.nn.Parameter(torch.zeros(3))], lr=1e-3)                   # CPU parameters: the stock step runs
    cpu_opt = torch.optim.Adam([torch.nn.Parameter(torch.ones(3))], lr=1e-3)
    cpu_opt.param_groups[0]["params"][0].grad = torch.ones(3)
    cpu_opt.step()
    assert type(ref) is not type(cpu_opt) and abs(cpu_opt.param_groups[0]["params"][0][0].item() - 0.999) < 1e-6


def test_gradients_are_views_of_one_arena_and_accumulate_like_autograd(pkg):
    p = R.synthetic_params(seed=3)
    m = models.Model(make_config()).eval()
    sd = m.state_dict(); sd.update({k: v for k, v in p.items() if k in sd}); m.load_state_dict(sd)
    for q in m.parameters():
        q.requires_grad = True
    x, y = R.synthetic_batch(3, 8000, seed=4)
    loss, _ = m(x, y)
    loss.backward(retain_graph=True)
    grads = [q.grad for q in m.parameters() if q.grad is not None]
    arena = pkg.grads.find(grads[0])
    assert arena is not None and all(arena.owns(g) for g in grads) and len(grads) == 48
    assert sum(g.dtype == torch.float64 for g in grads) == 2
    first = {k: q.grad.clone() for k, q in m.named_parameters() if q.grad is not None}
    loss.backward()                                       # second pass through the retained graph: .grad must double, not alias
    for k, q in m.named_parameters():
        if q.grad is not None:
            assert rel_err(q.grad, 2 * first[k]) < 1e-5, k
    # a fresh forward/backward after zero_grad is independent of the first arena
    m.zero_grad()
    loss, _ = m(x, y); loss.backward()
    for k, q in m.named_parameters():
        if q.grad is not None:
            assert rel_err(q.grad, first[k]) < 1e-5, k


def test_out_of_range_label_poisons_the_loss(pkg):
    m = models.Model(make_config()).eval()
    x, y = R.synthetic_batch(2, 8000, seed=1)
    y = y.clone(); y[1, 0] = 6                            # slot 0 has 6 values
    loss, acc = m(x, y)
    assert torch.isnan(loss)


def test_predict_intents_logits_are_differentiable_on_gpu(pkg):
    p = R.synthetic_params(seed=5)
    m = models.Model(make_config()).eval()
    sd = m.state_dict(); sd.update({k: v for k, v in p.items() if k in sd}); m.load_state_dict(sd)
    for q in m.parameters():
        q.requires_grad = True
    x, _ = R.synthetic_batch(2, 8000, seed=6)
    logits, _ = m.predict_intents(x)
    w = torch.linspace(-1, 1, 24, device="cuda")
    (logits * w).sum().backward()
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    (R.intent_logits(x, pr) * w.cpu()).sum().backward()
    for k in ("intent_layers.4.weight", "intent_layers.4.bias", "intent_layers.0.weight_hh_l0", "pretrained_model.phoneme_layers.5.weight"):
        assert rel_err(dict(m.named_parameters())[k].grad.cpu(), pr[k].grad) < 5e-3, k


def test_gru_size_limit_is_a_readable_error(pkg):
    gx = torch.empty(1, device="cuda")
    with pytest.raises(RuntimeError, match="split the batch"):
        pkg._lib.call("slu_gru_fwd_tc", gx.data_ptr(), gx.data_ptr(), gx.data_ptr(), None, 0.0, 0, None, 4096, 1024, 1, gx.data_ptr(),
                      gx.data_ptr(), None, pkg._lib.stream())


def test_prefetcher_keeps_pinned_sources_alive_until_the_copy_has_run(pkg):
    """batches from DataLoader(pin_memory=True) are dropped by the caller right after staging; the lean path must hold
    them until the DMA is done.  The sources here are freed (and their pinned blocks overwritten) as soon as the iterator moves on."""
    def batches():
        for i in range(12):
            yield (torch.full((64, 4000), float(i)).pin_memory(), torch.full((64, 3), i, dtype=torch.int64).pin_memory())
    pf = pkg.loader.DevicePrefetcher(batches())
    seen = []
    for x, y in pf:
        torch.cuda._sleep(2_000_000)                      # keep the compute stream (which the copies are ordered after) busy
        seen.append((x, y))
        if len(seen) > 1:
            a, b = seen.pop(0)
            assert a.min().item() == a.max().item() == float(b[0, 0].item())
    assert not pf._keep


def _train_model(cfg_over=None, seq2seq=False):
    from util import make_config
    if seq2seq:
        cfg = make_config("seq2seq")
        cfg.Sy_intent = ["<sos>"] + list("abcdefghij {}:'\",") + ["<eos>"]
    else:
        cfg = make_config(**(cfg_over or {}))
    torch.manual_seed(0)
    m = models.Model(cfg).train()
    for q in m.parameters():
        q.requires_grad = True
    return m, cfg


@pytest.mark.parametrize("seq2seq", [False, True])
def test_graphed_train_step_equals_the_eager_step(pkg, monkeypatch, seq2seq):
    """The training forward/backward captured as two CUDA graphs (engine.graphed_train_step) against eager execution, dropout
    probabilities set to 0 so both are deterministic: same loss, same gradients for every parameter, for new inputs on every
    replay, with the unchanged zero_grad / backward / optimizer.step sequence of the Trainer; dL/dloss != 1 is honoured."""
    eng = pkg.engine
    over = dict(phone_rnn_drop=[0.0, 0.0], word_rnn_drop=[0.0, 0.0], intent_rnn_drop=[0.0])
    m, cfg = _train_model(over, seq2seq)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    rs = np.random.RandomState(0)
    S = len(cfg.Sy_intent) if seq2seq else 0

    def batch(i):
        x = torch.from_numpy((0.1 * rs.standard_normal((6, 8000))).astype(np.float32))
        if seq2seq:
            idx = rs.randint(1, S - 1, size=(6, 7)); idx[:, 0] = 0; idx[:, -1] = S - 1
            return x, torch.nn.functional.one_hot(torch.from_numpy(idx), S).float()
        return x, torch.from_numpy(np.stack([rs.randint(0, v, size=6) for v in (6, 14, 4)], 1))
    batches = [batch(i) for i in range(6)]

    def run(graph):
        monkeypatch.setattr(eng, "STEP_GRAPH", graph)
        m.__dict__.pop("_step_graphs", None)
        out = []
        loss = None                                    # like the Trainer, the previous step's loss is still referenced during the next forward
        for i, (x, y) in enumerate(batches):
            m.zero_grad()
            loss, _ = m(x, y)
            (loss * (2.0 if i == 4 else 1.0)).backward()
            out.append((loss.item(), {k: q.grad.detach().clone() for k, q in m.named_parameters() if q.grad is not None}))
        return out
    eager = run(False)
    graphed = run(True)
    assert any(isinstance(v, eng._StepGraph) for v in m._step_graphs.values())          # the third occurrence was captured
    for (le, ge), (lg, gg) in zip(eager, graphed):
        assert abs(le - lg) < 1e-6 * abs(le)
        assert set(ge) == set(gg)
        for k in ge:
            assert ge[k].dtype == gg[k].dtype and rel_err(gg[k], ge[k]) < 2e-5, k
    # stale backward is refused, gradient accumulation falls back to eager semantics
    m.zero_grad()
    l_a, _ = m(*batches[0]); l_b, _ = m(*batches[1])
    with pytest.raises(RuntimeError, match="later forward"):
        l_a.backward()
    m.zero_grad()
    l1, _ = m(*batches[0]); l1.backward()
    l2, _ = m(*batches[0]); l2.backward()                      # .grad is not None -> eager, accumulates
    for k, q in m.named_parameters():
        if q.grad is not None:
            assert rel_err(q.grad, 2 * eager[0][1][k]) < 2e-5, k


def test_graphed_train_step_draws_new_dropout_masks_every_replay(pkg):
    m, _ = _train_model()
    x, y = R.synthetic_batch(6, 8000, seed=3)
    torch.manual_seed(1)
    losses = []
    for i in range(7):
        m.zero_grad()
        loss, _ = m(x, y)
        loss.backward()
        losses.append(loss.item())
    assert any(isinstance(v, pkg.engine._StepGraph) for v in m._step_graphs.values())
    assert len(set(losses[2:])) == len(losses[2:])              # same input, different masks on every replay
    m.eval()
    with torch.no_grad():
        l_eval = m(x, y)[0].item()
    assert abs(np.mean(losses) - l_eval) < 0.5 * abs(l_eval)    # and still the same model
