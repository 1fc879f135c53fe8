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
        pkg._lib.call("slu_gru_fwd_tc", gx.data_ptr(), gx.data_ptr(), gx.data_ptr(), None, 0.0, 0, 4096, 1024, 1, gx.data_ptr(),
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
