"""2-GPU NCCL test of the data-parallel gradient exchange (dp.allreduce_grads incl. the fp64 sinc gradients carried inside the
single fp32 bucket): sharding the minibatch over 2 ranks reproduces the single-GPU gradients of the full batch, and one
optimizer step leaves both ranks with identical parameters.  Needs 2 CUDA devices (`gpurun --gpus 2`); skipped otherwise."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _model(device):
    import models
    from oracle import torch_ref as R
    from util import make_config
    torch.cuda.set_device(device)
    torch.manual_seed(0)
    m = models.Model(make_config()).eval()
    sd = m.state_dict(); sd.update({k: v for k, v in R.synthetic_params(seed=1).items() if k in sd}); m.load_state_dict(sd)
    for q in m.parameters():
        q.requires_grad = True
    return m


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import importlib
    from oracle import torch_ref as R
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    dp = importlib.import_module("end-to-end-slu_b200.dp")
    dp.install()
    m = _model(rank)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    x, y = R.synthetic_batch(8, 16000, seed=2)
    shard = slice(rank * 4, rank * 4 + 4)
    loss, _ = m(x[shard], y[shard])
    opt.zero_grad(); loss.backward()
    params = [p for p in m.parameters()]
    dp.allreduce_grads(params)                               # what the optimizer pre-step hook runs
    grads = {k: p.grad.detach().cpu().clone() for k, p in m.named_parameters() if p.grad is not None}
    calls = dp.stats["allreduce_calls"]
    opt.step()                                               # hook fires again on already-averaged grads: mean of equal values
    torch.cuda.synchronize()
    torch.save({"grads": grads, "calls": calls, "params": {k: p.detach().cpu() for k, p in m.named_parameters()}}, out % rank)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 CUDA devices")
def test_two_gpu_nccl_allreduce_equals_full_batch(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import torch_ref as R
    from util import rel_err
    out = str(tmp_path / "g%d.pt")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r0, r1 = torch.load(out % 0), torch.load(out % 1)
    assert r0["calls"] == 1
    m = _model(0)
    x, y = R.synthetic_batch(8, 16000, seed=2)
    loss, _ = m(x, y); loss.backward()
    n = 0
    for k, p in m.named_parameters():
        if p.grad is None:
            assert k not in r0["grads"]
            continue
        assert r0["grads"][k].dtype == p.grad.dtype                       # fp64 sinc grads stay fp64
        assert rel_err(r0["grads"][k], p.grad.cpu()) < 2e-4, (k, rel_err(r0["grads"][k], p.grad.cpu()))
        assert torch.equal(r0["grads"][k], r1["grads"][k]), k             # both ranks hold the same averaged gradient
        n += 1
    assert n == 48
    for k in r0["params"]:
        assert torch.equal(r0["params"][k], r1["params"][k]), k           # replicas stay in lock step after the Adam step
