"""world_size-2 gloo test (CPU): sharding the minibatch over 2 ranks + the single flat all-reduce
reproduces the single-process gradients of the full batch.  Uses the CPU execution
path of models.py, so it runs without a GPU."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import importlib
    import models
    from oracle import torch_ref as R
    from util import make_config
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dp = importlib.import_module("end-to-end-slu_b200.dp")
    dp.install()
    torch.manual_seed(0)
    m = models.Model(make_config()).cpu().eval(); m.is_cuda = False
    sd = m.state_dict(); sd.update({k: v for k, v in R.synthetic_params(seed=1).items() if k in sd}); m.load_state_dict(sd)
    opt = torch.optim.SGD(m.parameters(), lr=0.0)            # lr 0: the hook runs, weights stay put
    x, y = R.synthetic_batch(4, 4000, seed=2)
    shard = slice(rank * 2, rank * 2 + 2)
    loss, _ = m(x[shard], y[shard])
    opt.zero_grad(); loss.backward(); opt.step()
    if rank == 0:
        torch.save({k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}, out)
        assert dp.stats["allreduce_calls"] == 1
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_full_batch(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import models
    from oracle import torch_ref as R
    from util import make_config, rel_err
    out = str(tmp_path / "g.pt")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    m = models.Model(make_config()).cpu().eval(); m.is_cuda = False
    sd = m.state_dict(); sd.update({k: v for k, v in R.synthetic_params(seed=1).items() if k in sd}); m.load_state_dict(sd)
    x, y = R.synthetic_batch(4, 4000, seed=2)
    loss, _ = m(x, y); loss.backward()
    n = 0
    for k, p in m.named_parameters():
        if p.grad is None:
            assert k not in got
            continue
        assert got[k].dtype == p.grad.dtype
        assert rel_err(got[k], p.grad) < 1e-5, k
        n += 1
    assert n == 48
