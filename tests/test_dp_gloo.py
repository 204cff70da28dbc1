"""CPU, world_size 2 over gloo: the data-parallel reduction semantics of fit_step's host logic
(spff_b200/dp.py) — sum all-reduce + 1/world scale == DistributedDataParallel's mean of rank gradients,
tally reduction for logging, slice sharding for inference."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spff_b200 import dp
    torch.manual_seed(100 + rank)
    # each rank: gradient of ITS batch-mean loss for a tiny least-squares model
    w = torch.arange(6, dtype=torch.float32).reshape(2, 3).requires_grad_()
    x = torch.randn(5, 3)
    y = torch.randn(5, 2)
    loss = ((x @ w.t() - y) ** 2).mean()
    loss.backward()
    flat = w.grad.reshape(-1).clone()
    scale = dp.allreduce_grads(flat)
    nll = torch.tensor([float(rank + 1)], dtype=torch.float64)
    cnt = torch.tensor([10 * (rank + 1)], dtype=torch.int64)
    conf = torch.full((3, 3), rank + 1, dtype=torch.int64)
    dp.allreduce_tally(nll, cnt, conf)
    out[rank] = dict(local=w.grad.reshape(-1).clone(), reduced=flat * scale, scale=scale, nll=float(nll), cnt=int(cnt),
                     conf=int(conf.sum()), shard=dp.shard_range(7), world=dp.world())
    dist.destroy_process_group()


def test_dp_reduction_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    mean = (r0["local"] + r1["local"]) / 2
    assert torch.allclose(r0["reduced"], mean) and torch.allclose(r1["reduced"], mean)
    assert r0["scale"] == 0.5 and r0["nll"] == 3.0 and r1["cnt"] == 30 and r0["conf"] == 27
    assert r0["shard"] == (0, 4) and r1["shard"] == (4, 7) and r1["world"] == (1, 2)


def test_dp_single_process_is_identity():
    from spff_b200 import dp
    g = torch.ones(4)
    assert dp.allreduce_grads(g) == 1.0 and torch.equal(g, torch.ones(4)) and dp.world() == (0, 1)
