"""GPU, 2 ranks over NCCL (skipped on a single-GPU box): the data-parallel fit_step — per-rank shards,
decoder-range all-reduce overlapped with the encoder backward, 1/world in the Adam kernel — gives every
rank the gradients and the parameters of one process stepping on the concatenated batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from innovative3D import config as C
    from oracle import spff_oracle as O
    torch.manual_seed(42)
    lit = dict((v[0], v[1]) for v in C.VARIANTS)["SPFF-UNet"]().cuda()
    x, lab = O.phantom_batch(4, 32, 32, seed=77)          # no ignored voxels: equal N_valid per rank
    lo, hi = rank * 2, rank * 2 + 2
    lit.hparams["lr"] = 1e-3
    o = lit.fit_step((x[lo:hi].cuda(), lab[lo:hi].cuda()), sample_group=1)
    torch.cuda.synchronize()
    out[rank] = dict(grad=lit._fused["grad"].cpu(), flat=lit.model._flat.detach().cpu(), loss=float(o["loss"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_step_equals_single_process_step():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    from innovative3D import config as C
    from oracle import spff_oracle as O
    torch.manual_seed(42)
    lit = dict((v[0], v[1]) for v in C.VARIANTS)["SPFF-UNet"]().cuda()
    x, lab = O.phantom_batch(4, 32, 32, seed=77)
    lit.hparams["lr"] = 1e-3
    lit.fit_step((x.cuda(), lab.cuda()), sample_group=1)
    g1 = lit._fused["grad"].cpu()
    rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-30))
    # summed rank gradients / world == full-batch gradient (both are means over the same valid voxels)
    assert torch.equal(out[0]["grad"], out[1]["grad"])
    assert rel(out[0]["grad"] / world, g1) < 2e-3
    assert torch.equal(out[0]["flat"], out[1]["flat"])
    # Adam's first step moves every weight by ~lr * sign(g): near-zero gradients may flip sign between the two
    # summation orders, so parameters agree to within 2 lr element-wise (and closely in the mean)
    diff = (out[0]["flat"] - lit.model._flat.detach().cpu()).abs()
    assert float(diff.max()) <= 2.1e-3 and float((diff > 1e-4).float().mean()) < 0.05


def _worker_3dunet(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from innovative3D import config as C
    from oracle import cicek_oracle as CO
    from oracle import spff_oracle as O
    lit = dict((v[0], v[1]) for v in C.VARIANTS)["3DUNet"]().cuda()
    lit.load_state_dict(CO.det_weights(seed=42), strict=True)
    lit.train()
    x, lab = O.phantom_batch(4, 32, 32, seed=78)
    lo, hi = rank * 2, rank * 2 + 2
    o = lit.fit_step((x[lo:hi].cuda(), lab[lo:hi].cuda()))
    torch.cuda.synchronize()
    out[rank] = dict(grad=lit._fused["grad"].cpu(), flat=lit.backbone._flat.detach().cpu(), loss=float(o["loss"]),
                     rm=lit.state_dict()["backbone.enc1.1.running_mean"].cpu())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_3dunet_step():
    """3DUNet under data parallelism: per-rank BatchNorm statistics (no SyncBN), ONE all-reduce, the same
    SGD step on every rank — the reduced gradient is the sum of the two shards' single-process gradients, and the
    parameters equal a single-process SGD step with their mean."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_3dunet, args=(world, _free_port(), out), nprocs=world, join=True)
    from innovative3D import config as C
    from oracle import cicek_oracle as CO
    from oracle import spff_oracle as O
    from spff_b200 import ops
    x, lab = O.phantom_batch(4, 32, 32, seed=78)
    weights = CO.det_weights(seed=42)
    shard = []
    for r in range(world):
        lit = dict((v[0], v[1]) for v in C.VARIANTS)["3DUNet"]().cuda()
        lit.load_state_dict(weights, strict=True)
        lit.train()
        lit.fit_step((x[2 * r:2 * r + 2].cuda(), lab[2 * r:2 * r + 2].cuda()), optimize=False)
        shard.append(lit._fused["grad"].clone())
        assert torch.equal(lit.state_dict()["backbone.enc1.1.running_mean"].cpu(), out[r]["rm"])   # per-rank statistics
    assert torch.equal(out[0]["grad"], out[1]["grad"]) and torch.equal(out[0]["flat"], out[1]["flat"])
    assert torch.equal(out[0]["grad"], (shard[0] + shard[1]).cpu())
    flat = lit.backbone._flat.detach().clone()     # still the initial weights (optimize=False)
    buf = torch.zeros_like(flat)
    ops.sgd_step(flat, shard[0] + shard[1], buf, 1e-2, 0.99, 0.0, False, True, 0.5)
    assert torch.equal(flat.cpu(), out[0]["flat"])
    assert not torch.equal(out[0]["rm"], out[1]["rm"])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_model_on_a_device_that_is_not_the_current_one():
    """One process, two GPUs: a model moved with plain `.to('cuda:1')` while cuda:0 stays the current device runs on
    cuda:1's stream with cuda:1 as the launch device (the binding follows its tensors), and both devices can be used in
    turn (the native per-device caches: shared-memory opt-ins, SM counts). Tensors of two devices in one call raise."""
    from innovative3D import config as C
    from oracle import spff_oracle as O
    from spff_b200 import ops
    torch.cuda.set_device(0)
    x, lab = O.phantom_batch(2, 32, 32, seed=5, ignore_frac=0.01)
    outs = []
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        torch.manual_seed(42)
        lit = dict((v[0], v[1]) for v in C.VARIANTS)["SPFF-UNet"]().to(dev)
        assert torch.cuda.current_device() == 0
        with torch.no_grad():
            logits = lit(x.to(dev))
        o = lit.fit_step((x.to(dev), lab.to(dev)), optimize=False)
        torch.cuda.synchronize(dev)
        assert logits.device == torch.device(dev) and torch.cuda.current_device() == 0
        outs.append((logits.cpu(), float(o["loss"]), lit._fused["grad"].cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][0], outs[2][0])
    assert outs[0][1] == outs[1][1]
    assert float((outs[0][2] - outs[1][2]).norm() / outs[0][2].norm()) < 1e-4     # atomics' ordering only
    a = torch.zeros(1, 5, 16, 16, 32, dtype=torch.bfloat16, device="cuda:0")
    b = torch.zeros(1, 5, 16, 16, 32, dtype=torch.bfloat16, device="cuda:1")
    w = torch.zeros(32, 32, 3, 3, 3, device="cuda:0")
    wf, _ = ops.pack_conv3_weight(w)
    with pytest.raises(RuntimeError, match="different devices"):
        ops.conv3d_k3_fwd(a, 32, wf, b, 32)
    ops.conv3d_k3_fwd(a, 32, wf, a.clone(), 32)      # the failed call left no stale device record behind
