"""world_size-2 gloo tests of the N>1 host logic: batch sharding + final all_gather for
sampling / encoding, and the data-parallel gradient all-reduce of a training step.  CPU only."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from infodiffusion_b200.distributed import gather_batch, local_slice, sharded_map, sharded_sampling
from infodiffusion_b200.layout import shard_range


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_sampler(n, xT=None, a=None):
    # per-sample independent stand-in for DiffusionProcess.sampling: every output row depends on its own row only
    assert xT.shape[0] == n and a.shape[0] == n
    return torch.tanh(xT) * 0.5 + a.sum(dim=1).view(-1, 1, 1, 1)


def _fake_encoder(x):
    flat = x.flatten(1)
    return flat[:, :4] * 2.0, flat[:, 4:6] - 1.0


def _worker(rank: int, world: int, port: int, total: int, tmp: str):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(64)                 # every rank draws the same full batch (reference seed)
        xT = torch.randn(total, 3, 8, 8, generator=g)
        a = torch.randn(total, 5, generator=g)
        want = _fake_sampler(total, xT=xT, a=a)
        got = sharded_sampling(_fake_sampler, total, xT=xT, a=a)
        assert got.shape == want.shape and torch.equal(got, want)
        lo, hi = shard_range(total, rank, world)
        assert torch.equal(local_slice(xT, total), xT[lo:hi])
        part = sharded_sampling(_fake_sampler, total, xT=xT, a=a, gather=False)
        assert torch.equal(part, want[lo:hi])
        z, w = sharded_map(_fake_encoder, xT)
        zr, wr = _fake_encoder(xT)
        assert torch.equal(z, zr) and torch.equal(w, wr)
        assert torch.equal(gather_batch(xT[lo:hi].contiguous(), total), xT)

        # data-parallel gradient exchange: average over ranks, parameters without a gradient are skipped
        from infodiffusion_b200.train import allreduce_gradients
        ps = [torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(2))]
        ps[0].grad = torch.full((3, 2), float(rank + 1))
        ps[2].grad = torch.arange(2.0) * (rank + 1)
        allreduce_gradients(ps, world)
        mean = sum(range(1, world + 1)) / world
        assert torch.allclose(ps[0].grad, torch.full((3, 2), mean))
        assert ps[1].grad is None
        assert torch.allclose(ps[2].grad, torch.arange(2.0) * mean)
        # overlapped variant: the conv-stack backward hands its flat buffer to GradSync.reduce, finish() covers the rest
        from infodiffusion_b200.train import GradSync
        sync = GradSync(world)
        qs = [torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(2, 2)), torch.nn.Parameter(torch.zeros(3))]
        flat = torch.arange(8.0) * (rank + 1)
        qs[0].grad, qs[1].grad = flat[:4], flat[4:].view(2, 2)
        qs[2].grad = torch.full((3,), 10.0 * (rank + 1))
        sync.reduce(flat, qs[:2])
        sync.finish(qs)
        assert torch.allclose(qs[0].grad, torch.arange(4.0) * mean)
        assert torch.allclose(qs[1].grad, (torch.arange(4.0) + 4).view(2, 2) * mean)
        assert torch.allclose(qs[2].grad, torch.full((3,), 10.0 * mean))
        assert not sync.pending and not sync.covered
        # a gradient that does NOT alias the reduced buffer (autograd cloned it, or it existed before backward: gradient
        # accumulation, zero_grad(set_to_none=False)) must still end up holding the REDUCED values
        rs = [torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(2, 2))]
        flat2 = torch.arange(8.0) * (rank + 1)
        views = [flat2[:4], flat2[4:].view(2, 2)]
        rs[0].grad = views[0]                              # stolen view: aliases flat2
        rs[1].grad = views[1].clone()                      # local copy: would stay un-reduced without the check
        sync.reduce(flat2, rs, views)
        del views
        sync.finish(rs)
        assert torch.allclose(rs[0].grad, torch.arange(4.0) * mean)
        assert torch.allclose(rs[1].grad, (torch.arange(4.0) + 4).view(2, 2) * mean)
        assert sync.fixed_up == 1
        with open(os.path.join(tmp, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [5, 8, 1])
def test_sharding_and_gather_world2(tmp_path, total):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_single_process_is_identity():
    x = torch.randn(3, 4)
    assert gather_batch(x, 3) is x
    assert torch.equal(local_slice(x, 3), x)
