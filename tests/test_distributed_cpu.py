"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: shard bookkeeping, the final ensemble gather
(equal and unequal shards) and the flat-buffer gradient all-reduce of the fine-tune exchange step."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from se3diff_b200.distributed import allreduce_gradients, gather_ensemble, sample_sharded, shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 256, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1 and sorted(sizes, reverse=True) == sizes


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _Out(dict):
    pass


def _worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = 5
        # equal shards
        frames = torch.full((3, L, 12), float(rank)) + torch.arange(3).view(3, 1, 1)
        g = gather_ensemble(frames)
        assert g.shape == (3 * world, L, 12)
        for r in range(world):
            assert torch.equal(g[3 * r:3 * r + 3], torch.full((3, L, 12), float(r)) + torch.arange(3).view(3, 1, 1))
        # unequal shards through sample_sharded: 5 samples over 2 ranks -> 3 + 2, seeded with the global offset
        def denoiser(*, batch, tag):
            n = batch
            pos = torch.randn(n * L, 3)                      # consumes the generator seeded with (seed + start)
            rot = torch.eye(3).repeat(n * L, 1, 1) * tag
            return _Out(pos=pos, node_orientations=rot)

        ens = sample_sharded(denoiser, make_batch=lambda n: n, num_samples=5, seed=100, tag=2.0)
        assert ens.shape == (5, L, 12)
        torch.manual_seed(100)
        first = torch.randn(3 * L, 3).view(3, L, 3)
        torch.manual_seed(103)
        second = torch.randn(2 * L, 3).view(2, L, 3)
        assert torch.equal(ens[:3, :, :3], first) and torch.equal(ens[3:, :, :3], second)
        assert torch.equal(ens[..., 3:], (torch.eye(3) * 2.0).reshape(9).expand(5, L, 9))
        # gradient all-reduce: mean over ranks, parameters without grad are skipped
        p1, p2, p3 = (torch.nn.Parameter(torch.zeros(4, 3)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2)))
        p1.grad = torch.full((4, 3), float(rank + 1))
        p2.grad = torch.arange(7.0) * (rank + 1)
        allreduce_gradients([p1, p2, p3])
        mean = sum(range(1, world + 1)) / world
        assert torch.allclose(p1.grad, torch.full((4, 3), mean)) and torch.allclose(p2.grad, torch.arange(7.0) * mean) and p3.grad is None
        results[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_gather_and_allreduce_world2_gloo():
    world, port = 2, _free_port()
    mgr = mp.get_context("spawn").Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    assert dict(results) == {0: "ok", 1: "ok"}
