"""Multi-GPU plumbing: one process per GPU, independent samples sharded across ranks, one collective.

The reference has no distributed code; sampling is embarrassingly parallel over
samples (attention never crosses samples, SDE scalars are per graph), so ranks never exchange data
on the path.  The only exchange is the final ensemble gather of `[B_local, L, 12]` fp32 frames
(3.1 MB/GPU at L = 512, B = 128: latency-bound, plain NCCL all_gather over NVSwitch).
Seeding follows sample.py:288-306: every sub-batch is seeded with its global sample offset.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(expected_world: int | None = None, backend: str | None = None):
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun) and joins the process group when
    WORLD_SIZE > 1.  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    if expected_world is not None and expected_world != world and rank == 0 and world == 1 and expected_world > 1:
        raise RuntimeError(f"--gpus {expected_world} needs a torchrun launch with {expected_world} ranks (WORLD_SIZE={world})")
    return rank, world, local_rank


def shard_range(num_samples: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [start, end) of the global sample index space owned by `rank`; the first
    `num_samples % world` ranks take one extra sample."""
    base, rem = divmod(num_samples, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_ensemble(frames: torch.Tensor, counts: list[int] | None = None) -> torch.Tensor:
    """All ranks contribute `[B_local, L, 12]` frames (pos | row-major rotation); every rank receives the
    `[sum B_local, L, 12]` ensemble in rank order.  Unequal shard sizes are padded to the maximum."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return frames
    world = dist.get_world_size()
    if counts is None:
        out = torch.empty((world * frames.shape[0],) + tuple(frames.shape[1:]), dtype=frames.dtype, device=frames.device)
        dist.all_gather_into_tensor(out, frames.contiguous())      # concatenated form: accepted by NCCL and gloo
        return out
    bmax = max(counts)
    pad = frames.new_zeros((bmax,) + tuple(frames.shape[1:]))
    pad[: frames.shape[0]] = frames
    out = torch.empty((world * bmax,) + tuple(frames.shape[1:]), dtype=frames.dtype, device=frames.device)
    dist.all_gather_into_tensor(out, pad)
    return torch.cat([out[r * bmax: r * bmax + counts[r]] for r in range(world)], dim=0)


def allreduce_gradients(params, average: bool = True) -> None:
    """Fine-tune exchange step (hook for finetune.py:625): ONE flat-buffer all_reduce(SUM) of all
    gradients (193,806 floats = 0.78 MB for the bioemu-v1.0 fine-tune model), divided by world size."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if average:
        flat /= dist.get_world_size()
    o = 0
    for g in grads:
        g.copy_(flat[o:o + g.numel()].view_as(g))
        o += g.numel()


def sample_sharded(denoiser, *, make_batch, num_samples: int, seed: int = 0, **denoiser_kwargs) -> torch.Tensor:
    """Runs `denoiser` on this rank's shard and returns the gathered ensemble `[num_samples, L, 12]`.
    `make_batch(n)` builds a batch of n graphs of ONE sequence (equal lengths)."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    start, end = shard_range(num_samples, rank, world)
    counts = [shard_range(num_samples, r, world)[1] - shard_range(num_samples, r, world)[0] for r in range(world)]
    torch.manual_seed(seed + start)
    out = denoiser(batch=make_batch(end - start), **denoiser_kwargs)
    n = end - start
    frames = torch.cat([out["pos"].view(n, -1, 3), out["node_orientations"].reshape(n, -1, 9)], dim=-1)
    return gather_ensemble(frames, counts if len(set(counts)) > 1 else None)
