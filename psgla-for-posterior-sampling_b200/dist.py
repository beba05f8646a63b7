"""Multi-GPU plumbing: chains (2D) and (image, chain) pairs are independent, so the hot path shards with no per-step
communication.  One process per GPU; ``torch.distributed`` (NCCL over NVLink on GPUs, gloo in the CPU
tests) is used once per run, to gather the final samples / moments on rank 0 for the metric."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

__all__ = ["init_from_env", "world", "shard_range", "gather_to_rank0", "reduce_mean_to_rank0", "max_over_ranks"]


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment (no-op for a single process).
    Returns (rank, world_size, local_rank)."""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if ws > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=ws)
    return rank, ws, local


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_total: int, rank: int, world_size: int):
    """Contiguous block [start, stop) of global ids owned by ``rank``; sizes differ by at most one and the blocks tile
    [0, n_total) in rank order, so ``chain_id0 = start`` keeps every chain's Philox stream independent of world_size."""
    if world_size < 1 or not (0 <= rank < world_size) or n_total < 0:
        raise ValueError("bad shard request")
    q, r = divmod(n_total, world_size)
    start = rank * q + min(rank, r)
    return start, start + q + (1 if rank < r else 0)


def gather_to_rank0(local: torch.Tensor, n_total: int | None = None):
    """Concatenate per-rank blocks (first axis, possibly ragged by one) on rank 0; other ranks get None."""
    rank, ws = world()
    if ws == 1:
        return local
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(ws)]
    dist.all_gather(sizes, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device))
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    padded = local
    if local.shape[0] < mx:
        pad = torch.zeros((mx - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], 0)
    bufs = [torch.empty_like(padded) for _ in range(ws)]
    dist.all_gather(bufs, padded.contiguous())
    if rank != 0:
        return None
    out = torch.cat([b[:s] for b, s in zip(bufs, sizes)], 0)
    if n_total is not None and out.shape[0] != n_total:
        raise RuntimeError("gathered %d rows, expected %d" % (out.shape[0], n_total))
    return out


def reduce_mean_to_rank0(local_sum: torch.Tensor, local_count: int):
    """Pooled mean over all ranks' chains (sum of per-rank sums / total count) on rank 0; None elsewhere."""
    rank, ws = world()
    if ws == 1:
        return local_sum / max(local_count, 1)
    buf = torch.cat([local_sum.reshape(-1).double(), torch.tensor([float(local_count)], dtype=torch.float64,
                                                                   device=local_sum.device)])
    dist.reduce(buf, dst=0, op=dist.ReduceOp.SUM)
    if rank != 0:
        return None
    return (buf[:-1] / buf[-1]).reshape(local_sum.shape).to(local_sum.dtype)


def max_over_ranks(value: float, device=None) -> float:
    rank, ws = world()
    if ws == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
