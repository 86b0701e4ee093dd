"""Image-sharded multi-GPU plumbing: contiguous index shards, one NCCL all-gather of [N,512].

The reference is single-device; this is new.  Row order contract: rank r
owns records [r*ceil(N/R), min(N,(r+1)*ceil(N/R))), so concatenating the ranks' rows reproduces the
order extract_embeddings returns on one device (src/feature_extraction.py:295,305).  Works on any
torch.distributed backend (NCCL on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size) of the current process group, (0, 1) when not distributed."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def env_world() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def ensure_process_group(backend: str) -> bool:
    """Initialise torch.distributed from the torchrun environment if WORLD_SIZE > 1."""
    rank, size, _ = env_world()
    if size <= 1:
        return False
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, rank=rank, world_size=size)
    return True


def shard_bounds(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n items for `rank`."""
    per = (n + world_size - 1) // world_size
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def allgather_rows(local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather row blocks of possibly different heights: [n_r, D] per rank -> [sum n_r, D].

    Counts are exchanged first (R ints), blocks are padded to the largest count so that a single
    fixed-size all_gather moves the payload, then the padding is dropped.
    """
    rank, size = world()
    if size == 1:
        return local
    d = local.shape[1]
    counts = torch.zeros(size, dtype=torch.int64, device=local.device)
    mine = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(counts, mine, group=group)
    counts_host = [int(c) for c in counts.cpu()]
    cap = max(counts_host)
    if cap == 0:
        return local[:0]
    padded = local
    if local.shape[0] != cap:
        padded = torch.zeros((cap, d), dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
    gathered = torch.empty((size * cap, d), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, padded.contiguous(), group=group)
    if all(c == cap for c in counts_host):
        return gathered
    return torch.cat([gathered[r * cap : r * cap + c] for r, c in enumerate(counts_host)], dim=0)


def allgather_inplace(buf: torch.Tensor, cap: int, count: int, group=None) -> Tuple[torch.Tensor, List[int]]:
    """The all-gather of SURVEY.md 8e without staging copies.  `buf` is [world * cap, D] on every rank and this rank's
    `count` (<= cap) rows already sit at buf[rank * cap : rank * cap + count] -- the trunk's last kernel wrote them there.
    ONE collective moves the payload: on NCCL it runs in place (send buffer = receive buffer + rank * cap rows, which is
    NCCL's in-place all-gather).  Returns (matrix [sum(counts), D], counts): a view of `buf` when the shards are full
    (always, unless files failed to decode), else the compacted rows in rank order."""
    rank, size = world()
    if buf.dim() != 2 or buf.shape[0] != size * cap or not buf.is_contiguous() or not 0 <= count <= cap:
        raise ValueError("allgather_inplace: buf must be a contiguous [world * cap, D] tensor and 0 <= count <= cap")
    if size == 1:
        return buf[:count], [count]
    counts = torch.zeros(size, dtype=torch.int64, device=buf.device)
    dist.all_gather_into_tensor(counts, torch.tensor([count], dtype=torch.int64, device=buf.device), group=group)
    counts_host = [int(c) for c in counts.cpu()]
    if cap > 0 and max(counts_host) > 0:
        mine = buf[rank * cap : (rank + 1) * cap]
        if dist.get_backend(group) != "nccl":
            mine = mine.clone()  # only NCCL documents the aliased form
        dist.all_gather_into_tensor(buf, mine, group=group)
    total = sum(counts_host)
    last = max((r for r, c in enumerate(counts_host) if c), default=-1)
    if all(c == cap for c in counts_host[:last]):  # every shard before the last non-empty one is full: rows are contiguous
        return buf[:total], counts_host
    return torch.cat([buf[r * cap : r * cap + c] for r, c in enumerate(counts_host)], dim=0), counts_host


def allgather_objects(obj, group=None) -> List:
    """Small host-side metadata (kept record indices, failure paths, timings)."""
    _, size = world()
    if size == 1:
        return [obj]
    out: List = [None] * size
    dist.all_gather_object(out, obj, group=group)
    return out


def concat_in_rank_order(parts: Sequence[Sequence]) -> List:
    merged: List = []
    for p in parts:
        merged.extend(p)
    return merged


def bind_to_gpu_numa_node(gpu_index: int) -> bool:
    """Pin this process to the CPU cores NVML reports as local to its GPU (host decode threads, pinned staging
    buffers and H2D copies then stay on the GPU's socket).  Best effort; returns whether the affinity changed."""
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = ((os.cpu_count() or 64) + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return True
    except Exception:
        pass
    return False
