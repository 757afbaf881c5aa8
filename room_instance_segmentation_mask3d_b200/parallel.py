"""Multi-GPU plumbing for the hot path: scene sharding and the training gradient all-reduce.

The path shards by scene: hash keys carry the batch index, kernel maps never cross scenes and
InstanceNorm is per scene, so forward needs NO data-path collective — every rank runs its own scenes and results are
gathered on the host.  Training adds exactly one exchange step per optimizer step, the gradient all-reduce (the reference
itself configures no distributed strategy, main_instance_segmentation.py:76-83; its only collective is the 1-float
all_reduce(num_masks) of criterion.py:244-246, mirrored by `allreduce_scalar`).

One process per GPU, `torch.distributed` (backend "nccl" on the GPUs, "gloo" in the CPU tests).  Nothing here touches
feature data on the host; buckets are flat device tensors.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard_scenes(voxel_counts: Sequence[int], world_size: int) -> List[List[int]]:
    """Size-balanced partition of scenes over ranks: greedy longest-processing-time (largest scene first onto the
    least-loaded rank, ties -> lowest rank).  Deterministic, so every rank computes the same partition without
    communication.  Returns, per rank, the scene indices it owns in ascending order (a scene is never split: a halo
    exchange per conv level would be needed)."""
    assert world_size >= 1
    loads = [0] * world_size
    owned: List[List[int]] = [[] for _ in range(world_size)]
    for i in sorted(range(len(voxel_counts)), key=lambda j: (-int(voxel_counts[j]), j)):
        r = min(range(world_size), key=lambda q: (loads[q], q))
        owned[r].append(i)
        loads[r] += int(voxel_counts[i])
    return [sorted(o) for o in owned]


def allreduce_gradients(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20, average: bool = True,
                        group: Optional[dist.ProcessGroup] = None) -> int:
    """Sum (or average) the gradients of `params` over all ranks through flat buckets.  Returns the number of collectives.

    Parameters without a gradient on this rank (e.g. `backbone.final.*`, which the model never calls, res16unet.py:231-238)
    contribute zeros, so the bucket layout is identical on every rank and ranks whose scenes produced no gradient for a
    parameter still take part.  NVSwitch reduces in the switch (NVLS), so buckets are sized for launch latency
    (64 MB default: 3 collectives for the 39.6 M-parameter model), not per link."""
    params = [p for p in params if p.requires_grad]
    if not params:
        return 0
    ws = dist.get_world_size(group) if dist.is_initialized() else 1
    n_coll = 0
    bucket: List[torch.nn.Parameter] = []
    size = 0

    def flush():
        nonlocal bucket, size, n_coll
        if not bucket:
            return
        dev, dt = bucket[0].device, torch.float32
        flat = torch.zeros(sum(p.numel() for p in bucket), dtype=dt, device=dev)
        off = 0
        for p in bucket:
            if p.grad is not None:
                flat[off:off + p.numel()].copy_(p.grad.reshape(-1))
            off += p.numel()
        if ws > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            n_coll += 1
            if average:
                flat.div_(ws)
        off = 0
        for p in bucket:
            g = flat[off:off + p.numel()].view_as(p)
            if p.grad is None:
                p.grad = g.to(p.dtype).clone()
            else:
                p.grad.copy_(g)
            off += p.numel()
        bucket, size = [], 0

    for p in params:
        bucket.append(p)
        size += p.numel() * 4
        if size >= bucket_bytes:
            flush()
    flush()
    return n_coll


def allreduce_scalar(value: float, device, average: bool = False, group: Optional[dist.ProcessGroup] = None) -> float:
    """The reference's num_masks normalisation (criterion.py:244-246): sum over ranks (then / world when `average`)."""
    t = torch.tensor([float(value)], dtype=torch.float32, device=device)
    if world() > 1:
        dist.all_reduce(t, group=group)
        if average:
            t /= world()
    return float(t.item())


def gather_scene_results(local: list, owned: Sequence[int], n_scenes: int, group: Optional[dist.ProcessGroup] = None) -> Optional[list]:
    """Inference: collect per-scene host results from all ranks on rank 0 in scene order (no device collective)."""
    if world() == 1:
        out = [None] * n_scenes
        for i, r in zip(owned, local):
            out[i] = r
        return out
    pairs = [None] * world()
    dist.all_gather_object(pairs, list(zip(owned, local)), group=group)
    if rank() != 0:
        return None
    out = [None] * n_scenes
    for pr in pairs:
        for i, r in pr:
            out[i] = r
    return out
