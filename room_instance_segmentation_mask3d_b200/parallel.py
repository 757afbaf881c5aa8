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


class GradientBuckets:
    """Bucketed gradient all-reduce overlapped with the backward pass (the one collective of the training path: data-parallel
    replicas over scenes; the reference itself only has the 1-float all_reduce(num_masks),
    mask3d/models/criterion.py:244-246).

    * Every parameter's `.grad` is a VIEW into a persistent flat fp32 bucket, so autograd accumulates straight into the buffer the
      collective reads and the optimizer reads the reduced values in place: no per-step flat-buffer allocation, no copy-in, no
      copy-out.  `zero()` clears the buckets (one memset each) instead of `zero_grad(set_to_none=True)`.
    * Buckets follow REVERSE registration order (backward produces the decoder's gradients first, the stem's last); a
      post-accumulate hook per parameter counts its bucket down and launches `all_reduce(bucket, async_op=True)` as soon as the
      bucket and all earlier ones are complete (same collective order on every rank), so the transfer runs on the
      communication stream under the rest of backward.  NVSwitch reduces in the
      switch (NVLS), so buckets are sized for launch latency and overlap (32 MB default: ~5 collectives for the 39.6 M
      parameters), not per link.
    * Parameters whose gradient this rank did not produce in a step (other ranks may have) still take part: their bucket slice
      stays zero, and `finish()` launches every bucket a hook has not launched.  Parameters that NO rank ever uses are left out
      by the caller (`backbone.final.*`, constructed and never called, res16unet.py:231-238): their `.grad` stays None, as under
      the reference optimizer, so AdamW never touches them.
    * world size 1: same code path without the collective (hooks only count)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 32 << 20, average: bool = True,
                 group: Optional[dist.ProcessGroup] = None):
        self.params = [p for p in params if p.requires_grad]
        self.group, self.average = group, average
        self.ws = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.buckets: List[torch.Tensor] = []
        self._members: List[List[torch.nn.Parameter]] = []
        self._bucket_of = {}
        cur: List[torch.nn.Parameter] = []
        size = 0
        for p in reversed(self.params):
            cur.append(p)
            size += p.numel() * 4
            if size >= bucket_bytes:
                self._close(cur)
                cur, size = [], 0
        if cur:
            self._close(cur)
        self._pending = [len(m) for m in self._members]
        self._handles: List = []
        self._launched = [False] * len(self.buckets)
        self._next = 0
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self.collectives = 0

    def _close(self, members: List[torch.nn.Parameter]) -> None:
        dev = members[0].device
        flat = torch.zeros(sum(p.numel() for p in members), dtype=torch.float32, device=dev)
        off = 0
        bi = len(self.buckets)
        for p in members:
            assert p.dtype == torch.float32, "gradient buckets are fp32 (the model's parameters are fp32, as in the reference)"
            p.grad = flat[off:off + p.numel()].view_as(p)
            self._bucket_of[id(p)] = bi
            off += p.numel()
        self.buckets.append(flat)
        self._members.append(list(members))

    def _launch(self, bi: int) -> None:
        self._launched[bi] = True
        if self.ws > 1:
            op = dist.ReduceOp.AVG if (self.average and dist.get_backend(self.group) == "nccl") else dist.ReduceOp.SUM
            self._handles.append((bi, dist.all_reduce(self.buckets[bi], op=op, group=self.group, async_op=True), op))
            self.collectives += 1

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        bi = self._bucket_of[id(p)]
        flat = self.buckets[bi]
        if p.grad is None or p.grad.data_ptr() < flat.data_ptr() or p.grad.data_ptr() >= flat.data_ptr() + flat.numel() * 4:
            # autograd replaced the view (first accumulation after a set_to_none): copy it back into the bucket and re-attach
            off = sum(q.numel() for q in self._members[bi][:self._members[bi].index(p)])
            view = flat[off:off + p.numel()].view_as(p)
            if p.grad is not None:
                view.copy_(p.grad)
            p.grad = view
        self._pending[bi] -= 1
        # collectives must be issued in the same order on every rank: buckets launch strictly in index order, each as soon as
        # it and all earlier ones are complete
        while self._next < len(self.buckets) and self._pending[self._next] == 0:
            self._launch(self._next)
            self._next += 1

    def zero(self) -> None:
        """Start of a step: clear the buckets (the gradients are views into them) and re-arm the hooks."""
        for b in self.buckets:
            b.zero_()
        self._pending = [len(m) for m in self._members]
        self._launched = [False] * len(self.buckets)
        self._handles = []
        self._next = 0

    def finish(self) -> int:
        """After backward: launch the buckets no hook completed (parameters without a gradient on this rank), wait for all
        collectives (the compute stream then waits for the communication stream).  Returns the number of collectives of the step."""
        while self._next < len(self.buckets):
            self._launch(self._next)
            self._next += 1
        n = len(self._handles)
        for bi, h, op in self._handles:
            h.wait()
            if self.average and op == dist.ReduceOp.SUM:
                self.buckets[bi].div_(self.ws)
        self._handles = []
        return n

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []


def allreduce_gradients(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20, average: bool = True,
                        group: Optional[dist.ProcessGroup] = None) -> int:
    """One-shot form (after backward, not overlapped): sum (or average) the gradients of `params` over all ranks through flat
    buckets; returns the number of collectives.  A parameter whose gradient is None on EVERY rank keeps `grad = None` (a 1-byte
    presence map is reduced first), so the optimizer skips it exactly as the reference's does; one that some rank produced is
    zero-filled on the others.  world size 1: nothing to do.  For training loops use `GradientBuckets` (persistent buffers,
    overlapped with backward)."""
    params = [p for p in params if p.requires_grad]
    ws = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    if not params or ws == 1:
        return 0
    dev = params[0].device
    present = torch.tensor([0 if p.grad is None else 1 for p in params], dtype=torch.int32, device=dev)
    dist.all_reduce(present, op=dist.ReduceOp.MAX, group=group)
    present = present.tolist()
    params = [p for p, f in zip(params, present) if f]
    n_coll = 1
    bucket: List[torch.nn.Parameter] = []
    size = 0

    def flush():
        nonlocal bucket, size, n_coll
        if not bucket:
            return
        flat = torch.zeros(sum(p.numel() for p in bucket), dtype=torch.float32, device=dev)
        off = 0
        for p in bucket:
            if p.grad is not None:
                flat[off:off + p.numel()].copy_(p.grad.reshape(-1))
            off += p.numel()
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
