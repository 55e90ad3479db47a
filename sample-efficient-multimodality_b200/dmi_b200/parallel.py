"""Data parallelism for the adapted-projector path: one process per GPU, replicated parameters, bucketed gradient
all-reduce over NCCL (NVLink 5 / NVSwitch) overlapped with the remaining backward.

The reference is single-process; its closest analogue is gradient accumulation (train_hypernet.py:119-149,
train_projector.py:51-73: ``loss / GA`` summed over GA micro-steps).  ``world`` ranks each running ``GA_local`` micro-steps and
all-reducing with ``op=SUM`` and scale ``1/(world*GA_local)`` give the same update as the reference with
``GA = world*GA_local``; ``tests/test_parallel_cpu.py`` checks that equivalence with gloo on the CPU.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist


class FlatGrads:
    """Named gradient tensors carved out of ONE flat fp32 buffer, grouped in buckets that are contiguous ranges, so that a
    step needs one memset and one all-reduce per bucket.  Buckets are listed in the order their gradients become available
    in the backward pass (layer 1 first)."""

    def __init__(self, shapes: Dict[str, Sequence[int]], buckets: List[List[str]], device, dtype=torch.float32):
        names = [n for b in buckets for n in b]
        assert sorted(names) == sorted(shapes), "every gradient must belong to exactly one bucket"
        total = 0
        self._ranges = {}
        self._bucket_ranges = []
        for b in buckets:
            start = total
            for n in b:
                numel = 1
                for s in shapes[n]:
                    numel *= int(s)
                numel_pad = (numel + 3) // 4 * 4          # keep every view 16-byte aligned
                self._ranges[n] = (total, numel, tuple(int(s) for s in shapes[n]))
                total += numel_pad
            self._bucket_ranges.append((start, total))
        self.flat = torch.zeros(total, dtype=dtype, device=device)
        self.views = {n: self.flat[o:o + k].view(shp) for n, (o, k, shp) in self._ranges.items()}
        self.buckets = [self.flat[a:b] for a, b in self._bucket_ranges]

    def zero_(self):
        self.flat.zero_()

    def __getitem__(self, name: str) -> torch.Tensor:
        return self.views[name]


class BucketAllReducer:
    """All-reduces the buckets of a FlatGrads on a side stream.  ``reduce_bucket(i, ready_event)`` may be called while the
    compute stream is still producing later buckets; ``wait()`` makes the compute stream wait for all reductions."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None, average: bool = True):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.average = average
        self._cuda = torch.cuda.is_available()
        self.stream = torch.cuda.Stream() if self._cuda else None
        self._pending = []

    def reduce_bucket(self, bucket: torch.Tensor, ready_event: Optional["torch.cuda.Event"] = None) -> None:
        if self.world == 1:
            return
        if bucket.is_cuda:
            if ready_event is not None:
                self.stream.wait_event(ready_event)
            else:
                self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                work = dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                if self.average:
                    work.wait()                      # stream-level wait on the NCCL stream, not a host block
                    bucket.mul_(1.0 / self.world)
                    work = None
            bucket.record_stream(self.stream)
        else:
            work = dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            work.wait()
            if self.average:
                bucket.mul_(1.0 / self.world)
            work = None
        if work is not None:
            self._pending.append(work)

    def wait(self) -> None:
        for w in self._pending:
            w.wait()
        self._pending.clear()
        if self._cuda and self.world > 1:
            torch.cuda.current_stream().wait_stream(self.stream)


def allreduce_module_grads(params, group=None, average: bool = True) -> None:
    """simple (non-overlapped) gradient all-reduce for a list of parameters, coalesced into one flat buffer"""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat.mul_(1.0 / dist.get_world_size(group))
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def global_grad_norm_clip_(params, max_norm: float) -> torch.Tensor:
    """clip_grad_norm_ semantics (train_hypernet.py:148) -- identical on every rank once gradients are all-reduced"""
    return torch.nn.utils.clip_grad_norm_(list(params), max_norm)


def shard_support_sets(items, rank: Optional[int] = None, world: Optional[int] = None, group=None):
    """Few-shot adapter generation shards by support set: rank r takes items[r::world].  Every rank then
    accumulates ``(1/N_total) * e_n`` over ITS support sets and ``allreduce_sum_`` of those partial means gives the global mean
    modality code on every rank (``HyperNetwork.mean_adapter(zs_local, n_total=N_total)`` does both)."""
    if rank is None or world is None:
        if dist.is_initialized():
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        else:
            rank, world = 0, 1
    return list(items[rank::world])


def allreduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    """in-place SUM all-reduce; a no-op in a single-process run"""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t
