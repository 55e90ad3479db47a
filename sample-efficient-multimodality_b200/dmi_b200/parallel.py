"""Data parallelism for the adapted-projector path: one process per GPU, replicated parameters, bucketed gradient
all-reduce over NCCL (NVLink 5 / NVSwitch) overlapped with the remaining backward and with the next step.

The reference is single-process; its closest analogue is gradient accumulation (train_hypernet.py:119-149,
train_projector.py:51-73: ``loss / GA`` summed over GA micro-steps).  ``world`` ranks each running ``GA_local`` micro-steps and
all-reducing with ``op=SUM`` and scale ``1/(world*GA_local)`` give the same update as the reference with
``GA = world*GA_local``; ``tests/test_parallel_cpu.py`` checks that equivalence with gloo on the CPU.

Four pieces:
  * ``FlatGrads`` + ``BucketAllReducer``: named gradient tensors carved out of one flat buffer, all-reduced bucket by bucket on a
    side stream in backward-availability order; every reduction leaves a CUDA event, so the consumer (optimizer step, or the
    next re-use of a double-buffered gradient buffer) waits at the point of USE instead of at the end of the backward.
  * ``SymmAllReducer``: for the small per-step adapter gradients -- a one-shot all-reduce kernel of this library over NVSwitch peer
    memory (``multimem.ld_reduce`` where NVLS is available), enqueued in the step's own stream.
  * ``GradSync``: the same for an ``nn.Module`` -- ``p.grad`` of every parameter becomes a view into a flat bucket buffer
    (no ``torch.cat`` / copy-back), buckets are released by post-accumulate-grad hooks as autograd finishes them.
  * ``Rank1FactorSync``: the generator-0 gradient of the hypernetwork is rank-1 per micro-step (``dG = dw (x) e``, SURVEY
    appendix A), so ranks exchange the factors (``[92160 + 768]`` floats per micro-step instead of 283 MB) with an all-gather and
    apply the rank-(world*GA) update locally: the same sum up to fp32 reassociation, ~750x less traffic.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def _world(group=None) -> int:
    return dist.get_world_size(group) if dist.is_initialized() else 1


class FlatGrads:
    """Named gradient tensors carved out of ONE flat fp32 buffer, grouped in buckets that are contiguous ranges, so that a
    step needs one memset and one all-reduce per bucket.  Buckets are listed in the order their gradients become available
    in the backward pass (layer 1 first)."""

    def __init__(self, shapes: Dict[str, Sequence[int]], buckets: List[List[str]], device, dtype=torch.float32,
                 storage: Optional[torch.Tensor] = None):
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
        if storage is not None:        # caller-owned flat buffer (e.g. peer-mapped symmetric memory for SymmAllReducer)
            assert storage.dtype == dtype and storage.dim() == 1 and storage.numel() >= total and storage.is_contiguous()
            self.flat = storage[:total]
            self.flat.zero_()
        else:
            self.flat = torch.zeros(total, dtype=dtype, device=device)
        self.views = {n: self.flat[o:o + k].view(shp) for n, (o, k, shp) in self._ranges.items()}
        self.buckets = [self.flat[a:b] for a, b in self._bucket_ranges]

    def zero_(self):
        self.flat.zero_()

    def __getitem__(self, name: str) -> torch.Tensor:
        return self.views[name]


class BucketAllReducer:
    """All-reduces buckets on a side stream.  ``reduce_bucket(bucket, ready_event)`` may be called while the compute stream is still
    producing later buckets.  Completion is tracked on the side stream: ``done_event()`` returns a CUDA event that fires when
    everything issued so far has been reduced (wait for it where the gradients are consumed -- e.g. one step later when the
    gradient buffers are double-buffered), ``wait()`` makes the compute stream wait for all of it right away."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None, average: bool = True):
        self.group = group
        self.world = _world(group)
        self.average = average
        self._cuda = torch.cuda.is_available()
        self.stream = torch.cuda.Stream() if self._cuda else None

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
                work.wait()                          # stream-level: the side stream waits for the NCCL stream, the host does not block
                if self.average:
                    bucket.mul_(1.0 / self.world)
            bucket.record_stream(self.stream)
        else:
            dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=self.group)
            if self.average:
                bucket.mul_(1.0 / self.world)

    def done_event(self) -> Optional["torch.cuda.Event"]:
        if not self._cuda or self.world == 1:
            return None
        ev = torch.cuda.Event()
        ev.record(self.stream)
        return ev

    def wait(self) -> None:
        if self._cuda and self.world > 1:
            torch.cuda.current_stream().wait_stream(self.stream)


class SymmAllReducer:
    """One-shot all-reduce of a small flat fp32 buffer through NVLink / NVSwitch peer memory (``dmi_allreduce_oneshot``), enqueued in
    the CALLER's stream: no side stream and no NCCL kernel competing with the step's persistent GEMMs for SMs.

    ``n_slots`` input buffers of ``numel`` floats live in torch symmetric memory (peer-mapped on every rank of the group, with an NVLS
    multicast mapping where the platform supports it -- the kernel then reduces inside the switch with ``multimem.ld_reduce``); the
    gradients of a step are accumulated straight into ``self.inputs[k]`` (hand it to ``FlatGrads(storage=...)``) and
    ``reduce(k)`` leaves ``scale * sum over ranks`` in ``self.outputs[k]`` (plain local memory).  Falls back to nothing: construction
    raises if symmetric memory cannot be set up, and the caller decides (bench.py then uses the NCCL reducer and says so)."""

    def __init__(self, numel: int, device, n_slots: int = 2, group=None, use_multicast: bool = True):
        import ctypes as C

        import torch.distributed._symmetric_memory as symm_mem

        from . import _lib
        self._C, self._lib = C, _lib
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        assert self.world <= 8, "SymmAllReducer: one NVSwitch domain (<= 8 GPUs)"
        self.numel = (numel + 3) // 4 * 4
        self.n_slots = n_slots
        lib = _lib.load()
        nflag = int(lib.dmi_allreduce_flag_words())
        self._buf = symm_mem.empty(n_slots * self.numel, dtype=torch.float32, device=device)
        self._buf.zero_()
        self._flags = symm_mem.empty(n_slots * nflag, dtype=torch.int32, device=device)
        self._flags.zero_()
        torch.cuda.synchronize(device)
        self._hb = symm_mem.rendezvous(self._buf, self.group)
        self._hf = symm_mem.rendezvous(self._flags, self.group)
        dist.barrier(self.group)               # every rank has zeroed its flags before anybody's kernel can signal
        mc = int(getattr(self._hb, "multicast_ptr", 0) or 0) if use_multicast else 0
        self.multicast = mc != 0
        self.inputs = [self._buf[k * self.numel:(k + 1) * self.numel] for k in range(n_slots)]
        self.outputs = [torch.zeros(self.numel, dtype=torch.float32, device=device) for _ in range(n_slots)]
        ptr_arr = C.c_void_p * self.world
        self._bufs, self._flgs, self._mc = [], [], []
        for k in range(n_slots):
            self._bufs.append(ptr_arr(*[int(b) + k * self.numel * 4 for b in self._hb.buffer_ptrs]))
            self._flgs.append(ptr_arr(*[int(b) + k * nflag * 4 for b in self._hf.buffer_ptrs]))
            self._mc.append(C.c_void_p(mc + k * self.numel * 4) if mc else None)
        self._epoch = [0] * n_slots
        self._side = None

    def reduce(self, k: int, scale: float = 1.0) -> torch.Tensor:
        """outputs[k] = scale * sum over ranks of inputs[k]; enqueued on the current stream; every rank must call it in the same order"""
        self._epoch[k] += 1
        C = self._C
        rc = self._lib.load().dmi_allreduce_oneshot(self._bufs[k], self._flgs[k], self._mc[k], self.rank, self.world,
                                                    C.c_void_p(self.outputs[k].data_ptr()), self.numel, float(scale), self._epoch[k],
                                                    C.c_void_p(torch.cuda.current_stream().cuda_stream))
        self._lib.check(rc, "dmi_allreduce_oneshot")
        return self.outputs[k]

    def reduce_async(self, k: int, scale: float = 1.0) -> "torch.cuda.Event":
        """Same reduction on this object's side stream, ordered after everything enqueued on the current stream so far.  The kernel's
        CTAs are small enough to co-reside with the step's persistent GEMM CTAs, so the next step proceeds underneath it -- including
        the time this rank spends at the kernel's barrier waiting for the slowest rank.  Returns the event to wait for before
        ``inputs[k]`` is written again / ``outputs[k]`` is read."""
        if self._side is None:
            self._side = torch.cuda.Stream()
        self._side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._side):
            self.reduce(k, scale)
            ev = torch.cuda.Event()
            ev.record(self._side)
        return ev


class GradSync:
    """Bucketed, overlapped gradient all-reduce for the parameters of a module (hypernet / projector / merged projector training).

    ``p.grad`` of every parameter is made a view into one flat buffer per bucket (so an all-reduce needs no gather / scatter
    copies; autograd and the library's fused accumulation both add in place).  Buckets are filled in the order given -- pass the
    parameters in BACKWARD-AVAILABILITY order (for the hypernet: generator weights first, then q/k/v and the prefix tokens) -- and
    each bucket is all-reduced on the side stream as soon as autograd has accumulated its last gradient
    (``register_post_accumulate_grad_hook``), overlapping the rest of the backward.  ``finish()`` is called where the gradients are
    consumed (before clipping / the optimizer step).

    Gradient accumulation: call ``backward()`` ``GA_local`` times, with ``sync.enabled = False`` for all but the last micro-step
    (DDP's ``no_sync``), then ``finish()``.  With ``scale = 1/(world*GA_local)`` folded into the loss this equals the reference's
    single-process accumulation over ``GA = world*GA_local`` micro-steps (train_hypernet.py:119-149)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20, group=None, average: bool = False):
        self.params = [p for p in params if p.requires_grad]
        self.reducer = BucketAllReducer(group, average=average)
        self.enabled = True
        self.buckets: List[torch.Tensor] = []
        self._bucket_of: Dict[int, int] = {}
        self._pending: List[int] = []
        cur: List[torch.nn.Parameter] = []
        cur_bytes = 0
        groups: List[List[torch.nn.Parameter]] = []
        for p in self.params:
            nb = p.numel() * 4
            if cur and cur_bytes + nb > bucket_bytes:
                groups.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nb
        if cur:
            groups.append(cur)
        for bi, grp in enumerate(groups):
            total = sum((p.numel() + 3) // 4 * 4 for p in grp)
            flat = torch.zeros(total, dtype=torch.float32, device=grp[0].device)
            off = 0
            for p in grp:
                assert p.dtype == torch.float32, "GradSync: fp32 master parameters only"
                p.grad = flat[off:off + p.numel()].view_as(p)
                self._bucket_of[id(p)] = bi
                off += (p.numel() + 3) // 4 * 4
            self.buckets.append(flat)
            self._pending.append(len(grp))
        self._sizes = [len(g) for g in groups]
        self._done = [False] * len(groups)         # reduction already issued this cycle
        self._touched = [False] * len(groups)      # at least one gradient landed in the bucket this cycle
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]

    def zero_grad(self) -> None:
        """zero the flat buffers in place (the views stay attached: never set ``p.grad = None`` under GradSync)"""
        for b in self.buckets:
            b.zero_()
        self._pending = list(self._sizes)
        self._done = [False] * len(self.buckets)
        self._touched = [False] * len(self.buckets)

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        bi = self._bucket_of[id(p)]
        self._touched[bi] = True
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._pending[bi] = self._sizes[bi]                      # re-armed for the next backward pass (gradient accumulation)
            if self.enabled and not self._done[bi]:
                self.reducer.reduce_bucket(self.buckets[bi], None)  # ready = everything enqueued on the compute stream so far
                self._done[bi] = True

    def notify(self, p: torch.nn.Parameter) -> None:
        """for gradients that a kernel accumulated into ``p.grad`` in place, bypassing autograd's AccumulateGrad (and therefore its
        hooks): ``HyperNetwork.grad_ready_callback = sync.notify`` with ``fuse_generator_grad_accumulation``"""
        if id(p) in self._bucket_of:
            self._on_grad(p)

    def finish(self) -> None:
        """all-reduce the buckets no hook released (a parameter of the bucket received no gradient in the last pass, or the bucket
        filled while ``enabled`` was off) -- buckets nothing was accumulated into are all-zero on every rank and are skipped -- and
        make the compute stream wait for every reduction.  Call where the gradients are consumed (clip / optimizer step)."""
        for bi in range(len(self.buckets)):
            if self._touched[bi] and not self._done[bi]:
                self.reducer.reduce_bucket(self.buckets[bi], None)
                self._done[bi] = True
        self.reducer.wait()

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()


class Rank1FactorSync:
    """Factor exchange for a gradient that is a sum of outer products: ``dG = sum_k dw_k (x) e_k`` (the hypernetwork's generator
    gradients, one term per micro-step).  Each rank appends its ``(dw, e)`` pairs; ``apply_(G_grad)``
    all-gathers the factors of all ranks and accumulates ``DW^T E`` into ``G_grad`` (and ``sum_k dw_k`` into the bias gradient):
    the same sum as a dense all-reduce up to fp32 reassociation, with ``(out + D)`` floats per micro-step on the wire instead of
    ``out * D``."""

    def __init__(self, out_features: int, in_features: int, device, max_terms: int = 64, group=None):
        self.out, self.inp, self.group = out_features, in_features, group
        self.world = _world(group)
        self.max_terms = max_terms
        self.dw = torch.zeros(max_terms, out_features, dtype=torch.float32, device=device)
        self.e = torch.zeros(max_terms, in_features, dtype=torch.float32, device=device)
        self.n = 0
        self._gather = None

    def push(self, dw: torch.Tensor, e: torch.Tensor) -> None:
        assert self.n < self.max_terms, "Rank1FactorSync: more micro-steps than max_terms"
        self.dw[self.n].copy_(dw.reshape(-1))
        self.e[self.n].copy_(e.reshape(-1))
        self.n += 1

    def apply_(self, G_grad: torch.Tensor, bias_grad: Optional[torch.Tensor] = None) -> None:
        """G_grad [out, in] += sum over ranks and local terms of dw (x) e;  bias_grad [out] += sum of dw.
        Every rank must have pushed the same number of terms (one per micro-step of the optimizer step); nothing here
        synchronises the host."""
        n = self.n
        dw, e = self.dw[:n], self.e[:n]
        if self.world > 1 and n > 0:
            if self._gather is None or self._gather[0].shape[0] != self.world * n:
                self._gather = (torch.empty(self.world * n, self.out, dtype=torch.float32, device=dw.device),
                                torch.empty(self.world * n, self.inp, dtype=torch.float32, device=dw.device))
            dw_all, e_all = self._gather
            dist.all_gather_into_tensor(dw_all, dw.contiguous(), group=self.group)
            dist.all_gather_into_tensor(e_all, e.contiguous(), group=self.group)
            dw, e = dw_all, e_all
        if dw.shape[0] > 0:
            G_grad.addmm_(dw.t(), e)               # fp32 rank-(world*GA) update (cuBLAS fp32; exact-sum semantics, no tf32)
            if bias_grad is not None:
                bias_grad.add_(dw.sum(0))
        self.n = 0


def allreduce_module_grads(params, group=None, average: bool = True, bucket_bytes: int = 64 << 20) -> None:
    """Non-overlapped gradient all-reduce for a list of parameters whose ``.grad`` tensors were produced by autograd (no flat
    views): gradients are reduced in buckets of at most ``bucket_bytes`` -- a single gradient larger than that is all-reduced in
    place without any copy, smaller ones are coalesced through one staging buffer per bucket.  (``GradSync`` is the overlapped,
    copy-free form.)"""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    world = dist.get_world_size(group)
    grads = [p.grad for p in params if p.grad is not None]
    small: List[torch.Tensor] = []
    small_bytes = 0

    def flush():
        nonlocal small, small_bytes
        if not small:
            return
        if len(small) == 1:
            g = small[0]
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
            if average:
                g.mul_(1.0 / world)
        else:
            flat = torch.cat([g.reshape(-1) for g in small])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            if average:
                flat.mul_(1.0 / world)
            off = 0
            for g in small:
                n = g.numel()
                g.copy_(flat[off:off + n].view_as(g))
                off += n
        small, small_bytes = [], 0

    for g in grads:
        nb = g.numel() * g.element_size()
        if nb >= bucket_bytes and g.is_contiguous():
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
            if average:
                g.mul_(1.0 / world)
            continue
        if small_bytes + nb > bucket_bytes:
            flush()
        small.append(g)
        small_bytes += nb
    flush()


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
