"""Fused gradient clipping + AdamW on the sm_100a kernels (C ABI ``dmi_grad_sqnorm`` / ``dmi_grad_clip`` / ``dmi_adamw_step``).

Drop-in for the optimizer step of the reference trainers::

    torch.nn.utils.clip_grad_norm_(self.model.hypernet.parameters(), self.train_args.max_grad_norm)   # train_hypernet.py:148
    self.optimizer.step()                                                                             # train_hypernet.py:149

with ``optimizer = optim.AdamW(params=..., lr=..., weight_decay=..., betas=..., eps=...)`` (train_hypernet.py:526-532):
``FusedAdamW`` takes the same constructor arguments and keeps the same state (``step``, ``exp_avg``, ``exp_avg_sq``), so
``optimizer_state_dict`` checkpoints are interchangeable; ``clip_grad_norm_`` has torch's signature and in-place effect.
``FusedAdamW.step(max_grad_norm=...)`` fuses both: one read-only pass for the norm and one pass for the update
(28 B of HBM traffic per parameter instead of ~60), with the clip factor computed on the device (no host sync).

fp32 CUDA parameters only; there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional, Union

import torch

from . import _lib
from ._lib import OptTensor


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _descriptors(entries):
    """entries: list of (p|None, g, m|None, v|None) tensors -> ctypes array of dmi_opt_tensor"""
    arr = (OptTensor * max(len(entries), 1))()
    for i, (p, g, m, v) in enumerate(entries):
        for t in (p, g, m, v):
            if t is None:
                continue
            if not t.is_cuda:
                raise RuntimeError("dmi_b200.optim runs on CUDA tensors only (there is no CPU fallback)")
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise RuntimeError("dmi_b200.optim needs contiguous fp32 tensors")
        arr[i].p = None if p is None else p.data_ptr()
        arr[i].g = g.data_ptr()
        arr[i].m = None if m is None else m.data_ptr()
        arr[i].v = None if v is None else v.data_ptr()
        arr[i].n = g.numel()
    return arr


def grad_sqnorm(grads: List[torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """sum of squares of all gradient elements, as a 1-element fp32 device tensor (accumulated into ``out`` if given)"""
    if out is None:
        out = torch.zeros(1, dtype=torch.float32, device=grads[0].device)
    arr = _descriptors([(None, g, None, None) for g in grads])
    _lib.check(_lib.load().dmi_grad_sqnorm(arr, len(grads), C.c_void_p(out.data_ptr()), _stream()), "dmi_grad_sqnorm")
    return out


def clip_grad_norm_(parameters: Union[torch.Tensor, Iterable[torch.Tensor]], max_norm: float, norm_type: float = 2.0) -> torch.Tensor:
    """``torch.nn.utils.clip_grad_norm_`` (2-norm): scales every ``.grad`` in place by min(1, max_norm / (total_norm + 1e-6)) and
    returns the total norm as a 0-d device tensor.  Two kernels over the gradients, no host synchronisation."""
    if norm_type != 2.0:
        raise NotImplementedError("only the 2-norm used by the reference trainers is implemented")
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return torch.tensor(0.0)
    sq = grad_sqnorm(grads)
    arr = _descriptors([(None, g, None, None) for g in grads])
    _lib.check(_lib.load().dmi_grad_clip(arr, len(grads), float(max_norm), C.c_void_p(sq.data_ptr()), _stream()), "dmi_grad_clip")
    return sq.sqrt().reshape(())


class FusedAdamW(torch.optim.Optimizer):
    """``torch.optim.AdamW`` (amsgrad=False, maximize=False) with the whole update in one kernel per parameter group."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.last_grad_norm: Optional[torch.Tensor] = None

    def _state_of(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    @torch.no_grad()
    def step(self, closure=None, *, max_grad_norm: Optional[float] = None, write_clipped_grads: bool = False,
             sqnorm: Optional[torch.Tensor] = None):
        """One AdamW update.  ``max_grad_norm`` fuses clip_grad_norm_ over ALL parameter groups' gradients into the update
        (``write_clipped_grads`` also stores the scaled gradients, as clip_grad_norm_ would); ``sqnorm`` may supply the squared
        global norm (e.g. already all-reduced over parameter shards) instead of computing it here."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        groups = []
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if any(p.grad.is_sparse for p in ps):
                raise RuntimeError("FusedAdamW does not support sparse gradients")
            groups.append((group, ps))
        if max_grad_norm is not None and sqnorm is None:
            grads = [p.grad for _, ps in groups for p in ps]
            if grads:
                sqnorm = grad_sqnorm(grads)
        self.last_grad_norm = None if sqnorm is None else sqnorm.sqrt().reshape(())
        lib = _lib.load()
        for group, ps in groups:
            if not ps:
                continue
            entries, steps = [], set()
            for p in ps:
                st = self._state_of(p)
                st["step"] += 1
                steps.add(int(st["step"].item()) if st["step"].device.type == "cpu" else int(st["step"]))
                entries.append((p, p.grad, st["exp_avg"], st["exp_avg_sq"]))
            if len(steps) != 1:       # parameters that joined later: one call per distinct step count
                by_step = {}
                for e, p in zip(entries, ps):
                    by_step.setdefault(int(self.state[p]["step"]), []).append(e)
            else:
                by_step = {steps.pop(): entries}
            b1, b2 = group["betas"]
            for step, ents in by_step.items():
                arr = _descriptors(ents)
                rc = lib.dmi_adamw_step(arr, len(ents), float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                        float(group["weight_decay"]), step, float(max_grad_norm) if max_grad_norm is not None else -1.0,
                                        None if sqnorm is None else C.c_void_p(sqnorm.data_ptr()), int(write_clipped_grads), _stream())
                _lib.check(rc, "dmi_adamw_step")
        return loss
