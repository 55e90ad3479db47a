"""Plain / merged MLP2 projector on the sm_100a kernels, with gradients to W1,b1,W2,b2.

Used by ``Projector.forward`` (reference projector.py:56-59, the ``train_projector.py`` path, dropout active in training)
and by the merged ``generated_projector`` that ``combine_lora`` returns for few-shot fine-tuning
(projector.py:76-116, train_hypernet.py:220-251)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .._lib import MLP_BASE_GRADS, MLP_DROPOUT, MLP_NO_ADAPTER


class _PlainMLP2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, keep, dropout_p, cache, grad_in_place):
        H, D = w1.shape
        dev = x.device
        pk = None
        key = (w1.data_ptr(), w1._version, w2.data_ptr(), w2._version)
        if cache is not None:
            hit = getattr(cache, "_plain_pk", None)
            if hit is not None and hit[0] == key:
                pk = hit[1]
        if pk is None:
            pk = ops.PackedProjector(D, H, 0, dev)
            pk.pack_base(w1.detach(), w2.detach())
            if cache is not None:
                cache._plain_pk = (key, pk)
        pk.pack_adapter(None, None, None, None, None, None, b1, b2)           # bias0 = b1, bias1 = b2
        B = x.shape[0]
        st = ops.MlpStash(B, D, H, 0, dev, full=True)
        y = torch.empty(B, H, dtype=torch.float32, device=dev)
        flags = MLP_NO_ADAPTER | (MLP_DROPOUT if keep is not None else 0)
        ops.adapted_mlp_fwd(pk, st, x.detach().float().contiguous(), y, flags=flags, keep=keep, dropout_p=dropout_p)
        ctx.pk, ctx.st, ctx.flags, ctx.keep, ctx.p = pk, st, flags, keep, dropout_p
        ctx.params, ctx.grad_in_place = (w1, b1, w2, b2), grad_in_place
        return y

    @staticmethod
    def backward(ctx, dy):
        pk, st = ctx.pk, ctx.st
        D, H = pk.D, pk.H
        dev = dy.device
        w1, b1, w2, b2 = ctx.params
        cb = ctx.grad_in_place
        in_place = bool(cb) and all(q.requires_grad and q.grad is not None and q.grad.is_contiguous() and q.grad.dtype == torch.float32 for q in ctx.params)
        if in_place:
            # the C entry accumulates (dW += ...): write straight into the existing .grad tensors -- no zero-filled temporaries and no
            # AccumulateGrad add kernels (43 of the 163 us of a B = 1024 step, profiles/r2_plain_launches.txt)
            grads = dict(dW1=w1.grad, db1=b1.grad, dW2=w2.grad, db2=b2.grad)
        else:
            z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
            grads = dict(dW1=z(H, D), db1=z(H), dW2=z(H, H), db2=z(H))
        ops.adapted_mlp_bwd(pk, st, dy.float().contiguous(), grads, flags=ctx.flags | MLP_BASE_GRADS, keep=ctx.keep, dropout_p=ctx.p)
        if in_place:
            if callable(cb):                   # e.g. parallel.GradSync.notify: these parameters' gradients are final for this backward
                for q in (w2, b2, w1, b1):
                    cb(q)
            return (None,) * 9
        return None, grads["dW1"], grads["db1"], grads["dW2"], grads["db2"], None, None, None, None


def plain_mlp2(x, w1, b1, w2, b2, *, dropout_p: float = 0.0, keep: torch.Tensor = None, cache=None, grad_in_place=None):
    """y = gelu_tanh(x W1^T + b1) [* keep/(1-p)] W2^T + b2.  ``keep`` (uint8/bool [B,H]) may be injected for parity tests;
    otherwise it is drawn with torch's generator on the device (RNG streams are never bit-matched).
    ``grad_in_place`` (True or a callback taking the parameter): when every parameter already has a contiguous fp32 ``.grad`` the
    backward accumulates into it directly and returns no gradients to autograd (tensor hooks on the parameters do not fire; the
    callback is how a gradient synchroniser learns that they are ready)."""
    if dropout_p > 0.0 and keep is None:
        # one kernel (rand >= p followed by a cast was three)
        keep = torch.empty(x.shape[0], w1.shape[0], dtype=torch.uint8, device=x.device).bernoulli_(1.0 - dropout_p)
    elif keep is not None:
        keep = keep.to(torch.uint8).contiguous()
    return _PlainMLP2Fn.apply(x, w1, b1, w2, b2, keep, float(dropout_p), cache, grad_in_place)


def merge_adapter(weight, bias, a_flat, b_flat, beta):
    return ops.merge_adapter(weight, bias, a_flat, b_flat, beta)


class MergedLinear(nn.Linear):
    """nn.Linear whose parameters are the merged (W + (AB)^T, b + beta); state-dict keys ``weight`` / ``bias``"""

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor):
        nn.Module.__init__(self)
        self.out_features, self.in_features = weight.shape
        self.weight = nn.Parameter(weight)
        self.bias = nn.Parameter(bias)


class MergedMLP2(nn.Sequential):
    """``nn.Sequential(Linear, GELU, Dropout, Linear)`` as returned by ``combine_lora`` (same child indices 0..3, hence the
    same ``generated_projector.{0,3}.*`` checkpoint keys), whose forward runs the fused kernels."""

    def forward(self, x):
        lin0, drop, lin1 = self[0], self[2], self[3]
        p = drop.p if (drop.training and drop.p > 0) else 0.0
        return plain_mlp2(x, lin0.weight, lin0.bias, lin1.weight, lin1.bias, dropout_p=p, cache=self, grad_in_place=getattr(self, "grad_in_place", None))
