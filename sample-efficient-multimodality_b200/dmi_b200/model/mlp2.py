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
    def forward(ctx, x, w1, b1, w2, b2, keep, dropout_p, cache):
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
        return y

    @staticmethod
    def backward(ctx, dy):
        pk, st = ctx.pk, ctx.st
        D, H = pk.D, pk.H
        dev = dy.device
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        grads = dict(dW1=z(H, D), db1=z(H), dW2=z(H, H), db2=z(H))
        ops.adapted_mlp_bwd(pk, st, dy.float().contiguous(), grads, flags=ctx.flags | MLP_BASE_GRADS, keep=ctx.keep, dropout_p=ctx.p)
        return None, grads["dW1"], grads["db1"], grads["dW2"], grads["db2"], None, None, None


def plain_mlp2(x, w1, b1, w2, b2, *, dropout_p: float = 0.0, keep: torch.Tensor = None, cache=None):
    """y = gelu_tanh(x W1^T + b1) [* keep/(1-p)] W2^T + b2.  ``keep`` (uint8/bool [B,H]) may be injected for parity tests;
    otherwise it is drawn with torch's generator on the device (RNG streams are never bit-matched)."""
    if dropout_p > 0.0 and keep is None:
        keep = (torch.rand(x.shape[0], w1.shape[0], device=x.device) >= dropout_p)
    if keep is not None:
        keep = keep.to(torch.uint8).contiguous()
    return _PlainMLP2Fn.apply(x, w1, b1, w2, b2, keep, float(dropout_p), cache)


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
        return plain_mlp2(x, lin0.weight, lin0.bias, lin1.weight, lin1.bias, dropout_p=p, cache=self)
