"""Shared MLP projector with an optional low-rank adapter -- drop-in for the reference's ``dmi/model/projector.py``.

Same constructor, attributes, state-dict keys (``net.0.{weight,bias}``, ``net.3.{weight,bias}``) and method names as
``Projector`` (projector.py:6-159); the arithmetic runs in the sm_100a kernels behind ``libdmi_b200.so``:
bf16 operands, fp32 accumulation, the adapter folded into the GEMMs as r extra K columns.

Compatibility switch ``lora_forward_mode``:
  * ``"as_written"`` (default, what the reference computes): ``lora_forward`` pairs ``zip(self.net, a, b, biases)``, so with
    n_layers=2 only ``Linear0 (+adapter 0)`` and the GELU are applied (projector.py:124).
  * ``"full"``: the complete 2-layer adapted MLP (what ``combine_lora`` / ``only_lora_forward`` compute).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from .. import ops
from .._lib import MLP_NO_ADAPTER, MLP_STOP_AFTER_FIRST_ACT
from ..utils.args import ProjectorArgs, setup_args


_KERNEL_RANKS = (8, 16, 32, 64)


def _kernel_rank(r: int) -> int:
    """the kernels are compiled for ranks 8/16/32/64; other ranks are zero-padded up (exact: padded factors contribute 0)"""
    for k in _KERNEL_RANKS:
        if r <= k:
            return k
    raise NotImplementedError(f"adapter rank {r} > 64 is not supported by the sm_100a kernels")


def _pad_cols(t: torch.Tensor, r: int) -> torch.Tensor:
    return t if t.shape[1] == r else torch.nn.functional.pad(t, (0, r - t.shape[1]))


def _pad_rows(t: torch.Tensor, r: int) -> torch.Tensor:
    return t if t.shape[0] == r else torch.nn.functional.pad(t, (0, 0, 0, r - t.shape[0]))


class _AdaptedMLPFn(torch.autograd.Function):
    """y = adapted_mlp(x; frozen base, adapter) with gradients to the adapter factors only (x is data, base is frozen)."""

    @staticmethod
    def forward(ctx, proj, flags, x, a0, b0, beta0, a1, b1, beta1):
        D, H = proj._in_dim(), proj.lm_emb_dim
        r_true = a0.numel() // D
        r = _kernel_rank(r_true)
        full = not (flags & MLP_STOP_AFTER_FIRST_ACT)
        B = x.shape[0]
        lin0, lin1 = proj.net[0], proj.net[-1]
        pa = lambda t, n_in: None if t is None else _pad_cols(t.detach().reshape(n_in, r_true), r)
        pb = lambda t: None if t is None else _pad_rows(t.detach().reshape(r_true, H), r)
        pk = proj._packed(r)
        repack = lambda: pk.pack_adapter(pa(a0, D), pb(b0), beta0, pa(a1, H) if full else None, pb(b1) if full else None,
                                         beta1 if full else None, lin0.bias, lin1.bias if full else None)
        repack()
        # the operand buffers are shared by every call on this projector: remember which adapter they hold, so that a backward
        # that runs after ANOTHER forward (two adapters in flight) re-packs its own adapter instead of using the wrong factors
        pk.adapter_epoch = getattr(pk, "adapter_epoch", 0) + 1
        ctx.adapter_epoch, ctx.repack = pk.adapter_epoch, repack
        ctx.r_true = r_true
        y = torch.empty(B, H, dtype=torch.float32, device=x.device)
        xd = x.detach()
        if xd.dtype == torch.bfloat16:
            # bf16 embeddings are consumed as they are (the GEMM operands are bf16 anyway -> identical results, half the bytes).
            # If x is the leading-D-column view of a [B, D+r] buffer it IS the operand buffer (zero copies); otherwise one device copy.
            from .._lib import MLP_X_PREPACKED
            base = xd._base if xd._base is not None else None
            inplace = (xd.stride(1) == 1 and xd.stride(0) == D + r and xd.storage_offset() == 0 and base is not None
                       and base.dtype == torch.bfloat16 and base.dim() == 2 and tuple(base.shape) == (B, D + r) and base.is_contiguous())
            if inplace:
                st = ops.MlpStash(B, D, H, r, x.device, full=full, xext=base)
            else:
                st = ops.MlpStash(B, D, H, r, x.device, full=full)
                st.xext[:, :D].copy_(xd)
            ops.adapted_mlp_fwd(pk, st, None, y, flags=flags | MLP_X_PREPACKED)
        else:
            st = ops.MlpStash(B, D, H, r, x.device, full=full)
            ops.adapted_mlp_fwd(pk, st, xd.float().contiguous(), y, flags=flags)
        ctx.proj, ctx.pk, ctx.st, ctx.flags, ctx.full = proj, pk, st, flags, full
        ctx.shapes = [None if t is None else t.shape for t in (a0, b0, beta0, a1, b1, beta1)]
        return y

    @staticmethod
    def backward(ctx, dy):
        pk, st, full = ctx.pk, ctx.st, ctx.full
        if getattr(pk, "adapter_epoch", None) != ctx.adapter_epoch:
            ctx.repack()
            pk.adapter_epoch = getattr(pk, "adapter_epoch", 0) + 1
            ctx.adapter_epoch = pk.adapter_epoch
        D, H, r = pk.D, pk.H, pk.r
        dev = dy.device
        def layer_grads(n_in, sfx):
            # [dA | dB | dbeta] of one layer live back to back in ONE zero-filled buffer -- the layout of the generated vector they
            # are slices of (hypernet.py:100-107), so the hypernet backward reads the buffer in place (no concatenation, one memset)
            flat = torch.zeros(n_in * r + r * H + H, dtype=torch.float32, device=dev)
            return {"dA" + sfx: flat[:n_in * r].view(n_in, r), "dB" + sfx: flat[n_in * r:n_in * r + r * H].view(r, H), "dbeta" + sfx: flat[n_in * r + r * H:]}
        grads = layer_grads(D, "0")
        if full:
            grads.update(layer_grads(H, "1"))
        if dy.dtype != torch.float32 or dy.stride(-1) != 1:      # a strided fp32 view (e.g. the prefix slot of inputs_embeds.grad) is fine
            dy = dy.float().contiguous()
        ops.adapted_mlp_bwd(pk, st, dy, grads, flags=ctx.flags)
        proj = ctx.proj
        if proj.accumulate_frozen_base_grads:
            # in the reference the frozen-by-omission projector has requires_grad=True, so autograd also fills
            # projector.net.0.{weight,bias}.grad (and net.3.* in the full form) on every micro-step, nothing ever zeroes them, and
            # clip_grad_norm_(HyperNetWrapper.parameters()) sees them (train_hypernet.py:148, hypernet.py:276-280).  Reproduced on request:
            # dW1 += dpre^T x, db1 += 1^T dpre (dW2 += dY^T h, db2 += 1^T dY) straight into .grad, like AccumulateGrad would.
            lin0, lin1 = proj.net[0], proj.net[-1]
            for q in (lin0.weight, lin0.bias) + ((lin1.weight, lin1.bias) if full else ()):
                if q.grad is None:
                    q.grad = torch.zeros_like(q)
            ops.gemm_mn(st.dpre, st.xext[:, :D], lin0.weight.grad, accumulate=True)
            lin0.bias.grad += st.dpre.float().sum(0)
            if full:
                ops.gemm_mn(st.dyext[:, :H], st.hext[:, :H], lin1.weight.grad, accumulate=True)
                lin1.bias.grad += dy.sum(0)
        order = ["dA0", "dB0", "dbeta0", "dA1", "dB1", "dbeta1"]
        rt = ctx.r_true
        outs = []
        for name, shp in zip(order, ctx.shapes):
            g = grads.get(name)
            if shp is None or g is None:
                outs.append(None)
                continue
            if rt != r:          # rank was zero-padded up to a kernel rank: drop the padding
                if name.startswith("dA"):
                    g = g[:, :rt].contiguous()
                elif name.startswith("dB"):
                    g = g[:rt].contiguous()
            outs.append(g.reshape(shp))
        return (None, None, None, *outs)


class Projector(nn.Module):
    def __init__(self, projector_args: ProjectorArgs, lm_emb_dim, mm_emb_dim, device):
        super().__init__()
        self.lm_emb_dim = lm_emb_dim
        self.mm_emb_dim = mm_emb_dim
        self.device = device
        setup_args(self, prefix="proj_", args=projector_args)
        self.lora_forward_mode = "as_written"
        # SURVEY H6 compatibility switch (default: the sane behaviour -- a frozen projector gets no gradients): see _AdaptedMLPFn.backward
        self.accumulate_frozen_base_grads = False
        # forward(): True / a callback(param) makes the backward accumulate dW, db straight into existing .grad tensors (mlp2.plain_mlp2)
        self.grad_in_place = None
        self._pk = {}
        self.build_model()

    # -- construction ------------------------------------------------------------------------------------------
    def act_function(self):
        if self.act != "quick_gelu":           # the only activation the reference knows (projector.py:17-22)
            raise NotImplementedError
        return nn.GELU                          # instantiated with approximate='tanh' 

    def build_model(self):
        layers: List[nn.Module] = []
        if self.arch == "linear":
            layers += [nn.Linear(self.mm_emb_dim, self.lm_emb_dim), nn.Dropout(self.dropout)]
        elif self.arch == "mlp":
            assert self.n_layers >= 2, f"MLP should at least have depth of two, cur depth = {self.n_layers}"
            width_in = self.mm_emb_dim
            for _ in range(self.n_layers - 1):
                layers += [nn.Linear(width_in, self.lm_emb_dim), self.act_function()(approximate="tanh"), nn.Dropout(self.dropout)]
                width_in = self.lm_emb_dim
            layers.append(nn.Linear(self.lm_emb_dim, self.lm_emb_dim))
        else:
            raise NotImplementedError
        self.net = nn.ModuleList(layers)
        self.to(self.device)

    def load_model(self):
        """reads ``torch.load(path)['projector_state_dict']``; column-prunes ``net.0.weight`` (projector.py:46-54)"""
        assert self.name_or_path is not None
        ckpt = torch.load(self.name_or_path, map_location=self.device)
        sd = ckpt["projector_state_dict"]
        if self.prune is not None:
            for k in list(sd):
                if "net.0.weight" in k:
                    sd[k] = sd[k][:, : self.prune]
        self.load_state_dict(sd)
        self._pk.clear()

    # -- operand cache -------------------------------------------------------------------------------------------
    def _in_dim(self) -> int:
        return self.net[0].weight.shape[1]

    def _is_mlp2(self) -> bool:
        return self.arch == "mlp" and self.n_layers == 2

    def _packed(self, r: int) -> "ops.PackedProjector":
        """bf16 operand cache for rank r, re-packed when the fp32 master weights change (optimizer step, load)."""
        w1, w2 = self.net[0].weight, self.net[-1].weight
        key = (r, w1.data_ptr(), w1._version, w2.data_ptr(), w2._version)
        hit = self._pk.get(r)
        if hit is None or hit[0] != key:
            pk = ops.PackedProjector(self._in_dim(), self.lm_emb_dim, r, w1.device)
            pk.pack_base(w1.detach(), w2.detach())
            self._pk[r] = (key, pk)
            return pk
        return hit[1]

    def _require_kernel_shape(self):
        if not self._is_mlp2():
            raise NotImplementedError("the sm_100a path implements the MLP2 projector (proj_arch='mlp', proj_n_layers=2)")

    # -- forward paths -------------------------------------------------------------------------------------------
    def forward(self, x):
        """plain MLP2 (projector.py:56-59)"""
        self._require_kernel_shape()
        from .mlp2 import plain_mlp2
        lin0, drop, lin1 = self.net[0], self.net[2], self.net[3]
        p = drop.p if (self.training and drop.p > 0) else 0.0
        return plain_mlp2(x, lin0.weight, lin0.bias, lin1.weight, lin1.bias, dropout_p=p, cache=self, grad_in_place=self.grad_in_place)

    def lora_forward(self, x, a_weights, b_weights, biases):
        """x:[B,D]; flat generated factors per Linear layer (projector.py:118-159)."""
        self._require_kernel_shape()
        n = len(a_weights)
        if biases is None:
            biases = [None] * n
        if self.lora_forward_mode == "as_written" or n < 2:
            if n < 2:
                raise NotImplementedError("lora_forward with a single weight set stops before the GELU in the reference; not built")
            flags = MLP_STOP_AFTER_FIRST_ACT
        else:
            flags = 0
        return _AdaptedMLPFn.apply(self, flags, x, a_weights[0], b_weights[0], biases[0], a_weights[1], b_weights[1], biases[1])

    def lora_forward_first_layer(self, x, a0, b0, beta0):
        """``lora_forward`` as written, given only what it reads (projector.py:124 / SURVEY H1): gelu(net[0](x) + (x A0) B0 + beta0).
        Used by ``HyperNetWrapper.forward`` so that the second generator need not run."""
        self._require_kernel_shape()
        return _AdaptedMLPFn.apply(self, MLP_STOP_AFTER_FIRST_ACT, x, a0, b0, beta0, None, None, None)

    def only_lora_forward(self, x, loras):
        """free-parameter LoRA on the frozen projector (projector.py:61-74, lora.py:15-17): scale alpha/r folded into B."""
        self._require_kernel_shape()
        l0, l1 = loras[0], loras[1]
        s0, s1 = l0.alpha / l0.rank, l1.alpha / l1.rank
        return _AdaptedMLPFn.apply(self, 0, x, l0.A, l0.B * s0, None, l1.A, l1.B * s1, None)

    def combine_lora(self, a_weights, b_weights, biases):
        """merge W' = (A B)^T + W, b' = beta + b per Linear into a new nn.Sequential (projector.py:76-116).
        The GELU / Dropout module *instances* are shared with this projector, as in the reference."""
        from .mlp2 import MergedLinear, merge_adapter
        if biases is None:
            biases = [torch.zeros(self.lm_emb_dim, device=self.device) for _ in range(len(a_weights))]
        modules, idx = [], 0
        for layer in self.net:
            if not isinstance(layer, nn.Linear):
                modules.append(layer)
                continue
            if idx >= len(a_weights):
                raise ValueError("Not enough weights provided for all linear layers")
            w, b = merge_adapter(layer.weight, layer.bias, a_weights[idx], b_weights[idx], biases[idx])
            modules.append(MergedLinear(w, b))
            idx += 1
        if idx < len(a_weights):
            raise ValueError("Too many weights provided")
        from .mlp2 import MergedMLP2
        return MergedMLP2(*modules).to(self.device)
