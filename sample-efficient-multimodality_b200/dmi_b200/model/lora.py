"""Free-parameter LoRA baseline on the frozen projector -- drop-in for the reference's ``dmi/model/lora.py``.

``LoRALayer.{A [in,r], B [r,out]}``, ``LoraAdapters.loras``, ``LoraWrapper.forward/trainable_parameters/train`` keep the
reference names, init (A ~ N(0, 1/r), B = 0, lora.py:9-11) and state-dict keys; ``LoraWrapper.forward`` runs the same fused
adapted-MLP kernels as the hypernetwork path with A, B as the gradient targets."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..utils.args import LoraArgs, setup_args
from .projector import Projector


class _LoRAFn(torch.autograd.Function):
    """y = s * (x A) B with u = x A kept for the backward: dB = s u^T dy, du = s dy B^T, dA = x^T du (x is data: no gradient)."""

    @staticmethod
    def forward(ctx, x, A, B, s):
        from .. import ops
        ops._need_cuda(x, A, B)
        bf = torch.bfloat16
        M, r, N = x.shape[0], A.shape[1], B.shape[1]
        if r not in (8, 16, 32, 64) or A.shape[0] % 8 or N % 8:
            raise NotImplementedError("LoRALayer.forward: rank must be 8/16/32/64 and the widths multiples of 8")
        xb = x.detach().to(bf).contiguous()
        u = torch.empty(M, r, dtype=bf, device=x.device)
        ops.skinny_rows(xb, A.detach().t().to(bf).contiguous(), u)                       # u = x A
        y = torch.empty(M, N, dtype=torch.float32, device=x.device)
        ops.gemm_tn(u, B.detach().t().to(bf).contiguous(), out0=y, alpha=s)              # y = s u B
        ctx.save_for_backward(xb, u, B)
        ctx.s = s
        return y

    @staticmethod
    def backward(ctx, dy):
        from .. import ops
        xb, u, B = ctx.saved_tensors
        bf = torch.bfloat16
        M, r = u.shape
        dyb = dy.detach().to(bf).contiguous()
        dB = torch.zeros(r, dyb.shape[1], dtype=torch.float32, device=dy.device)
        ops.outer_reduce(u, dyb, dB, scale=ctx.s)                                        # dB = s u^T dy
        du = torch.empty(M, r, dtype=bf, device=dy.device)
        ops.skinny_rows(dyb, (B.detach() * ctx.s).to(bf).contiguous(), du)               # du = s dy B^T
        dA = torch.zeros(xb.shape[1], r, dtype=torch.float32, device=dy.device)
        ops.outer_reduce(du, xb, dA, transpose_out=True)                                 # dA = x^T du
        return None, dA, dB, None


class LoRALayer(nn.Module):
    def __init__(self, in_dim, out_dim, rank, alpha):
        super().__init__()
        self.A = nn.Parameter(torch.randn(in_dim, rank) / torch.sqrt(torch.tensor(rank).float()))
        self.B = nn.Parameter(torch.zeros(rank, out_dim))
        self.rank = rank
        self.alpha = alpha

    def forward(self, x):
        """(alpha/r) * x A B (lora.py:15-17) on the sm_100a kernels (bf16 operands, fp32 accumulation), gradients to A and B.
        Stand-alone use only: inside the projector the term is fused into the GEMMs as r extra K columns."""
        return _LoRAFn.apply(x, self.A, self.B, float(self.alpha) / float(self.rank))


class LoraAdapters(nn.Module):
    def __init__(self, lora_args: LoraArgs, lm_emb_dim, mm_emb_dim, device):
        super().__init__()
        self.lm_emb_dim = lm_emb_dim
        self.mm_emb_dim = mm_emb_dim
        self.device = device
        setup_args(self, prefix="lora_", args=lora_args)
        self.build_model()

    def build_model(self):
        self.loras = nn.ModuleList(
            LoRALayer(self.mm_emb_dim if i == 0 else self.lm_emb_dim, self.lm_emb_dim, self.rank, self.alpha)
            for i in range(self.n_proj_layers))
        self.to(self.device)

    def forward(self, x):
        pass


class LoraWrapper(nn.Module):
    def __init__(self, lora_args, proj_args, lm_emb_dim, mm_emb_dim, device):
        super().__init__()
        self.device = device
        self.lora_adapters = LoraAdapters(lora_args, lm_emb_dim, mm_emb_dim, device)
        self.projector = Projector(proj_args, lm_emb_dim, mm_emb_dim, device)
        self.projector.load_model()

    def train(self, mode=True):
        if not isinstance(mode, bool):
            raise ValueError("training mode is expected to be boolean")
        self.training = mode
        for child in self.children():
            child.train(mode)
        self.projector.eval()          # the frozen projector never leaves eval (lora.py:47-55)
        return self

    def forward(self, x):
        return self.projector.only_lora_forward(x, self.lora_adapters.loras)

    def trainable_parameters(self):
        return self.lora_adapters.parameters()
