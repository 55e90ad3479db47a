"""Free-parameter LoRA baseline on the frozen projector -- drop-in for the reference's ``dmi/model/lora.py``.

``LoRALayer.{A [in,r], B [r,out]}``, ``LoraAdapters.loras``, ``LoraWrapper.forward/trainable_parameters/train`` keep the
reference names, init (A ~ N(0, 1/r), B = 0, lora.py:9-11) and state-dict keys; ``LoraWrapper.forward`` runs the same fused
adapted-MLP kernels as the hypernetwork path with A, B as the gradient targets."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..utils.args import LoraArgs, setup_args
from .projector import Projector


class LoRALayer(nn.Module):
    def __init__(self, in_dim, out_dim, rank, alpha):
        super().__init__()
        self.A = nn.Parameter(torch.randn(in_dim, rank) / torch.sqrt(torch.tensor(rank).float()))
        self.B = nn.Parameter(torch.zeros(rank, out_dim))
        self.rank = rank
        self.alpha = alpha

    def forward(self, x):
        """(alpha/r) * x A B (lora.py:15-17).  Stand-alone use only; inside the projector the term is fused into the GEMMs."""
        raise RuntimeError("LoRALayer is applied through Projector.only_lora_forward (fused sm_100a kernels); "
                           "there is no stand-alone / CPU path")


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
