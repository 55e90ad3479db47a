"""Multimodal wrappers -- drop-in for ``dmi/model/mmmodel.py``: project the modality embedding, splice it as ONE prefix
token in front of the text embeddings, call the frozen LLM.

The projection and the splice (mmmodel.py:36-48, :118-135, :205-221) run in the sm_100a kernels; the LLM call itself is the
stock Hugging Face module, exactly as in the reference (out of scope of the hot path).  ``embeds_dtype`` selects the dtype of
``inputs_embeds``: ``torch.float32`` reproduces the reference (``torch.cat`` promotes fp32 projected + bf16 table to fp32),
``torch.bfloat16`` is the bandwidth-saving mode.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.nn as nn

from .. import _lib, ops


class _SpliceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, projected, table, input_ids, labels, attention_masks, out_dtype):
        ops._need_cuda(projected, table, input_ids, labels, attention_masks)
        B, H = projected.shape
        T = input_ids.shape[1]
        dev = projected.device
        ids = input_ids.to(torch.int64).contiguous()
        out = torch.empty(B, 1 + T, H, dtype=out_dtype, device=dev)
        lab = labels.to(torch.int64).contiguous() if labels is not None else None
        lab_out = torch.empty(B, 1 + T, dtype=torch.int64, device=dev) if lab is not None else None
        mask = None
        mask_out = None
        if attention_masks is not None:
            mask = attention_masks.contiguous()
            if mask.dtype not in (torch.int64, torch.float32):
                mask = mask.float()
            mask_out = torch.empty(B, 1 + T, dtype=torch.float32, device=dev)
        proj = projected.detach()
        if proj.dtype not in (torch.float32, torch.bfloat16) or proj.stride(1) != 1:
            proj = proj.float().contiguous()
        tab = table.detach()
        assert tab.dtype in (torch.float32, torch.bfloat16) and tab.stride(1) == 1
        errf = ops.splice_error_flag(dev)
        errf.poll()                         # an out-of-range id seen by an EARLIER launch raises here (no synchronisation)
        err = errf.flag
        rc = _lib.load().dmi_splice(
            ops._ptr(proj) if proj.dtype == torch.float32 else None, ops._ptr(proj) if proj.dtype == torch.bfloat16 else None, proj.stride(0),
            ops._ptr(tab), int(tab.dtype == torch.bfloat16), tab.stride(0), tab.shape[0], ops._ptr(ids), B, T, H,
            ops._ptr(out), int(out_dtype == torch.bfloat16), ops._ptr(lab), ops._ptr(lab_out),
            ops._ptr(mask), int(mask is not None and mask.dtype == torch.int64), ops._ptr(mask_out), ops._ptr(err), ops._stream())
        _lib.check(rc, "dmi_splice")
        errf.arm()
        ctx.proj_dtype = projected.dtype
        ctx.mark_non_differentiable(*[t for t in (lab_out, mask_out) if t is not None])
        return out, lab_out, mask_out

    @staticmethod
    def backward(ctx, g_out, g_lab, g_mask):
        # only the prefix slot carries gradient back to the projector; the embedding table is frozen
        g = g_out[:, 0, :]
        if g.dtype != ctx.proj_dtype:
            g = g.to(ctx.proj_dtype)
        return g, None, None, None, None, None


def splice_prefix(projected, embed_table, input_ids, attention_masks=None, labels=None, embeds_dtype=torch.float32):
    """inputs_embeds [B,1+T,H], attention_masks [B,1+T] (float, leading 1), labels [B,1+T] (leading -100)"""
    out, lab, mask = _SpliceFn.apply(projected, embed_table, input_ids, labels, attention_masks, embeds_dtype)
    return out, mask, lab


class _MMBase(nn.Module):
    def __init__(self, llm, device, mm_emb_dim, name, pad_token_id):
        super().__init__()
        self.llm = llm
        self.device = device
        self.name = name
        self.pad_token_id = pad_token_id
        self.llm_dim = self.llm.config.hidden_size
        self.mm_emb_dim = mm_emb_dim
        self.embeddings = self.llm.get_input_embeddings()
        self.embeds_dtype = torch.float32            # reference behaviour (mmmodel.py:42); torch.bfloat16 = bandwidth mode
        for p in self.llm.parameters():
            p.requires_grad = False

    def train(self, mode=True):
        if not isinstance(mode, bool):
            raise ValueError("training mode is expected to be boolean")
        self.training = mode
        for child in self.children():
            child.train(mode)
        self.llm.eval()
        return self

    def _run_llm(self, fn):
        # the reference enables autocast only when ``self.device == 'cuda'`` (mmmodel.py:53-55): with 'cuda:0' or a torch.device the
        # LLM runs WITHOUT autocast there, and so it does here (loss numerics follow the oracle for every spelling of the device)
        if str(self.device) == "cuda":
            with torch.amp.autocast("cuda"):
                return fn()
        return fn()

    def _loss(self, out_embeds, input_ids, attention_masks, labels):
        embeds, _mask, lab = splice_prefix(out_embeds, self.embeddings.weight, input_ids, attention_masks, labels, self.embeds_dtype)
        # like the reference, the LLM is called WITHOUT the attention mask (mmmodel.py:51; the mask built above is unused there too)
        outputs = self._run_llm(lambda: self.llm(inputs_embeds=embeds, labels=lab))
        return outputs.loss

    def _generate(self, out_embeds, max_new_tokens, prefix):
        B, H = out_embeds.shape
        if prefix is not None:
            embeds, _, _ = splice_prefix(out_embeds, self.embeddings.weight, prefix, None, None, self.embeds_dtype)
        else:
            embeds = out_embeds.unsqueeze(1)
        with torch.no_grad():
            return self._run_llm(lambda: self.llm.generate(inputs_embeds=embeds, max_new_tokens=max_new_tokens, pad_token_id=self.pad_token_id))


class HypernetMMModel(_MMBase):
    def __init__(self, llm, hypernet, device, mm_emb_dim, name, pad_token_id):
        super().__init__(llm, device, mm_emb_dim, name, pad_token_id)
        self.hypernet = hypernet

    def forward(self, mm_embeds, mm_stat_embeds, input_ids, attention_masks, labels):
        out_embeds = self.hypernet(mm_embeds, mm_stat_embeds)
        return self._loss(out_embeds, input_ids, attention_masks, labels), out_embeds

    def generate(self, mm_embeds, mm_subset_embeds, max_new_tokens, prefix=None):
        return self._generate(self.hypernet(mm_embeds, mm_subset_embeds), max_new_tokens, prefix)


class ProjectorMMModel(_MMBase):
    def __init__(self, llm, projector, device, mm_emb_dim, name, pad_token_id):
        super().__init__(llm, device, mm_emb_dim, name, pad_token_id)
        self.projector = projector

    def forward(self, mm_embeds, input_ids, attention_masks, labels):
        return self._loss(self.projector(mm_embeds), input_ids, attention_masks, labels)

    def generate(self, mm_embeds, max_new_tokens, prefix=None):
        return self._generate(self.projector(mm_embeds), max_new_tokens, prefix)


class LoraMMModel(_MMBase):
    def __init__(self, llm, lora_model, device, mm_emb_dim, name, pad_token_id):
        super().__init__(llm, device, mm_emb_dim, name, pad_token_id)
        self.lora_model = lora_model

    def forward(self, mm_embeds, input_ids, attention_masks, labels):
        return self._loss(self.lora_model(mm_embeds), input_ids, attention_masks, labels)

    def generate(self, mm_embeds, max_new_tokens, prefix=None):
        return self._generate(self.lora_model(mm_embeds), max_new_tokens, prefix)
