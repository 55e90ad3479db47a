"""Device-resident embedding store -- the embedding side of the reference's data path.

The reference keeps extracted embeddings as a pickled ``{id: {'caption': ..., 'emb': [...]}}`` dict, turns it into a
``datasets.Dataset`` and rebuilds every batch on the host, one ``torch.FloatTensor(item['emb'])[selected_features]`` per sample,
``torch.stack``, ``- emb_mean`` (``dmi/data/base.py:159-185, 222-268``), before ``EmbeddingManager.get_embeddings`` moves it to the
GPU and L2-normalises it (``dmi/utils/model_utils.py:47-62``).  Here the table is ONE flat ``[N, D]`` tensor in HBM (a full
extracted-feature split is a few GB; 180 GB are available) and a batch is a single gather kernel over sample indices that also
applies the feature selection, the mean subtraction and the normalisation, and can emit the bf16 operand of the projector GEMM
directly.  Tokenisation / captions stay on the host and are out of scope."""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Mapping, Optional, Sequence

import numpy as np
import torch

from . import _lib, ops
from ._lib import AUG_NORMALIZE


class EmbeddingStore:
    def __init__(self, table: torch.Tensor, ids: Optional[Sequence] = None, *, selected_features=None, mean: Optional[torch.Tensor] = None):
        """table: [N, D_store] fp32 or bf16 CUDA tensor; ids: the reference's sample keys in row order (optional);
        selected_features: column indices kept by ``_select_features`` (InfFS, base.py:222-225) or None; mean: ``emb_mean`` of
        ``subtract_mean`` (already restricted to the selected features, as in the reference) or None."""
        ops._need_cuda(table)
        assert table.dim() == 2 and table.stride(1) == 1 and table.dtype in (torch.float32, torch.bfloat16)
        self.table = table
        self.ids = list(ids) if ids is not None else None
        self.row_of = {k: i for i, k in enumerate(self.ids)} if self.ids is not None else None
        dev = table.device
        self.selected = None if selected_features is None else torch.as_tensor(np.asarray(selected_features), dtype=torch.int32, device=dev).contiguous()
        if self.selected is not None:
            assert int(self.selected.min()) >= 0 and int(self.selected.max()) < table.shape[1], "selected feature outside the stored width"
        self.d_out = table.shape[1] if self.selected is None else self.selected.numel()
        self.mean = None if mean is None else mean.to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
        if self.mean is not None:
            assert self.mean.numel() == self.d_out
        self._errf = ops.DeferredErrorFlag(dev, "EmbeddingStore.gather: sample index outside the table")

    @classmethod
    def from_items(cls, items: Mapping, emb_name: str = "emb", device="cuda", dtype=torch.float32, **kw) -> "EmbeddingStore":
        """``items`` is the reference's pickled split dict ``{id: {'caption': str, emb_name: sequence of floats}}`` (base.py:159-177)"""
        ids = list(items.keys())
        table = torch.from_numpy(np.asarray([np.asarray(items[k][emb_name], dtype=np.float32) for k in ids], dtype=np.float32))
        return cls(table.to(device=device, dtype=dtype), ids, **kw)

    def __len__(self) -> int:
        return self.table.shape[0]

    def rows(self, ids: Iterable) -> torch.Tensor:
        """sample keys -> row indices (int64 on the store's device)"""
        assert self.row_of is not None, "the store was built without ids"
        return torch.tensor([self.row_of[k] for k in ids], dtype=torch.int64, device=self.table.device)

    def gather(self, idx: Optional[torch.Tensor], *, normalize: bool = True, out: Optional[torch.Tensor] = None,
               out_bf16: Optional[torch.Tensor] = None, want_f32: bool = True, check: bool = False):
        """batch of embeddings for sample rows ``idx`` (int64 CUDA tensor, or None for rows 0..len-1 of ``out``):
        column gather, mean subtraction, L2 normalisation fused.  Returns the fp32 batch (and fills ``out_bf16`` if given).
        An out-of-range index poisons its output row with NaN and raises ``IndexError`` -- immediately with ``check=True`` (one
        synchronisation), otherwise at the next ``gather`` / ``check()`` call once the asynchronous flag copy has landed."""
        dev = self.table.device
        self._errf.poll()                   # a bad index seen by an EARLIER launch raises here without synchronising
        if idx is not None:
            ops._need_cuda(idx)
            assert idx.dtype == torch.int64 and idx.dim() == 1 and idx.is_contiguous()
            B = idx.numel()
        else:
            B = (out if out is not None else out_bf16).shape[0]
        if out is None and want_f32:
            out = torch.empty(B, self.d_out, dtype=torch.float32, device=dev)
        for t in (out, out_bf16):
            if t is not None:
                ops._need_cuda(t)
                assert t.shape[0] == B and t.shape[1] >= self.d_out and t.stride(1) == 1
        assert out is None or out.dtype == torch.float32
        assert out_bf16 is None or out_bf16.dtype == torch.bfloat16
        rc = _lib.load().dmi_gather_rows(ops._ptr(self.table), int(self.table.dtype == torch.bfloat16), self.table.stride(0), self.table.shape[0],
                                         self.table.shape[1], ops._ptr(idx), B, self.d_out, ops._ptr(self.selected), ops._ptr(self.mean),
                                         AUG_NORMALIZE if normalize else 0, ops._ptr(out), 0 if out is None else out.stride(0),
                                         ops._ptr(out_bf16), 0 if out_bf16 is None else out_bf16.stride(0), ops._ptr(self._errf.flag), ops._stream())
        _lib.check(rc, "dmi_gather_rows")
        if check:
            self._errf.check()
        else:
            self._errf.arm()
        return out

    def check(self) -> None:
        """blocking check of the deferred out-of-range flag (end of step, or after replaying a captured graph)"""
        self._errf.check()
