"""Embedding preparation of the hypernetwork path -- mirrors the three helpers of ``HypernetTrainer``
(``dmi/train_hypernet.py:56-108``) and ``EmbeddingManager.get_embeddings`` (``dmi/utils/model_utils.py:47-62``).

In the reference these are separate ATen calls (norm, div, 2 matmuls, pad, stack/transpose/reshape, cat).  Here one C-ABI call
(``dmi_augment``) normalises, gathers/sign-flips, rotates on the tensor cores (3xTF32, fp32-accurate) and writes the rotated
support rows straight into their interleaved slots of ``z``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib, ops
from ._lib import AUG_NORMALIZE, AugmentArgs


def get_rotation_matrix(mm_dim: int, device, random_state=None) -> torch.Tensor:
    """Haar-random orthogonal matrix exactly as the reference draws it (train_hypernet.py:56-57):
    ``torch.FloatTensor(scipy.stats.ortho_group.rvs(mm_dim)).to(device)`` -- host LAPACK QR on the NumPy global RNG (or the
    given RandomState), so that a seeded run reproduces the reference's R bit for bit.  R is data for the kernels."""
    from scipy.stats import ortho_group
    return torch.from_numpy(np.asarray(ortho_group.rvs(mm_dim, random_state=random_state))).to(torch.float32).to(device)


def get_rotation_matrix_device(mm_dim: int, device, generator: Optional[torch.Generator] = None, gauss: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Haar-random orthogonal [mm_dim, mm_dim] matrix drawn ON THE DEVICE (``dmi_haar_orthogonal``): same distribution as
    ``ortho_group.rvs`` (Householder reflectors of independent Gaussian directions, sign-fixed), none of the host LAPACK QR that
    costs the reference 72 ms (D=768) to 0.7 s (D=2048) per micro-step.  Not bit-comparable with the host draw -- use
    ``get_rotation_matrix`` for a seeded run that must reproduce the reference's R.  ``gauss`` may supply the N(0,1) samples."""
    if gauss is None:
        gauss = torch.randn(mm_dim, mm_dim, device=device, dtype=torch.float32, generator=generator)
    ops._need_cuda(gauss)
    assert gauss.shape == (mm_dim, mm_dim) and gauss.dtype == torch.float32 and gauss.is_contiguous()
    lib = _lib.load()
    nbytes = int(lib.dmi_haar_workspace_bytes(mm_dim))
    ws = torch.empty(nbytes // 4, dtype=torch.float32, device=gauss.device)
    Q = torch.empty(mm_dim, mm_dim, dtype=torch.float32, device=gauss.device)
    _lib.check(lib.dmi_haar_orthogonal(ops._ptr(gauss), mm_dim, ops._ptr(Q), ops._ptr(ws), nbytes, ops._stream()), "dmi_haar_orthogonal")
    return Q


def l2_normalize(x: torch.Tensor) -> torch.Tensor:
    """x / x.norm(dim=1, keepdim=True) (model_utils.py:54-59)"""
    ops._need_cuda(x)
    x = x.float().contiguous()
    out = torch.empty_like(x)
    rc = _lib.load().dmi_l2_normalize(ops._ptr(x), x.stride(0), x.shape[0], x.shape[1], ops._ptr(out), out.stride(0), ops._stream())
    _lib.check(rc, "dmi_l2_normalize")
    return out


def get_embeddings(inputs, feed_txt_embs: bool, device):
    """``EmbeddingManager.get_embeddings`` for ``load_extracted_features=True``: move to the device, L2-normalise every row.
    ``inputs`` is a tensor, or the support tuple ``(embs, text_embs, prefix_emb)`` when ``feed_txt_embs``."""
    if feed_txt_embs and isinstance(inputs, (list, tuple)):
        embs, text_embs, prefix_emb = (t.to(device, non_blocking=True) for t in inputs)
        return l2_normalize(embs), l2_normalize(text_embs), l2_normalize(prefix_emb)
    return l2_normalize(inputs.to(device, non_blocking=True))


def interleave_embeddings(mm_subset_membs: torch.Tensor, txt_embs: torch.Tensor) -> torch.Tensor:
    """rows m0,t0,m1,t1,... (train_hypernet.py:76-83) -- pure data movement, done by dmi_augment without rotation"""
    _, z = process_embeddings(None, (mm_subset_membs, txt_embs, None), R=None, feed_txt_embs=True)
    return z


def process_embeddings(mm_embs: Optional[torch.Tensor], mm_subset_embs, *, R: Optional[torch.Tensor] = None, feed_txt_embs: bool = True,
                       prune: Optional[int] = None, finetune_mm_dim: Optional[int] = None, normalize: bool = False,
                       perm: Optional[torch.Tensor] = None, sign: Optional[torch.Tensor] = None,
                       mm_out_bf16: Optional[torch.Tensor] = None) -> Tuple[Optional[torch.Tensor], torch.Tensor]:
    """``HypernetTrainer._process_embeddings`` (train_hypernet.py:85-108).

    ``R`` is the rotation to apply (None when ``can_rotate`` / ``augment_emb_space`` is off).  With ``feed_txt_embs`` the
    support argument is the tuple ``(mm_subset_membs [K,D], txt_embs [K,Dh], prefix_emb [1,Dh] | None)`` and the result is
    ``(mm_embs', z [1+2K, Dh])``; without it nothing is assembled (the reference returns its inputs unchanged in that case).
    ``normalize=True`` folds the L2 normalisation of ``get_embeddings`` into the same call.  ``perm`` / ``sign`` are the
    optional column gather / sign flip of the isometry (identity / +1 reproduces the reference)."""
    if not feed_txt_embs:
        return mm_embs, mm_subset_embs
    m, t, p = mm_subset_embs
    dev = m.device
    ops._need_cuda(mm_embs, m, t, p, R, perm, sign, mm_out_bf16)
    K = m.shape[0]
    D_src = m.shape[1]
    D = D_src if perm is None else perm.numel()
    if prune is not None:
        assert finetune_mm_dim is not None and D == prune, "pruned projector: support rows are padded from `prune` to finetune_mm_dim"
        Dh = finetune_mm_dim
    else:
        Dh = t.shape[1]
    assert t.shape == (K, Dh) and Dh >= D
    B = 0 if mm_embs is None else mm_embs.shape[0]
    a = AugmentArgs()
    a.B, a.K, a.D, a.Dh, a.D_src = B, K, D, Dh, D_src
    a.flags = AUG_NORMALIZE if normalize else 0
    keep = []          # keep temporaries alive until the call is enqueued

    def f32(x):
        x = x.float().contiguous()
        keep.append(x)
        return x
    mm_out = None
    if B:
        mm_embs = f32(mm_embs)
        assert mm_embs.shape[1] == D_src
        mm_out = torch.empty(B, D, dtype=torch.float32, device=dev)
        a.mm, a.ld_mm = mm_embs.data_ptr(), mm_embs.stride(0)
        a.mm_out, a.ld_mm_out = mm_out.data_ptr(), mm_out.stride(0)
        if mm_out_bf16 is not None:
            assert mm_out_bf16.dtype == torch.bfloat16 and mm_out_bf16.shape[0] == B and mm_out_bf16.stride(1) == 1
            a.mm_out_bf16, a.ld_mm_bf16 = mm_out_bf16.data_ptr(), mm_out_bf16.stride(0)
    m, t = f32(m), f32(t)
    a.sup, a.ld_sup = m.data_ptr(), m.stride(0)
    a.txt, a.ld_txt = t.data_ptr(), t.stride(0)
    n_rows = 2 * K + (1 if p is not None else 0)
    zbuf = torch.empty(1 + 2 * K, Dh, dtype=torch.float32, device=dev)
    a.z = zbuf.data_ptr()
    if p is not None:
        p = f32(p)
        assert p.shape == (1, Dh)
        a.prefix = p.data_ptr()
    if R is not None:
        R = f32(R)
        assert R.shape == (D, D)
        a.R = R.data_ptr()
        nbytes = int(_lib.load().dmi_augment_workspace_bytes(B, K, D))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        keep.append(ws)
        a.workspace, a.workspace_bytes = ws.data_ptr(), nbytes
    if perm is not None:
        perm = perm.to(torch.int32).contiguous()
        keep.append(perm)
        a.perm = perm.data_ptr()
    if sign is not None:
        sign = f32(sign)
        a.sign = sign.data_ptr()
    _lib.check(_lib.load().dmi_augment(C.byref(a), ops._stream()), "dmi_augment")
    for x in keep:
        x.record_stream(torch.cuda.current_stream())
    z = zbuf if p is not None else zbuf[1:]
    assert z.shape[0] == n_rows
    return mm_out, z
