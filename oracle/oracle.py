"""CPU oracle for the adapted-projector hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU in plain torch/numpy fp32 (fp64 on request), the arithmetic of the
reference's hot path.  It exists so that the CUDA kernels can be
checked on a machine where ``/root/reference`` is not present.  Nothing in the product package may
import it: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference
arm do.

Parity pinning: the reference ships no tests and no golden vectors ("parity unpinned" by its own
suite).  The oracle is therefore pinned against the reference ITSELF: ``oracle/make_golden.py`` imports
the reference's modules from ``/root/reference`` (in the build container), runs them on seeded inputs
and stores inputs + outputs + gradients under ``tests/golden/``;
``tests/test_oracle_golden.py`` replays those vectors through this file.

All functions take parameters as a dict keyed with the reference's state-dict names
(``projector.net.0.weight`` ... ``hypernet.generators.1.bias``) so that a reference checkpoint
can be fed unchanged.  Citations are ``file:line`` relative to the reference repo root.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

Tensor = torch.Tensor

SQRT_2_OVER_PI = math.sqrt(2.0 / math.pi)
GELU_C3 = 0.044715


# ----------------------------------------------------------------------------------------------
# a1  EmbeddingManager.get_embeddings  (dmi/utils/model_utils.py:47-62)
# ----------------------------------------------------------------------------------------------
def l2_normalize(x: Tensor) -> Tensor:
    """Row-wise x / ||x||_2, no epsilon (model_utils.py:54-55,59)."""
    return x / x.norm(dim=1, keepdim=True)


# ----------------------------------------------------------------------------------------------
# a2  HypernetTrainer._get_rotation_matrix  (dmi/train_hypernet.py:56-57)
#     algorithm lives in scipy==1.14.1 scipy.stats.ortho_group.rvs (un-vendored dependency):
#     Gaussian matrix -> numpy.linalg.qr -> multiply column j of Q by sign(R_jj).
# ----------------------------------------------------------------------------------------------
def ortho_group_rvs(dim: int, random_state: np.random.RandomState) -> np.ndarray:
    z = random_state.normal(size=(dim, dim))
    q, r = np.linalg.qr(z)
    d = r.diagonal()
    q = q * (d / np.abs(d))[np.newaxis, :]
    return q


def get_rotation_matrix(dim: int, random_state: np.random.RandomState) -> Tensor:
    """float64 Haar draw cast to float32, as torch.FloatTensor(ortho_group.rvs(d)) does."""
    return torch.from_numpy(ortho_group_rvs(dim, random_state)).to(torch.float32)


# ----------------------------------------------------------------------------------------------
# a3  _interleave_embeddings / _process_embeddings  (dmi/train_hypernet.py:76-108)
# ----------------------------------------------------------------------------------------------
def interleave_embeddings(mm: Tensor, txt: Tensor) -> Tensor:
    """rows m0,t0,m1,t1,... (train_hypernet.py:76-83)."""
    k, d = mm.shape
    out = torch.empty(2 * k, d, dtype=mm.dtype)
    out[0::2] = mm
    out[1::2] = txt
    return out


def process_embeddings(mm_embs: Tensor, support: Tuple[Tensor, Tensor, Tensor], R: Optional[Tensor],
                       prune: Optional[int] = None, finetune_mm_dim: Optional[int] = None
                       ) -> Tuple[Tensor, Tensor]:
    """feed_txt_embs=True branch of _process_embeddings (train_hypernet.py:85-108).

    R is None when can_rotate/augment_emb_space is off.  Text rows and the instruction-prefix row
    are never rotated (:96-97).  With a pruned projector the support rows are zero-padded on the
    right up to finetune_mm_dim (:99-100).
    """
    m, t, p = support
    if R is not None:
        mm_embs = mm_embs @ R
        m = m @ R
    if prune is not None:
        m = torch.nn.functional.pad(m, (0, finetune_mm_dim - prune, 0, 0))
    z = torch.cat([p, interleave_embeddings(m, t)], dim=0)
    return mm_embs, z


# ----------------------------------------------------------------------------------------------
# a4  sinusoidal positional encoding  (dmi/model/hypernet.py:16-43)
# ----------------------------------------------------------------------------------------------
def sinusoidal_pe(d_model: int, max_len: int) -> Tensor:
    """[1, max_len, d_model] buffer, already scaled by 1/sqrt(d_model) (hypernet.py:31-32)."""
    pe = torch.zeros(max_len, d_model)
    pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float) * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return (pe * (1.0 / math.sqrt(d_model))).unsqueeze(0)


# ----------------------------------------------------------------------------------------------
# a5  MultiheadSelfAttention.forward  (dmi/model/hypernet.py:46-82), n_heads == 1 restated with
#     all S query rows exactly like the reference (the CUDA path only evaluates rows 0..1).
# ----------------------------------------------------------------------------------------------
def self_attention(seq: Tensor, wq, bq, wk, bk, wv, bv, n_valid: Optional[int] = None,
                   keep_mask: Optional[Tensor] = None, p_drop: float = 0.05, n_heads: int = 1) -> Tensor:
    """seq [S, d] -> [S, d].  Keys >= n_valid get -inf (hypernet.py:72-73).  keep_mask [heads,S,S]
    (1 = keep) reproduces nn.Dropout on the attention weights (:76) when given."""
    s, d = seq.shape
    hd = d // n_heads
    q = (seq @ wq.T + bq).view(s, n_heads, hd).transpose(0, 1)
    k = (seq @ wk.T + bk).view(s, n_heads, hd).transpose(0, 1)
    v = (seq @ wv.T + bv).view(s, n_heads, hd).transpose(0, 1)
    scores = q @ k.transpose(-2, -1) / math.sqrt(d)      # scale uses d_model, not head_dim (:70)
    if n_valid is not None and n_valid < s:
        scores[:, :, n_valid:] = float("-inf")
    w = torch.softmax(scores, dim=-1)
    if keep_mask is not None:
        w = w * keep_mask.to(w.dtype) / (1.0 - p_drop)
    out = (w @ v).transpose(0, 1).reshape(s, d)
    return out


# ----------------------------------------------------------------------------------------------
# a6  HyperNetwork.forward  (dmi/model/hypernet.py:140-196), hn_arch == "attention"
# ----------------------------------------------------------------------------------------------
def hypernetwork_forward(params: Dict[str, Tensor], z: Tensor, *, n_tokens: int, rank: int, alpha: float,
                         lm_dim: int, mm_dim: int, n_proj_layers: int = 2, predict_bias: bool = True,
                         use_pos_encs: bool = True, keep_mask: Optional[Tensor] = None,
                         n_heads: int = 1, prefix: str = "hypernet."):
    """Returns (a_weights, b_weights, biases) as flat tensors, like the reference."""
    g = lambda k: params[prefix + k]
    ptok = g("prefix_tokens")
    hyp_dim = ptok.shape[1]
    n_pref = ptok.shape[0]
    seq_len = n_pref + z.shape[0]
    ctx = 2 * n_tokens + n_pref + 1
    if seq_len < ctx:                                              # :144-151
        pad = torch.zeros(ctx - seq_len, z.shape[1], dtype=z.dtype)
        seq = torch.cat([ptok, z, pad], dim=0)
        n_valid = seq_len
    else:
        seq = torch.cat([ptok, z], dim=0)                          # :163
        n_valid = None
    if use_pos_encs:                                               # :166-167 / :41-43
        pe = params.get(prefix + "pos_encs.pe")
        if pe is None:
            pe = sinusoidal_pe(hyp_dim, ctx)
        seq = seq + pe[0, : seq.shape[0]].to(seq.dtype)
    enc = self_attention(seq, g("hypnet.q.weight"), g("hypnet.q.bias"), g("hypnet.k.weight"),
                         g("hypnet.k.bias"), g("hypnet.v.weight"), g("hypnet.v.bias"),
                         n_valid=n_valid, keep_mask=keep_mask, n_heads=n_heads)
    pref = enc[:n_pref]                                            # :175
    a_w, b_w, biases = [], [], ([] if predict_bias else None)
    for idx in range(n_proj_layers):                               # :181-194
        a_dim = (hyp_dim if idx == 0 else lm_dim) * rank
        b_dim = rank * lm_dim
        w = (alpha / rank) * (g(f"generators.{idx}.weight") @ pref[idx] + g(f"generators.{idx}.bias"))
        a = w[:a_dim]
        b = w[a_dim:a_dim + b_dim]
        if idx == 0 and hyp_dim > mm_dim:                          # :187-188
            a = a[: mm_dim * rank]
        a_w.append(a)
        b_w.append(b)
        if predict_bias:
            biases.append(w[a_dim + b_dim:])
    return a_w, b_w, biases


# ----------------------------------------------------------------------------------------------
# a7-a10  Projector  (dmi/model/projector.py)
# ----------------------------------------------------------------------------------------------
def gelu_tanh(x: Tensor) -> Tensor:
    """nn.GELU(approximate='tanh') — what proj_act='quick_gelu' really builds (projector.py:17-22,32)."""
    return 0.5 * x * (1.0 + torch.tanh(SQRT_2_OVER_PI * (x + GELU_C3 * x * x * x)))


def projector_forward(params: Dict[str, Tensor], x: Tensor, drop_keep: Optional[Tensor] = None,
                      p_drop: float = 0.1, prefix: str = "projector.") -> Tensor:
    """Projector.forward, MLP2 (projector.py:56-59): Linear -> GELU(tanh) -> Dropout -> Linear."""
    h = gelu_tanh(x @ params[prefix + "net.0.weight"].T + params[prefix + "net.0.bias"])
    if drop_keep is not None:
        h = h * drop_keep.to(h.dtype) / (1.0 - p_drop)
    return h @ params[prefix + "net.3.weight"].T + params[prefix + "net.3.bias"]


def adapted_mlp_full(w1, b1, w2, b2, x, a_w: Sequence[Tensor], b_w: Sequence[Tensor],
                     biases: Optional[Sequence[Tensor]]) -> Tensor:
    """The complete 2-layer adapted MLP (what combine_lora / only_lora_forward compute,
    projector.py:61-116): y = gelu(x W1^T + b1 + (x A0) B0 + beta0) W2^T + b2 + (h A1) B1 + beta1."""
    a0 = a_w[0].reshape(w1.shape[1], -1)
    b0 = b_w[0].reshape(-1, w1.shape[0])
    a1 = a_w[1].reshape(w2.shape[1], -1)
    bb1 = b_w[1].reshape(-1, w2.shape[0])
    beta0 = biases[0] if biases is not None else 0.0
    beta1 = biases[1] if biases is not None else 0.0
    h = gelu_tanh(x @ w1.T + b1 + (x @ a0) @ b0 + beta0)
    return h @ w2.T + b2 + (h @ a1) @ bb1 + beta1


def lora_forward_as_written(params: Dict[str, Tensor], x: Tensor, a_w, b_w, biases,
                            prefix: str = "projector.") -> Tensor:
    """Projector.lora_forward exactly as written (projector.py:118-159).

    ``zip(self.net, a_weights, b_weights, biases)`` pairs the 4 modules of ``net`` with the
    n_proj_layers=2 weight lists, so iteration stops after (Linear0, GELU): the second Linear, the
    Dropout and the second adapter are never applied.  Returns gelu(pre)."""
    w1 = params[prefix + "net.0.weight"]
    b1 = params[prefix + "net.0.bias"]
    if biases is None:
        biases = [torch.zeros(w1.shape[0], dtype=x.dtype) for _ in a_w]
    a0 = a_w[0].reshape(w1.shape[1], -1)
    b0 = b_w[0].reshape(-1, w1.shape[0])
    pre = x @ w1.T + b1 + ((x @ a0) @ b0 + biases[0])
    if len(a_w) < 2:            # zip would stop after the first Linear
        return pre
    return gelu_tanh(pre)


def only_lora_forward(params: Dict[str, Tensor], x: Tensor, loras: Sequence[Tuple[Tensor, Tensor]],
                      alpha: float, rank: int, prefix: str = "projector.") -> Tensor:
    """Projector.only_lora_forward + LoRALayer.forward (projector.py:61-74, lora.py:15-17).
    Eval-mode projector (LoraWrapper.train keeps it in eval, lora.py:47-55) so Dropout is identity."""
    s = alpha / rank
    w1, b1 = params[prefix + "net.0.weight"], params[prefix + "net.0.bias"]
    w2, b2 = params[prefix + "net.3.weight"], params[prefix + "net.3.bias"]
    (a0, b0), (a1, bb1) = loras
    h = gelu_tanh(x @ w1.T + b1 + s * (x @ a0 @ b0))
    return h @ w2.T + b2 + s * (h @ a1 @ bb1)


def combine_lora(params: Dict[str, Tensor], a_w, b_w, biases, prefix: str = "projector."
                 ) -> Dict[str, Tensor]:
    """Projector.combine_lora (projector.py:76-116): W' = (A B)^T + W, b' = beta + b for each Linear.
    Returns the merged nn.Sequential's state-dict keys ('0.weight','0.bias','3.weight','3.bias')."""
    lin = [(prefix + "net.0.weight", prefix + "net.0.bias", "0"),
           (prefix + "net.3.weight", prefix + "net.3.bias", "3")]
    if len(a_w) < len(lin):
        raise ValueError("Not enough weights provided for all linear layers")
    if len(a_w) > len(lin):
        raise ValueError("Too many weights provided")
    out = {}
    for i, (wk, bk, name) in enumerate(lin):
        w, b = params[wk], params[bk]
        a = a_w[i].reshape(w.shape[1], -1)
        bm = b_w[i].reshape(-1, w.shape[0])
        beta = biases[i] if biases is not None else torch.zeros_like(b)
        out[name + ".weight"] = (a @ bm).T + w
        out[name + ".bias"] = beta + b
    return out


def merged_forward(merged: Dict[str, Tensor], x: Tensor) -> Tensor:
    """nn.Sequential(Linear, GELU, Dropout(eval), Linear) produced by combine_lora."""
    h = gelu_tanh(x @ merged["0.weight"].T + merged["0.bias"])
    return h @ merged["3.weight"].T + merged["3.bias"]


def average_adapters(adapters: Sequence[Tuple[List[Tensor], List[Tensor], Optional[List[Tensor]]]]):
    """HyperNetWrapper.generate_projector_from_multiple_adapters mean step (hypernet.py:251-262)."""
    n_layers = len(adapters[0][0])
    a = [torch.stack([ad[0][i] for ad in adapters]).mean(0) for i in range(n_layers)]
    b = [torch.stack([ad[1][i] for ad in adapters]).mean(0) for i in range(n_layers)]
    bias = None
    if adapters[0][2] is not None:
        bias = [torch.stack([ad[2][i] for ad in adapters]).mean(0) for i in range(n_layers)]
    return a, b, bias


# ----------------------------------------------------------------------------------------------
# a11  HyperNetWrapper.forward  (dmi/model/hypernet.py:268-274)
# ----------------------------------------------------------------------------------------------
def hypernet_wrapper_forward(params, x, z, *, n_tokens, rank, alpha, lm_dim, mm_dim,
                             keep_mask=None, full_mlp: bool = False, **kw) -> Tensor:
    a_w, b_w, biases = hypernetwork_forward(params, z, n_tokens=n_tokens, rank=rank, alpha=alpha,
                                            lm_dim=lm_dim, mm_dim=mm_dim, keep_mask=keep_mask, **kw)
    if full_mlp:
        return adapted_mlp_full(params["projector.net.0.weight"], params["projector.net.0.bias"],
                                params["projector.net.3.weight"], params["projector.net.3.bias"],
                                x, a_w, b_w, biases)
    return lora_forward_as_written(params, x, a_w, b_w, biases)


# ----------------------------------------------------------------------------------------------
# a12  prefix splice  (dmi/model/mmmodel.py:36-48, same block at :118-135 and :205-221)
# ----------------------------------------------------------------------------------------------
def splice_prefix(projected: Tensor, embed_table: Tensor, input_ids: Tensor, attention_masks: Tensor,
                  labels: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """inputs_embeds = cat(projected[:,None,:], embed(ids)); mask gets a leading 1, labels a leading -100.
    torch.cat type-promotes (fp32 projected + bf16 table -> fp32 inputs_embeds, mmmodel.py:42)."""
    b = projected.shape[0]
    text = embed_table[input_ids]
    embeds = torch.cat((projected.unsqueeze(1), text), dim=1)
    mask = torch.cat((torch.ones(b, 1), attention_masks), dim=-1)
    lab = torch.cat((torch.full((b, 1), -100), labels), dim=-1)
    return embeds, mask, lab


# ----------------------------------------------------------------------------------------------
# a13  gradients: autograd over the restatement (the reference uses autograd too)
# ----------------------------------------------------------------------------------------------
def adapted_mlp_full_grads(w1, b1, w2, b2, x, a_w, b_w, biases, dy, base_grads: bool = False):
    """y and d(sum(y*dy)) w.r.t. the adapter factors (and optionally W1,b1,W2,b2)."""
    leaves = [t.detach().clone().requires_grad_(True) for t in (*a_w, *b_w, *biases)]
    n = len(a_w)
    base = [t.detach().clone().requires_grad_(base_grads) for t in (w1, b1, w2, b2)]
    y = adapted_mlp_full(*base, x, leaves[:n], leaves[n:2 * n], leaves[2 * n:])
    targets = leaves + (base if base_grads else [])
    grads = torch.autograd.grad((y * dy).sum(), targets)
    return y.detach(), list(grads)


def gelu_tanh_grad(a: Tensor) -> Tensor:
    """d gelu_tanh / da."""
    t = torch.tanh(SQRT_2_OVER_PI * (a + GELU_C3 * a ** 3))
    return 0.5 * (1 + t) + 0.5 * a * (1 - t * t) * SQRT_2_OVER_PI * (1 + 3 * GELU_C3 * a * a)


# ----------------------------------------------------------------------------------------------
# a13 / 8f-1  optimizer step: clip_grad_norm_ + AdamW  (dmi/train_hypernet.py:148-149, optimizer built at :526-532)
# The arithmetic lives in torch (torch.nn.utils.clip_grad_norm_, torch.optim.AdamW single-tensor path); restated here
# operation by operation so that it can be checked without torch's optimizer classes.
# ----------------------------------------------------------------------------------------------
def clip_grad_norm(grads: Sequence[Tensor], max_norm: float) -> Tuple[List[Tensor], Tensor]:
    """returns (clipped gradients, total 2-norm): coef = clamp(max_norm / (total + 1e-6), max=1)"""
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g, 2.0) for g in grads]), 2.0)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return [g * coef for g in grads], total


def adamw_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, *, lr: float, betas=(0.9, 0.999), eps: float = 1e-8,
               weight_decay: float = 1e-2) -> Tuple[Tensor, Tensor, Tensor]:
    """one torch.optim.AdamW update (amsgrad=False); ``step`` counts from 1.  Returns new (p, exp_avg, exp_avg_sq)."""
    b1, b2 = betas
    p = p * (1.0 - lr * weight_decay)
    m = torch.lerp(m, g, 1.0 - b1)
    v = v * b2 + (1.0 - b2) * g * g
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


# ----------------------------------------------------------------------------------------------
# 8f-4  embedding side of the collate functions + get_embeddings  (dmi/data/base.py:222-232, dmi/utils/model_utils.py:47-62)
# ----------------------------------------------------------------------------------------------
def collate_embeddings(items: Sequence[dict], selected_features=None, emb_mean: Optional[Tensor] = None, normalize: bool = True,
                       emb_name: str = "emb") -> Tensor:
    """train_collate's embedding part, literally: per-item FloatTensor (+ feature selection), stack, subtract mean; then the
    row L2 normalisation of EmbeddingManager.get_embeddings."""
    if selected_features is not None:
        embs = [torch.FloatTensor(item[emb_name])[selected_features] for item in items]
    else:
        embs = [torch.FloatTensor(item[emb_name]) for item in items]
    embs = torch.stack(embs, dim=0)
    if emb_mean is not None:
        embs = embs - emb_mean
    if normalize:
        embs = embs / embs.norm(dim=1, keepdim=True)
    return embs


# ----------------------------------------------------------------------------------------------
# 8f-3  Haar orthogonal matrix from Gaussian samples without a matrix QR (statistically equivalent to a2)
# ----------------------------------------------------------------------------------------------
def haar_from_gaussian(gauss: np.ndarray) -> np.ndarray:
    """Q = H_0 H_1 ... H_{n-1} diag(d) in float64, with v_k = x_k + sign(x_k0)||x_k|| e_0 built from row k of ``gauss``
    (entries k..n-1) and d_k = -sign(x_k0): the reflectors and sign fix that Householder QR + ``q *= sign(diag(r))``
    (scipy.stats.ortho_group.rvs, used at train_hypernet.py:57) apply to a Gaussian matrix, whose k-th reduced column is an
    independent Gaussian vector (Stewart 1980)."""
    g = np.asarray(gauss, dtype=np.float64)
    n = g.shape[0]
    Q = np.eye(n)
    d = np.empty(n)
    for k in range(n - 1, -1, -1):
        x = g[k, k:].copy()
        sgn = 1.0 if x[0] >= 0 else -1.0
        d[k] = -sgn
        v = x
        v[0] += sgn * np.linalg.norm(g[k, k:])
        tau = 2.0 / (v @ v)
        Q[k:, :] -= tau * np.outer(v, v @ Q[k:, :])
    return Q * d[None, :]
