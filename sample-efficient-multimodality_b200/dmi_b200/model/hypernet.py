"""Hypernetwork that turns a K-shot support set into per-layer LoRA factors -- drop-in for ``dmi/model/hypernet.py``.

Same classes, constructor signatures, parameter / buffer names (``prefix_tokens``, ``hypnet.{q,k,v}.{weight,bias}``,
``generators.{l}.{weight,bias}``, ``pos_encs.pe``) and return conventions as the reference (hypernet.py:84-280), so a
reference checkpoint loads unchanged.  ``forward`` runs the fp32 sm_100a kernels behind ``dmi_hypernet_fwd/bwd``: the 1-head
attention pooling is evaluated for the two prefix rows only, in a reduced form that never materialises K/V projections of
the S support tokens, and the generators are streamed once from HBM.

Only ``hn_arch == "attention"`` with one head is built (the only architecture any reference config uses;
``"att_w_nonlinear"`` cannot run in the reference either).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional

import torch
import torch.nn as nn

from .. import _lib, ops
from .._lib import MAX_GEN_LAYERS, HypernetArgs as _CArgs
from ..utils.args import HypnetArgs, setup_args
from .projector import Projector


def sinusoidal_pos_embedding(d_model: int, max_len: int = 5000, pos_offset: int = 0, device=None):
    """fixed sin/cos table [max_len, d_model] (hypernet.py:16-23); built once on the host, it is a constant buffer"""
    position = torch.arange(0, max_len, dtype=torch.float, device=device).unsqueeze(1) + pos_offset
    freq = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float, device=device) * (-math.log(10000.0) / d_model))
    table = torch.zeros(max_len, d_model, device=device)
    table[:, 0::2] = torch.sin(position * freq)
    table[:, 1::2] = torch.cos(position * freq)
    return table


class PositionalEncoding(nn.Module):
    """holds the ``pe`` buffer [1, max_len, d_model] scaled by 1/sqrt(d_model) (hypernet.py:26-43); the addition itself is
    fused into the pooling kernel"""

    def __init__(self, d_model, dropout=0.0, max_len=128, device=None):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        self.batch_dim = 0
        self.register_buffer("pe", (sinusoidal_pos_embedding(d_model, max_len, 0, device) / math.sqrt(d_model)).unsqueeze(0))

    def get(self, n: int, offset: int) -> torch.Tensor:
        return self.pe.narrow(1, start=offset, length=n)


class MultiheadSelfAttention(nn.Module):
    """parameter holder for the q/k/v projections and the attention-weight dropout (hypernet.py:46-55)"""

    def __init__(self, d_model=128, nhead=4, dropout=0.05):
        super().__init__()
        self.d_model = d_model
        self.nhead = nhead
        self.q = nn.Linear(d_model, d_model)
        self.k = nn.Linear(d_model, d_model)
        self.v = nn.Linear(d_model, d_model)
        self.dropout = nn.Dropout(dropout)


def _join_piece_grads(gs, cuts, dev):
    """gradients of the consecutive slices [cuts[i], cuts[i+1]) of one generated vector -> gradient of the vector (None if all None).
    Zero-copy when they already are consecutive slices of one fp32 buffer."""
    if all(g is None for g in gs):
        return None
    if all(g is not None and g.dtype == torch.float32 and g.is_contiguous() and g.numel() == cuts[i + 1] - cuts[i] for i, g in enumerate(gs)):
        g0 = gs[0]
        st = g0.untyped_storage()
        if all(g.untyped_storage().data_ptr() == st.data_ptr() and g.data_ptr() == g0.data_ptr() + 4 * cuts[i] for i, g in enumerate(gs)) \
                and st.nbytes() - 4 * g0.storage_offset() >= 4 * cuts[-1]:
            return torch.as_strided(g0, (cuts[-1],), (1,))
    full = torch.zeros(cuts[-1], dtype=torch.float32, device=dev)
    for i, g in enumerate(gs):
        if g is not None:
            full[cuts[i]:cuts[i + 1]].copy_(g.reshape(-1))
    return full


class _HyperNetFn(torch.autograd.Function):
    """(z, prefix_tokens, Wq,bq,Wk,bk,Wv,bv, G_0,c_0, G_1,c_1, ...) -> per layer (A_flat, B_flat[, beta]): consecutive slices of ONE
    generated vector per layer, returned as separate outputs so that their gradients arrive separately (no slice-backward kernels);
    when the three gradients are again consecutive slices of one buffer (``_AdaptedMLPFn.backward`` allocates them so) the backward
    kernel reads that buffer in place."""

    @staticmethod
    def forward(ctx, hn, z, keep, prefix_tokens, wq, bq, wk, bk, wv, bv, *gens):
        ctx.set_materialize_grads(False)
        dev = z.device
        n_layers = len(gens) // 2
        NQ, D = prefix_tokens.shape
        S_z = z.shape[0]
        a = _CArgs()
        a.S_z, a.NQ, a.D, a.n_layers = S_z, NQ, D, n_layers
        a.out_scale = float(hn.alpha) / float(hn.rank)
        z = z.detach().float().contiguous()
        a.z, a.ldz = z.data_ptr(), z.stride(0)
        a.prefix_tokens = prefix_tokens.data_ptr()
        pe = None
        if hn.use_pos_encs:
            pe = hn.pos_encs.pe[0]
            assert pe.shape[0] >= NQ + S_z, "support set longer than the positional-encoding table (n_tokens too small)"
            a.pe, a.ldpe = pe.data_ptr(), pe.stride(0)
        for name, t in (("wq", wq), ("bq", bq), ("wk", wk), ("bk", bk), ("wv", wv), ("bv", bv)):
            assert t.is_contiguous() and t.dtype == torch.float32
            setattr(a, name, t.data_ptr())
        outs = []
        for l in range(n_layers):
            gw, gb = gens[2 * l], gens[2 * l + 1]
            assert gw.is_contiguous() and gw.dtype == torch.float32 and gw.shape[1] == D
            a.gen_w[l], a.gen_b[l], a.gen_out[l] = gw.data_ptr(), gb.data_ptr(), gw.shape[0]
            o = torch.empty(gw.shape[0], dtype=torch.float32, device=dev)
            a.w_out[l] = o.data_ptr()
            outs.append(o)
        if keep is not None:
            keep = keep.float().contiguous()
            assert keep.shape == (NQ, NQ + S_z)
            a.keep, a.dropout_p = keep.data_ptr(), float(hn.hypnet.dropout.p)
        lib = _lib.load()
        stash = torch.empty(int(lib.dmi_hypernet_stash_floats(NQ, S_z, D)), dtype=torch.float32, device=dev)
        a.stash = stash.data_ptr()
        _lib.check(lib.dmi_hypernet_fwd(C.byref(a), ops._stream()), "dmi_hypernet_fwd")
        ctx.hn, ctx.args, ctx.n_layers = hn, a, n_layers
        ctx.keepalive = (z, keep, pe, stash, outs)
        ctx.params = (prefix_tokens, wq, bq, wk, bk, wv, bv) + tuple(gens)
        pieces, ctx.cuts = [], []
        for l, o in enumerate(outs):
            cuts = [0, hn.a_dims[l], hn.a_dims[l] + hn.b_dims[l]] + ([o.numel()] if hn.predict_bias else [])
            ctx.cuts.append(cuts)
            pieces += [o[cuts[i]:cuts[i + 1]] for i in range(len(cuts) - 1)]
        return tuple(pieces)

    @staticmethod
    def backward(ctx, *dpieces):
        a, hn, n_layers = ctx.args, ctx.hn, ctx.n_layers
        dws, k = [], 0
        for cuts in ctx.cuts:                  # reassemble one gradient vector per layer from the gradients of its pieces
            n = len(cuts) - 1
            dws.append(_join_piece_grads(dpieces[k:k + n], cuts, ctx.params[0].device))
            k += n
        prefix_tokens, wq, bq, wk, bk, wv, bv = ctx.params[:7]
        gens = ctx.params[7:]
        dev = prefix_tokens.device
        NQ, D = prefix_tokens.shape
        lib = _lib.load()
        small = dict(dprefix=prefix_tokens, dwq=wq, dbq=bq, dwk=wk, dbk=bk, dwv=wv, dbv=bv)
        fused = getattr(hn, "fuse_generator_grad_accumulation", False)
        ready_cb = getattr(hn, "grad_ready_callback", None)      # told about parameters whose .grad was updated in place (GradSync.notify)
        in_place = []
        if fused and all(t.grad is not None and t.grad.is_contiguous() and t.grad.dtype == torch.float32 for t in small.values()):
            # the C entry accumulates (+=): write straight into the existing .grad tensors (7 zero-fills and 7 add kernels fewer)
            g = {k: None for k in small}
            for k, t in small.items():
                setattr(a, k, t.grad.data_ptr())
                in_place.append(t)
        else:
            sizes = [t.numel() for t in small.values()]
            flat = torch.zeros(sum((n + 3) // 4 * 4 for n in sizes), dtype=torch.float32, device=dev)      # one memset for all seven
            g, off = {}, 0
            for (k, t), n in zip(small.items(), sizes):
                g[k] = flat[off:off + n].view_as(t)
                setattr(a, k, g[k].data_ptr())
                off += (n + 3) // 4 * 4
        scratch = torch.empty(int(lib.dmi_hypernet_scratch_floats(NQ, int(a.S_z), D)), dtype=torch.float32, device=dev)
        a.scratch = scratch.data_ptr()
        gen_grads: List[Optional[torch.Tensor]] = []
        hold = []
        sinks = getattr(hn, "factor_sinks", None) or {}          # {layer: parallel.Rank1FactorSync}: keep (dw, e) instead of a dense dG
        factor_layers = []
        a.overwrite_gen_grads = 0 if fused else 1
        for l in range(n_layers):
            dw = dws[l]
            gw, gb = gens[2 * l], gens[2 * l + 1]
            if dw is None:                         # layer without gradient (reference-as-written: generators.1)
                a.dw[l] = None
                gen_grads += [None, None]
                continue
            dw = dw.float().contiguous()
            hold.append(dw)
            a.dw[l] = dw.data_ptr()
            if l in sinks:
                # rank-1 factor mode: the kernel only reads G (de = G^T dw); (out_scale * dw, e_l) go to the sink and
                # the dense gradient is formed once per optimizer step from the factors of all micro-steps and ranks
                a.dgen_w[l], a.dgen_b[l] = None, None
                gen_grads += [None, None]
                factor_layers.append((l, dw))
                continue
            if fused and gw.grad is not None and gb.grad is not None:
                # accumulate the rank-1 gradient straight into the existing .grad (saves a 2x283 MB read-modify-write pass)
                a.dgen_w[l], a.dgen_b[l] = gw.grad.data_ptr(), gb.grad.data_ptr()
                gen_grads += [None, None]
                in_place += [gw, gb]
            else:
                dgw = torch.empty_like(gw) if not fused else torch.zeros_like(gw)
                dgb = torch.empty_like(gb) if not fused else torch.zeros_like(gb)
                a.dgen_w[l], a.dgen_b[l] = dgw.data_ptr(), dgb.data_ptr()
                gen_grads += [dgw, dgb]
        if all(d is None for d in dws):
            return (None,) * (10 + 2 * n_layers)
        _lib.check(lib.dmi_hypernet_bwd(C.byref(a), ops._stream()), "dmi_hypernet_bwd")
        if factor_layers:
            stash = ctx.keepalive[3]
            off = int(lib.dmi_hypernet_stash_code_offset(NQ, int(a.S_z), D))
            for l, dw in factor_layers:
                sinks[l].push(dw * float(a.out_scale), stash[off + l * D: off + (l + 1) * D])
        if ready_cb is not None:
            for t in in_place:
                ready_cb(t)
        return (None, None, None, g["dprefix"], g["dwq"], g["dbq"], g["dwk"], g["dbk"], g["dwv"], g["dbv"], *gen_grads)


class HyperNetwork(nn.Module):
    def __init__(self, hn_args: HypnetArgs, lm_emb_dim, mm_emb_dim, n_tokens, device):
        super().__init__()
        setup_args(self, prefix="hn_", args=hn_args)
        self.lm_emb_dim = lm_emb_dim
        self.mm_emb_dim = mm_emb_dim
        self.n_tokens = n_tokens
        self.device = device

        if self.arch == "attention":
            self.hypnet = MultiheadSelfAttention(d_model=self.hypnet_dim, nhead=self.n_heads)
        elif self.arch in ("transformer", "att_w_nonlinear"):
            raise NotImplementedError(f"hn_arch='{self.arch}' is not built for sm_100a: every reference config uses 'attention'")
        else:
            raise ValueError(f"Unknown hypernetwork architecture: {self.arch}")
        if self.n_heads != 1:
            raise NotImplementedError("the pooling kernel implements the 1-head attention every reference config uses")
        if self.n_proj_layers > MAX_GEN_LAYERS:
            raise NotImplementedError(f"more than {MAX_GEN_LAYERS} projector layers")

        self.a_dims, self.b_dims = [], []
        gens = []
        for layer_idx in range(self.n_proj_layers):
            in_width = self.hypnet_dim if layer_idx == 0 else self.lm_emb_dim
            a_dim, b_dim = in_width * self.rank, self.rank * self.lm_emb_dim
            out_dim = a_dim + b_dim + (self.lm_emb_dim if self.predict_bias else 0)
            gens.append(nn.Linear(self.hypnet_dim, out_dim))
            self.a_dims.append(a_dim)
            self.b_dims.append(b_dim)
        self.generators = nn.ModuleList(gens)
        self.prefix_tokens = nn.Parameter(torch.randn((self.n_proj_layers, self.hypnet_dim)))
        if self.use_pos_encs:
            # one position per support row (2 per shot) + the prefix tokens + the instruction-prefix embedding
            self.pos_encs = PositionalEncoding(self.hypnet_dim, max_len=2 * n_tokens + self.n_proj_layers + 1, device=self.device)
        self._init_weights()
        self.fuse_generator_grad_accumulation = False
        self.to(self.device)

    def _init_weights(self):
        nn.init.xavier_uniform_(self.prefix_tokens)
        for gen in self.generators:
            nn.init.xavier_uniform_(gen.weight)
            nn.init.zeros_(gen.bias)

    def forward(self, z, keep_mask: Optional[torch.Tensor] = None, n_layers: Optional[int] = None):
        """z: [n, hypnet_dim] support sequence -> (a_weights, b_weights, biases | None), flat tensors per projector layer.
        ``n_layers`` (default: all) limits the pass to the first generators -- ``HyperNetWrapper.forward`` asks for one when the
        projector applies ``lora_forward`` as written, which never reads the second adapter: 409 MB less to stream.

        A sequence shorter than the context (2*n_tokens + n_prefix + 1) is zero-padded and key-masked by the reference
        (hypernet.py:144-151); masked keys get weight exactly 0, so attending over the valid rows only is the same function.
        ``keep_mask`` [n_prefix, n_prefix + n] injects the attention-dropout mask (parity tests); in training mode without
        it a mask is drawn on the device (torch RNG streams are not bit-matched to the reference's)."""
        ops._need_cuda(z)
        self._check_z(z)
        n_pref = self.prefix_tokens.shape[0]
        p_drop = self.hypnet.dropout.p
        if keep_mask is None and self.training and p_drop > 0:
            keep_mask = torch.empty(n_pref, n_pref + z.shape[0], dtype=torch.float32, device=z.device).bernoulli_(1.0 - p_drop)
        if keep_mask is not None:
            if keep_mask.device != z.device or tuple(keep_mask.shape) != (n_pref, n_pref + z.shape[0]):
                raise ValueError(f"keep_mask must be a [{n_pref}, {n_pref + z.shape[0]}] tensor on {z.device}, got {tuple(keep_mask.shape)} on {keep_mask.device}")
        att = self.hypnet
        gens = []
        for gen in list(self.generators)[: (n_layers if n_layers is not None else len(self.generators))]:
            gens += [gen.weight, gen.bias]
        pieces = _HyperNetFn.apply(self, z, keep_mask, self.prefix_tokens, att.q.weight, att.q.bias, att.k.weight, att.k.bias,
                                   att.v.weight, att.v.bias, *gens)
        per = 3 if self.predict_bias else 2
        a_weights, b_weights = list(pieces[0::per]), list(pieces[1::per])
        biases = list(pieces[2::per]) if self.predict_bias else None
        if self.hypnet_dim > self.mm_emb_dim:
            a_weights[0] = a_weights[0][: self.mm_emb_dim * self.rank]        # encoder narrower than the hypernet: first mm_dim rows of A0
        return a_weights, b_weights, biases

    def _check_z(self, z) -> None:
        """the reference's torch.cat([prefix_tokens, z]) raises on a width mismatch; the C side only needs ldz >= D and would
        silently read the first hypnet_dim columns of a wider z"""
        if z.dim() != 2 or z.shape[1] != self.hypnet_dim:
            raise ValueError(f"support sequence must be [n, {self.hypnet_dim}] (hypnet_dim), got {tuple(z.shape)}")

    @torch.no_grad()
    def mean_adapter(self, zs, n_total: Optional[int] = None, group=None):
        """Element-wise mean of the adapters of N support sets, computed as ONE generator pass over the mean modality code
        (the generators are linear: mean_n(G e_n + c) = G mean_n(e_n) + c), i.e. the 692 MB of generator weights are streamed once
        instead of N times (reference: HyperNetWrapper.generate_projector_from_multiple_adapters, hypernet.py:234-266).
        Eval-mode semantics (no attention dropout).  Returns (a_weights, b_weights, biases | None) like ``forward``.
        Data parallel: pass this rank's shard of the support sets (``parallel.shard_support_sets``) and the global count
        ``n_total``; the partial mean codes are summed over the ranks before the single generator pass."""
        for z in zs:
            self._check_z(z)
        assert len(zs) > 0 or n_total is not None
        n_all = len(zs) if n_total is None else int(n_total)
        dev = self.prefix_tokens.device
        NQ, D = self.prefix_tokens.shape
        att = self.hypnet
        lib = _lib.load()
        e_mean = torch.zeros(NQ, D, dtype=torch.float32, device=dev)
        keep = []
        pe = self.pos_encs.pe[0] if self.use_pos_encs else None
        for z in zs:
            ops._need_cuda(z)
            z = z.detach().float().contiguous()
            a = _CArgs()
            a.S_z, a.NQ, a.D, a.n_layers = z.shape[0], NQ, D, 0
            a.out_scale = float(self.alpha) / float(self.rank)
            a.z, a.ldz = z.data_ptr(), z.stride(0)
            a.prefix_tokens = self.prefix_tokens.data_ptr()
            if pe is not None:
                assert pe.shape[0] >= NQ + z.shape[0], "support set longer than the positional-encoding table (n_tokens too small)"
                a.pe, a.ldpe = pe.data_ptr(), pe.stride(0)
            for name, t in (("wq", att.q.weight), ("bq", att.q.bias), ("wk", att.k.weight), ("bk", att.k.bias), ("wv", att.v.weight), ("bv", att.v.bias)):
                setattr(a, name, t.data_ptr())
            stash = torch.empty(int(lib.dmi_hypernet_stash_floats(NQ, z.shape[0], D)), dtype=torch.float32, device=dev)
            a.stash = stash.data_ptr()
            keep.append((z, stash))
            _lib.check(lib.dmi_hypernet_pool(C.byref(a), C.c_void_p(e_mean.data_ptr()), 1.0 / n_all, ops._stream()), "dmi_hypernet_pool")
        if n_total is not None:
            from ..parallel import allreduce_sum_
            allreduce_sum_(e_mean, group)
        g = _CArgs()
        g.NQ, g.D, g.n_layers = NQ, D, len(self.generators)
        g.out_scale = float(self.alpha) / float(self.rank)
        outs = []
        for l, gen in enumerate(self.generators):
            g.gen_w[l], g.gen_b[l], g.gen_out[l] = gen.weight.data_ptr(), gen.bias.data_ptr(), gen.weight.shape[0]
            o = torch.empty(gen.weight.shape[0], dtype=torch.float32, device=dev)
            g.w_out[l] = o.data_ptr()
            outs.append(o)
        _lib.check(lib.dmi_hypernet_generate(C.byref(g), C.c_void_p(e_mean.data_ptr()), ops._stream()), "dmi_hypernet_generate")
        return self._split_generated(outs)

    def _split_generated(self, outs):
        a_weights, b_weights = [], []
        biases = [] if self.predict_bias else None
        for idx, w in enumerate(outs):
            a_dim, b_dim = self.a_dims[idx], self.b_dims[idx]
            a_w = w[:a_dim]
            if idx == 0 and self.hypnet_dim > self.mm_emb_dim:
                a_w = a_w[: self.mm_emb_dim * self.rank]        # encoder narrower than the hypernet: first mm_dim rows of A0
            a_weights.append(a_w)
            b_weights.append(w[a_dim:a_dim + b_dim])
            if self.predict_bias:
                biases.append(w[a_dim + b_dim:])
        return a_weights, b_weights, biases


class HyperNetWrapper(nn.Module):
    def __init__(self, hn_args, proj_args, lm_emb_dim, mm_emb_dim, n_tokens, device):
        super().__init__()
        self.device = device
        self.hn_args = hn_args
        self.proj_args = proj_args
        self.hypernet = HyperNetwork(hn_args, lm_emb_dim, mm_emb_dim, n_tokens, device)
        self.projector = Projector(proj_args, lm_emb_dim, mm_emb_dim, device)
        self.projector.load_model()
        self.generated_projector = None

    # -- SURVEY H6 ------------------------------------------------------------------------------------------------------------
    @property
    def clip_includes_frozen_projector(self) -> bool:
        """Compatibility switch for an exact reference train step.  The reference clips ``clip_grad_norm_(self.model.hypernet
        .parameters())`` over this WHOLE wrapper (train_hypernet.py:148), i.e. including the frozen-by-omission projector whose
        ``net.0.{weight,bias}.grad`` autograd keeps filling and nothing ever zeroes (the optimizer only owns the hypernet,
        train_hypernet.py:526-532): the clip norm grows step after step.  False (default): the projector receives no gradients and
        ``clip_parameters()`` is the hypernet alone.  True: the kernels also accumulate the useless dW1 / db1 and
        ``clip_parameters()`` returns every parameter of the wrapper, reproducing the reference's clip factor."""
        return self.projector.accumulate_frozen_base_grads

    @clip_includes_frozen_projector.setter
    def clip_includes_frozen_projector(self, on: bool) -> None:
        self.projector.accumulate_frozen_base_grads = bool(on)

    def clip_parameters(self):
        """the parameter set the trainer's clip_grad_norm_ runs over (see ``clip_includes_frozen_projector``)"""
        if self.generated_projector is not None:
            return self.generated_projector.parameters()
        return self.parameters() if self.clip_includes_frozen_projector else self.hypernet.parameters()

    def train(self, mode=True):
        if not isinstance(mode, bool):
            raise ValueError("training mode is expected to be boolean")
        self.training = mode
        for child in self.children():
            child.train(mode)
        self.projector.eval()          # the shared projector stays in eval (hypernet.py:219-227)
        return self

    def generate_projector(self, z):
        with torch.no_grad():
            a_w, b_w, biases = self.hypernet(z)
            self.generated_projector = self.projector.combine_lora(a_w, b_w, biases)

    def generate_projector_from_multiple_adapters(self, zs):
        """N support sets -> N adapters -> element-wise mean -> merged projector (hypernet.py:234-266).  The mean adapter is
        produced by one generator pass over the mean modality code (``HyperNetwork.mean_adapter``); in training mode with
        attention dropout the reference's N independent forward passes are kept."""
        with torch.no_grad():
            if not (self.hypernet.training and self.hypernet.hypnet.dropout.p > 0):
                a_w, b_w, biases = self.hypernet.mean_adapter(zs)
                self.generated_projector = self.projector.combine_lora(a_w, b_w, biases)
                return
            n = len(zs)
            acc_a = acc_b = acc_bias = None
            for z in zs:
                a_w, b_w, biases = self.hypernet(z)
                if acc_a is None:
                    acc_a = [t.clone() for t in a_w]
                    acc_b = [t.clone() for t in b_w]
                    acc_bias = None if biases is None else [t.clone() for t in biases]
                else:
                    for dst, src in zip(acc_a, a_w):
                        dst += src
                    for dst, src in zip(acc_b, b_w):
                        dst += src
                    if biases is not None:
                        for dst, src in zip(acc_bias, biases):
                            dst += src
            avg = lambda ts: None if ts is None else [t / n for t in ts]
            self.generated_projector = self.projector.combine_lora(avg(acc_a), avg(acc_b), avg(acc_bias))

    def forward(self, x, z):
        if self.generated_projector is not None:
            return self.generated_projector(x)
        if self.projector.lora_forward_mode == "as_written" and self.hypernet.n_proj_layers == 2:
            # lora_forward as written stops after the first GELU (projector.py:124): the second adapter is never read and
            # generators.1 never receives a gradient, so its 409 MB GEMV is not run at all (same outputs, same gradients, grad None)
            a_w, b_w, biases = self.hypernet(z, n_layers=1)
            return self.projector.lora_forward_first_layer(x, a_w[0], b_w[0], None if biases is None else biases[0])
        a_w, b_w, biases = self.hypernet(z)
        return self.projector.lora_forward(x, a_w, b_w, biases)

    def trainable_parameters(self):
        if self.generated_projector is not None:
            return self.generated_projector.parameters()
        return self.hypernet.parameters()
