"""torch-facing wrappers over the C ABI (device pointers + current CUDA stream).  No fallbacks: CPU tensors raise."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import MlpArgs

KIND_BF16, KIND_TF32 = 0, 1
EPI_STORE, EPI_GELU, EPI_GELU_BWD = 0, 1, 2


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("dmi_b200 runs on CUDA tensors only (there is no CPU fallback)")


class DeferredErrorFlag:
    """Device-side error flag with a host check that costs no synchronisation.

    Kernels that validate indices on the device (prefix splice: token ids, embedding store: sample rows) raise ``flag`` instead of
    trapping.  The reference raises on the spot (``nn.Embedding`` device assert / host ``IndexError``); a silent flag would let training
    continue on garbage, so after every launch the flag is copied to pinned host memory asynchronously (``arm``) and the NEXT call --
    or ``check()``, which blocks -- inspects the copy once its event has completed and raises ``IndexError``."""

    def __init__(self, device, what: str):
        self.what = what
        self.flag = torch.zeros(1, dtype=torch.int32, device=device)
        self._host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._event = torch.cuda.Event()
        self._armed = False

    def arm(self) -> None:
        if torch.cuda.is_current_stream_capturing():
            return                               # inside a CUDA graph the flag still accumulates on the device; check() after replays
        self._host.copy_(self.flag, non_blocking=True)
        self._event.record()
        self._armed = True

    def poll(self, block: bool = False) -> None:
        if torch.cuda.is_current_stream_capturing():
            return
        if not self._armed and not block:
            return
        if block:
            self._host.copy_(self.flag)           # synchronous copy on the current stream
        elif not self._event.query():
            return
        self._armed = False
        if int(self._host[0]) != 0:
            self.flag.zero_()
            self._host.zero_()
            raise IndexError(f"{self.what}: index out of range (detected on the device by an earlier launch)")

    def check(self) -> None:
        """blocking check, e.g. at the end of a step or after replaying a captured graph"""
        self.poll(block=True)


_SPLICE_FLAGS = {}


def splice_error_flag(device) -> DeferredErrorFlag:
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _SPLICE_FLAGS:
        _SPLICE_FLAGS[key] = DeferredErrorFlag(torch.device("cuda", key), "prefix splice: token id outside the embedding table")
    return _SPLICE_FLAGS[key]


def check_device_errors() -> None:
    """blocking check of every deferred index-error flag of the splice path (call where the reference would have synchronised anyway,
    e.g. next to ``loss.item()``)"""
    for f in _SPLICE_FLAGS.values():
        f.check()


def _rows(t: torch.Tensor) -> int:
    """leading dimension (elements) of a 2-D row-major view"""
    assert t.dim() == 2 and (t.stride(1) == 1 or t.shape[1] == 1), "need unit stride along the last dim"
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def gemm_tn(a: torch.Tensor, b: torch.Tensor, *, mode: int = EPI_STORE, bias: Optional[torch.Tensor] = None,
            out0: torch.Tensor, out1: Optional[torch.Tensor] = None, aux: Optional[torch.Tensor] = None,
            alpha: float = 1.0) -> torch.Tensor:
    """out0[M,N] = epilogue(alpha * a[M,K] @ b[N,K]^T + bias).  a, b both bf16 (tcgen05 kind::f16) or both fp32 (kind::tf32)."""
    _need_cuda(a, b, out0, out1, aux, bias)
    assert a.dtype == b.dtype and a.dtype in (torch.bfloat16, torch.float32)
    kind = KIND_BF16 if a.dtype == torch.bfloat16 else KIND_TF32
    M, K = a.shape
    N, K2 = b.shape
    assert K == K2 and out0.shape == (M, N)
    assert out0.dtype in (torch.float32, torch.bfloat16)
    rc = _lib.load().dmi_gemm_tn(kind, mode, _ptr(a), _rows(a), _ptr(b), _rows(b), M, N, K, alpha, _ptr(bias),
                                 _ptr(out0), _rows(out0), int(out0.dtype == torch.float32),
                                 _ptr(out1), 0 if out1 is None else _rows(out1),
                                 _ptr(aux), 0 if aux is None else _rows(aux), _stream())
    _lib.check(rc, "dmi_gemm_tn")
    return out0


def outer_reduce(L: torch.Tensor, R: torch.Tensor, G: torch.Tensor, *, transpose_out: bool = False,
                 colsum: Optional[torch.Tensor] = None, scale: float = 1.0) -> torch.Tensor:
    """G += scale * L^T R   (L [B,P], R [B,Q] bf16; G fp32 [P,Q] or [Q,P] when transpose_out)."""
    _need_cuda(L, R, G, colsum)
    B, P = L.shape
    B2, Q = R.shape
    assert B == B2 and L.dtype == torch.bfloat16 and R.dtype == torch.bfloat16 and G.dtype == torch.float32
    assert G.shape == ((Q, P) if transpose_out else (P, Q))
    rc = _lib.load().dmi_outer_reduce(_ptr(L), _rows(L), _ptr(R), _rows(R), B, P, Q, _ptr(G), _rows(G), int(transpose_out),
                                      _ptr(colsum), scale, _stream())
    _lib.check(rc, "dmi_outer_reduce")
    return G


class PackedProjector:
    """bf16 operand cache of a frozen MLP2 projector plus one adapter, in the layouts the kernels consume.

    ``w1ext = [W1 | B0^T]``, ``w2ext = [W2 | B1^T]``, ``w2text = [W2^T | A1]`` so that the low-rank term is r extra K columns
    of each GEMM.  fp32 master weights stay in the nn.Module; these are derived caches and are never saved.

    Single-stream contract: the adapter columns of these buffers are REWRITTEN by every ``pack_adapter`` (one adapter at a time), so all
    calls that use one PackedProjector -- pack, forward, backward -- must be enqueued on the same CUDA stream (or be ordered by events).
    Two forwards in flight on different streams or host threads race on the operands.  ``Projector`` re-packs in backward when another
    adapter was packed in between (``adapter_epoch``), which covers interleaved forward/backward on ONE stream only; capture a graph
    of the step (``graphs.GraphedStep``) on the stream that also warmed it up, and replay instead of mixing eager calls on another stream.
    """

    def __init__(self, D: int, H: int, r: int, device):
        self.D, self.H, self.r = D, H, r
        bf = dict(dtype=torch.bfloat16, device=device)
        self.w1ext = torch.zeros(H, D + r, **bf)
        self.w2ext = torch.zeros(H, H + r, **bf)
        self.w2text = torch.zeros(H, H + r, **bf)
        self.a0t = torch.zeros(max(r, 1), D, **bf)
        self.a1t = torch.zeros(max(r, 1), H, **bf)
        self.b0 = torch.zeros(max(r, 1), H, **bf)
        self.b1 = torch.zeros(max(r, 1), H, **bf)
        self.bias0 = torch.zeros(H, dtype=torch.float32, device=device)
        self.bias1 = torch.zeros(H, dtype=torch.float32, device=device)
        self.base_version = None

    def pack_base(self, W1: torch.Tensor, W2: Optional[torch.Tensor]):
        """W1 [H, >=D] (column-pruned view allowed), W2 [H,H]; fp32."""
        _need_cuda(W1, W2)
        assert W1.dtype == torch.float32 and W1.shape[0] == self.H and W1.shape[1] >= self.D and W1.stride(1) == 1
        rc = _lib.load().dmi_projector_pack_base(_ptr(W1), W1.stride(0), _ptr(W2), self.D, self.H, self.r,
                                                 _ptr(self.w1ext), _ptr(self.w2ext), _ptr(self.w2text), _stream())
        _lib.check(rc, "dmi_projector_pack_base")

    def pack_adapter(self, A0, B0, beta0, A1, B1, beta1, b1, b2, scale: float = 1.0):
        """flat or shaped fp32 adapter tensors (A0 [D*r], B0 [r*H], beta0 [H] | None, ...); b1/b2 = base biases."""
        ts = [t if t is None else t.detach().contiguous().float() for t in (A0, B0, beta0, A1, B1, beta1, b1, b2)]
        _need_cuda(*ts)
        rc = _lib.load().dmi_adapter_pack(*[_ptr(t) for t in ts], self.D, self.H, self.r, scale,
                                          _ptr(self.w1ext), _ptr(self.w2ext), _ptr(self.w2text), _ptr(self.a0t), _ptr(self.a1t),
                                          _ptr(self.b0), _ptr(self.b1), _ptr(self.bias0), _ptr(self.bias1), _stream())
        _lib.check(rc, "dmi_adapter_pack")


class MlpStash:
    """activation buffers of one adapted-MLP step (bf16), reusable across steps of the same shape"""

    def __init__(self, B: int, D: int, H: int, r: int, device, full: bool = True, xext: Optional[torch.Tensor] = None):
        bf = dict(dtype=torch.bfloat16, device=device)
        self.B = B
        if xext is not None:       # caller-owned operand buffer whose columns [0,D) already hold bf16 x (e.g. the target of an H2D copy)
            assert xext.dtype == torch.bfloat16 and xext.shape == (B, D + r) and xext.is_contiguous()
        self.xext = xext if xext is not None else torch.empty(B, D + r, **bf)
        self.pre = torch.empty(B, H, **bf)
        self.dpre = torch.empty(B, H, **bf)
        self.du = torch.empty(B, max(r, 8), **bf)
        self.hext = torch.empty(B, H + r, **bf) if full else None
        self.dyext = torch.empty(B, H + r, **bf) if full else None


def _fill_args(pk: PackedProjector, st: MlpStash, B: int, flags: int) -> MlpArgs:
    a = MlpArgs()
    a.B, a.D, a.H, a.r = B, pk.D, pk.H, pk.r
    a.flags = flags
    a.grad_scale = 1.0
    a.w1ext, a.w2ext, a.w2text = pk.w1ext.data_ptr(), pk.w2ext.data_ptr(), pk.w2text.data_ptr()
    a.a0t, a.a1t, a.b0, a.b1 = pk.a0t.data_ptr(), pk.a1t.data_ptr(), pk.b0.data_ptr(), pk.b1.data_ptr()
    a.bias0, a.bias1 = pk.bias0.data_ptr(), pk.bias1.data_ptr()
    a.xext, a.pre, a.dpre, a.du = st.xext.data_ptr(), st.pre.data_ptr(), st.dpre.data_ptr(), st.du.data_ptr()
    if st.hext is not None:
        a.hext, a.dyext = st.hext.data_ptr(), st.dyext.data_ptr()
    return a


def adapted_mlp_fwd(pk: PackedProjector, st: MlpStash, x: Optional[torch.Tensor], y: Optional[torch.Tensor], *, flags: int = 0,
                    y_bf16: Optional[torch.Tensor] = None, keep: Optional[torch.Tensor] = None, dropout_p: float = 0.0) -> None:
    _need_cuda(x, y, y_bf16, keep)
    B = st.B if x is None else x.shape[0]
    a = _fill_args(pk, st, B, flags)
    if keep is not None:
        assert keep.dtype == torch.uint8 and keep.shape == (B, pk.H) and keep.is_contiguous()
        a.keep, a.dropout_p = keep.data_ptr(), dropout_p
    if x is not None:
        assert x.dtype == torch.float32 and x.shape[1] == pk.D
        a.x, a.ldx = x.data_ptr(), _rows(x)
    if y is not None:
        a.y, a.ldy = y.data_ptr(), _rows(y)
    if y_bf16 is not None:
        a.y_bf16, a.ldy_bf16 = y_bf16.data_ptr(), _rows(y_bf16)
    _lib.check(_lib.load().dmi_adapted_mlp_fwd(C.byref(a), _stream()), "dmi_adapted_mlp_fwd")


def adapted_mlp_bwd(pk: PackedProjector, st: MlpStash, dy: torch.Tensor, grads: dict, *, flags: int = 0, grad_scale: float = 1.0,
                    keep: Optional[torch.Tensor] = None, dropout_p: float = 0.0, layer1_event: Optional[torch.cuda.Event] = None) -> None:
    """grads: dict with fp32 tensors dA0 [D,r], dB0 [r,H], dbeta0 [H] (and dA1 [H,r], dB1, dbeta1 in full mode), or
    dW1 [H,D], db1, dW2 [H,H], db2 with MLP_BASE_GRADS; accumulated into."""
    _need_cuda(dy, keep, *grads.values())
    assert dy.dtype == torch.float32
    a = _fill_args(pk, st, dy.shape[0], flags)
    if keep is not None:
        a.keep, a.dropout_p = keep.data_ptr(), dropout_p
    if layer1_event is not None:
        layer1_event.record()          # materialises the lazily-created cudaEvent_t; re-recorded by the library mid-backward
        a.ev_layer1_grads = layer1_event.cuda_event
    a.grad_scale = grad_scale
    a.dy, a.lddy = dy.data_ptr(), _rows(dy)
    for k, t in grads.items():
        assert t.dtype == torch.float32 and t.is_contiguous()
        setattr(a, k, t.data_ptr())
    _lib.check(_lib.load().dmi_adapted_mlp_bwd(C.byref(a), _stream()), "dmi_adapted_mlp_bwd")


def gemm_mn(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, *, alpha: float = 1.0, accumulate: bool = False) -> torch.Tensor:
    """out[M,N] (+)= alpha * a[K,M]^T @ b[K,N]  (bf16 operands contracted over their rows, fp32 out)."""
    _need_cuda(a, b, out)
    K, M = a.shape
    K2, N = b.shape
    assert K == K2 and out.shape == (M, N) and out.dtype == torch.float32 and a.dtype == b.dtype == torch.bfloat16
    rc = _lib.load().dmi_gemm_mn(_ptr(a), _rows(a), _ptr(b), _rows(b), M, N, K, alpha, _ptr(out), _rows(out), int(accumulate), _stream())
    _lib.check(rc, "dmi_gemm_mn")
    return out


def merge_adapter(W: torch.Tensor, bias: torch.Tensor, A: torch.Tensor, B: torch.Tensor, beta: Optional[torch.Tensor],
                  scale: float = 1.0):
    """exact fp32 W' = W + scale*(A B)^T, b' = bias + beta   (Projector.combine_lora, projector.py:95-103)"""
    _need_cuda(W, bias, A, B, beta)
    H, in_dim = W.shape
    A = A.detach().reshape(in_dim, -1).contiguous().float()
    r = A.shape[1]
    B = B.detach().reshape(r, H).contiguous().float()
    Wm = torch.empty(H, in_dim, dtype=torch.float32, device=W.device)
    bm = torch.empty(H, dtype=torch.float32, device=W.device)
    Wd = W.detach()
    rc = _lib.load().dmi_merge_adapter(_ptr(Wd), Wd.stride(0), _ptr(bias.detach()), _ptr(A), _ptr(B),
                                       _ptr(None if beta is None else beta.detach().contiguous().float()),
                                       in_dim, H, r, scale, _ptr(Wm), in_dim, _ptr(bm), _stream())
    _lib.check(rc, "dmi_merge_adapter")
    return Wm, bm


def launch_count() -> int:
    return int(_lib.load().dmi_launch_count())


def set_option(name: str, value: int) -> None:
    """tuning switches of the library (A/B measurements and tests), e.g. ``set_option("gemm_pair", 0)``"""
    _lib.check(_lib.load().dmi_set_option(name.encode(), int(value)), "dmi_set_option")


def skinny_rows(inp: torch.Tensor, W: torch.Tensor, out: torch.Tensor, copy: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[M,R] = inp[M,K] @ W[R,K]^T (bf16 out).  fp32 ``inp`` is converted on the fly and its bf16 copy stored in ``copy``."""
    _need_cuda(inp, W, out, copy)
    M, K = inp.shape
    R = W.shape[0]
    assert W.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and out.shape == (M, R) and inp.dtype in (torch.float32, torch.bfloat16)
    rc = _lib.load().dmi_skinny_rows(_ptr(inp), _rows(inp), int(inp.dtype == torch.float32), _ptr(W), _rows(W), _ptr(out), _rows(out),
                                     _ptr(copy), 0 if copy is None else _rows(copy), M, K, R, _stream())
    _lib.check(rc, "dmi_skinny_rows")
    return out


def panel_fused_tc(inp: torch.Tensor, W: torch.Tensor, L: torch.Tensor, out: torch.Tensor, G: torch.Tensor, *,
                   colsum: Optional[torch.Tensor] = None, scale: float = 1.0) -> torch.Tensor:
    """One tcgen05 sweep over bf16 ``inp`` [M,K]: out = inp @ W^T (bf16), G += scale * L^T inp, colsum += scale * 1^T inp.

    The fused form of :func:`skinny_rows` + :func:`outer_reduce` over the same matrix (rank 32, K in {1024, 2048})."""
    _need_cuda(inp, W, L, out, G, colsum)
    M, K = inp.shape
    R = W.shape[0]
    assert inp.dtype == W.dtype == L.dtype == out.dtype == torch.bfloat16 and G.dtype == torch.float32
    assert L.shape == (M, R) and out.shape == (M, R) and G.shape == (R, K)
    rc = _lib.load().dmi_panel_fused_tc(_ptr(inp), _rows(inp), _ptr(W), _rows(W), _ptr(out), _rows(out), _ptr(L), _rows(L), _ptr(G), _rows(G),
                                        _ptr(colsum), scale, M, K, R, _stream())
    _lib.check(rc, "dmi_panel_fused_tc")
    return out


def panel_tc_project(inp: torch.Tensor, W: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out[M,R] = inp @ W^T with the tcgen05 panel kernel (bf16, rank 32, K in {768, 1024, 2048}); same contract as skinny_rows."""
    _need_cuda(inp, W, out)
    M, K = inp.shape
    R = W.shape[0]
    assert inp.dtype == W.dtype == out.dtype == torch.bfloat16 and out.shape == (M, R)
    rc = _lib.load().dmi_panel_tc_project(_ptr(inp), _rows(inp), _ptr(W), _rows(W), _ptr(out), _rows(out), M, K, R, _stream())
    _lib.check(rc, "dmi_panel_tc_project")
    return out


def panel_fused_tc32(inp: torch.Tensor, W: torch.Tensor, L: torch.Tensor, out: torch.Tensor, G: torch.Tensor, *,
                     colsum: Optional[torch.Tensor] = None, copy: Optional[torch.Tensor] = None, scale: float = 1.0) -> torch.Tensor:
    """fp32-input form of :func:`panel_fused_tc` (the dY pass; also writes the bf16 ``copy``)."""
    _need_cuda(inp, W, L, out, G, colsum, copy)
    M, K = inp.shape
    R = W.shape[0]
    assert inp.dtype == torch.float32 and W.dtype == L.dtype == out.dtype == torch.bfloat16 and G.dtype == torch.float32
    assert L.shape == (M, R) and out.shape == (M, R) and G.shape == (R, K) and (copy is None or copy.dtype == torch.bfloat16)
    rc = _lib.load().dmi_panel_fused_tc32(_ptr(inp), _rows(inp), _ptr(W), _rows(W), _ptr(out), _rows(out), _ptr(copy),
                                          0 if copy is None else _rows(copy), _ptr(L), _rows(L), _ptr(G), _rows(G), _ptr(colsum), scale,
                                          M, K, R, _stream())
    _lib.check(rc, "dmi_panel_fused_tc32")
    return out


