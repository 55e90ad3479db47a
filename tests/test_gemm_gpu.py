"""Parity of the tcgen05 GEMM building block and the mma.sync batch-reduction against fp32 torch matmul on the same
bf16-rounded inputs (tolerance: bf16 path 2e-3 relative to the fp32 result; outputs stored in bf16 get 1e-2)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def rows_rel(a, b):
    """worst row: a handful of wrong rows in a ragged last tile must not hide inside a Frobenius norm"""
    d = (a.double() - b.double()).norm(dim=1)
    return (d / b.double().norm(dim=1).clamp_min(1e-6)).max().item()


def gelu_tanh(x):
    return torch.nn.functional.gelu(x, approximate="tanh")


@pytest.fixture(scope="module")
def ops():
    from dmi_b200 import ops
    return ops


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 128), (256, 512, 800), (4, 2048, 800), (300, 2048, 2080),
                                   (1000, 32, 768), (257, 8, 2048), (129, 64, 96), (2048, 2048, 832), (5000, 2048, 800)])
def test_gemm_store_f32(ops, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    b = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.full((M, N), float("nan"), device="cuda")
    ops.gemm_tn(a, b, bias=bias, out0=out)
    ref = a.float() @ b.float().T + bias
    assert torch.isfinite(out).all()
    assert rel(out, ref) < 2e-3, rel(out, ref)


def test_gemm_strided_operands_and_bf16_out(ops):
    # A is a column slice of a wider buffer (ld > K), output goes into a column slice: the [x | u] packing pattern
    M, D, r = 700, 768, 32
    g = torch.Generator(device="cuda").manual_seed(1)
    xext = torch.randn(M, D + r, device="cuda", generator=g).to(torch.bfloat16)
    keep = xext.clone()
    a0t = (torch.randn(r, D, device="cuda", generator=g) / math.sqrt(D)).to(torch.bfloat16)
    ops.gemm_tn(xext[:, :D], a0t, out0=xext[:, D:])
    ref = keep[:, :D].float() @ a0t.float().T
    assert torch.equal(xext[:, :D], keep[:, :D])           # untouched
    assert rel(xext[:, D:].float(), ref) < 1e-2


@pytest.mark.parametrize("M,N,K", [(300, 2048, 800), (4, 256, 64)])
def test_gemm_gelu_epilogue(ops, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(2)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    b = (torch.randn(N, K, device="cuda", generator=g) * (2.0 / math.sqrt(K))).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g) * 0.5
    h = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    pre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm_tn(a, b, mode=ops.EPI_GELU, bias=bias, out0=h, out1=pre)
    pre_ref = a.float() @ b.float().T + bias
    assert rel(pre.float(), pre_ref) < 1e-2
    assert rel(h.float(), gelu_tanh(pre_ref)) < 1e-2


def test_gemm_gelu_bwd_epilogue(ops):
    M, N, K = 300, 2048, 2080
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    b = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(torch.bfloat16)
    pre = (torch.randn(M, N, device="cuda", generator=g) * 1.5).to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm_tn(a, b, mode=ops.EPI_GELU_BWD, out0=out, aux=pre)
    p = pre.float().requires_grad_(True)
    gelu_tanh(p).sum().backward()
    ref = (a.float() @ b.float().T) * p.grad
    assert rel(out.float(), ref) < 1e-2


@pytest.mark.parametrize("M,N,K", [(260, 768, 768), (132, 512, 512), (4, 768, 768)])
def test_gemm_tf32(ops, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(4)
    a = torch.randn(M, K, device="cuda", generator=g)
    b = torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)
    out = torch.empty(M, N, device="cuda")
    ops.gemm_tn(a, b, out0=out)
    ref = (a.double() @ b.double().T).float()
    assert rel(out, ref) < 1e-3, rel(out, ref)


@pytest.mark.parametrize("B,P,Q,tr,cs", [(1000, 32, 2048, False, True), (64, 32, 768, True, False), (5, 8, 256, False, True),
                                         (4097, 64, 2048, True, False), (333, 16, 520, False, True)])
def test_outer_reduce(ops, B, P, Q, tr, cs):
    g = torch.Generator(device="cuda").manual_seed(B + P + Q)
    L = torch.randn(B, P + 8, device="cuda", generator=g).to(torch.bfloat16)[:, :P]      # strided view
    R = torch.randn(B, Q, device="cuda", generator=g).to(torch.bfloat16)
    G = torch.ones((Q, P) if tr else (P, Q), device="cuda")
    colsum = torch.ones(Q, device="cuda") if cs else None
    ops.outer_reduce(L, R, G, transpose_out=tr, colsum=colsum, scale=0.5)
    ref = 0.5 * (L.float().T @ R.float())
    ref = (ref.T if tr else ref) + 1.0
    assert rel(G, ref) < 1e-4, rel(G, ref)
    if cs:
        assert rel(colsum, 0.5 * R.float().sum(0) + 1.0) < 1e-4


@pytest.mark.parametrize("M,N,K,mode", [(256, 256, 64, 0), (300, 2048, 800, 1), (1000, 512, 2080, 0), (2500, 2048, 2080, 2), (20000, 2048, 832, 1),
                                        (130, 256, 128, 0), (16384, 2048, 2080, 0), (1, 128, 64, 0), (33, 2048, 800, 2), (8191, 2048, 800, 3),
                                        (20001, 1024, 2080, 3), (257, 128, 800, 1)])
def test_gemm_cta_pair_variant(ops, M, N, K, mode):
    """tcgen05.mma.cta_group::2 (256x256 tile per CTA pair, TMA-store epilogue) against fp32 matmul; odd M-tile counts leave a phantom
    half-tile, ragged M is clipped by the store's tensor map.  Frobenius AND worst-row error; every output is pre-filled with NaN and
    lives in a wider buffer (row stride > N) whose remaining columns must stay untouched."""
    ops.set_option("gemm_pair", 1)
    try:
        g = torch.Generator(device="cuda").manual_seed(M + N + K + mode)
        a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
        b = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(torch.bfloat16)
        bias = torch.randn(N, device="cuda", generator=g) * 0.3
        acc = a.float() @ b.float().T
        bf = torch.bfloat16
        wide = lambda dt: torch.full((M, N + 32), float("nan"), device="cuda", dtype=dt)
        if mode == 0:
            buf = wide(torch.float32)
            ops.gemm_tn(a, b, bias=bias, out0=buf[:, :N])
            assert rel(buf[:, :N], acc + bias) < 2e-3 and rows_rel(buf[:, :N], acc + bias) < 5e-3
            assert torch.isnan(buf[:, N:]).all()
        elif mode == 3:
            buf = wide(bf)
            ops.gemm_tn(a, b, bias=bias, out0=buf[:, :N])
            assert rel(buf[:, :N].float(), acc + bias) < 1e-2 and rows_rel(buf[:, :N].float(), acc + bias) < 1e-2
            assert torch.isnan(buf[:, N:]).all()
        elif mode == 1:
            hb, pb = wide(bf), wide(bf)
            ops.gemm_tn(a, b, mode=ops.EPI_GELU, bias=bias, out0=hb[:, :N], out1=pb[:, :N])
            assert rel(pb[:, :N].float(), acc + bias) < 1e-2 and rel(hb[:, :N].float(), gelu_tanh(acc + bias)) < 1e-2
            assert rows_rel(pb[:, :N].float(), acc + bias) < 1e-2 and rows_rel(hb[:, :N].float(), gelu_tanh(acc + bias)) < 1.5e-2
            assert torch.isnan(hb[:, N:]).all() and torch.isnan(pb[:, N:]).all()
        else:
            pre = torch.randn(M, N + 8, device="cuda", generator=g).to(bf)[:, :N]
            buf = wide(bf)
            ops.gemm_tn(a, b, mode=ops.EPI_GELU_BWD, out0=buf[:, :N], aux=pre)
            pp = pre.float().requires_grad_(True)
            gelu_tanh(pp).sum().backward()
            assert rel(buf[:, :N].float(), acc * pp.grad) < 1e-2 and rows_rel(buf[:, :N].float(), acc * pp.grad) < 1.5e-2
            assert torch.isnan(buf[:, N:]).all()
    finally:
        ops.set_option("gemm_pair", -1)


@pytest.mark.parametrize("M,K,R", [(1000, 768, 32), (64, 128, 8), (5, 2048, 64), (16384, 2048, 32), (777, 520, 16), (16, 800, 32), (1, 768, 32), (17, 2048, 32)])
@pytest.mark.parametrize("f32", [False, True])
def test_skinny_rows(ops, M, K, R, f32):
    """row-panel rank-r projection; the fp32 variant also emits the bf16 copy of its input (bit-exact round-to-nearest)"""
    g = torch.Generator(device="cuda").manual_seed(M + K + R)
    x = torch.randn(M, K, device="cuda", generator=g)
    W = (torch.randn(R, K, device="cuda", generator=g) / math.sqrt(K)).to(torch.bfloat16)
    buf = torch.full((M, K + R), 7.0, device="cuda", dtype=torch.bfloat16)          # [copy | out] like xext
    if f32:
        ops.skinny_rows(x, W, buf[:, K:], copy=buf[:, :K])
        assert torch.equal(buf[:, :K], x.to(torch.bfloat16))
    else:
        buf[:, :K] = x.to(torch.bfloat16)
        ops.skinny_rows(buf[:, :K], W, buf[:, K:])
    ref = x.to(torch.bfloat16).float() @ W.float().T
    assert rel(buf[:, K:].float(), ref) < 1e-2
