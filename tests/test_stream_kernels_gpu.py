"""Register-streaming side kernels (dmi_stream_project / dmi_stream_reduce / dmi_lq_pack) and the merged-weight, overlapped
schedule of the adapted MLP (DMI_MLP_MERGED, side products from the GEMM's staged tiles) against torch references / the CPU oracle.

Index work (the pair-interleaved LQ layout, the bf16 copy) is bit-exact; products are bf16 x bf16 with fp32 accumulation."""
import math

import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-2


def rel(a, b):
    return ((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def ops():
    from dmi_b200 import ops as _ops
    return _ops


def lq_reference(X):
    """plain [B,P] bf16 -> the LQ word layout, computed on the host: u32 LQ[b/2][g][jh] = {X[2p][8jh+g], X[2p+1][8jh+g]}"""
    X = X.cpu()
    B, P = X.shape
    PJ = max(P, 16)
    JH = PJ // 8
    bits = X.view(torch.int16).to(torch.int32) & 0xFFFF
    if B % 2:
        bits = torch.cat([bits, torch.zeros(1, P, dtype=torch.int32)])
    pairs = bits[0::2] | (bits[1::2] << 16)                      # [ceil(B/2), P]
    out = torch.zeros(pairs.shape[0], 8, JH, dtype=torch.int32)
    for j in range(P):
        out[:, j % 8, j // 8] = pairs[:, j]
    return out.reshape(-1)


@pytest.mark.parametrize("B,P", [(64, 32), (7, 8), (1001, 16), (300, 64)])
def test_lq_pack_layout_bit_exact(ops, B, P):
    g = torch.Generator(device="cuda").manual_seed(B + P)
    X = torch.randn(B, P + 8, device="cuda", generator=g).to(torch.bfloat16)[:, :P]
    got = ops.lq_pack(X)
    assert torch.equal(got.cpu(), lq_reference(X))


@pytest.mark.parametrize("M,K,R", [(1000, 768, 32), (64, 128, 8), (5, 2048, 64), (16384, 2048, 32), (777, 520, 16), (2, 64, 32)])
@pytest.mark.parametrize("f32", [False, True])
def test_stream_project(ops, M, K, R, f32):
    g = torch.Generator(device="cuda").manual_seed(M + K + R)
    x = torch.randn(M, K, device="cuda", generator=g)
    W = (torch.randn(R, K + 8, device="cuda", generator=g) / math.sqrt(K)).to(torch.bfloat16)[:, :K]      # strided weight view
    xb = x.to(torch.bfloat16)
    out = torch.full((M, R), 7.0, device="cuda", dtype=torch.bfloat16)
    lq = torch.full((ops.lq_words(M, R),), -1, device="cuda", dtype=torch.int32)
    if f32:
        copy = torch.full((M, K + 16), 7.0, device="cuda", dtype=torch.bfloat16)
        ops.stream_project(x, W, out=out, out_lq=lq, copy=copy[:, :K])
        assert torch.equal(copy[:, :K], xb)
        assert bool((copy[:, K:] == 7.0).all())
    else:
        ops.stream_project(xb, W, out=out, out_lq=lq)
    ref = xb.float() @ W.float().T
    assert rel(out.float(), ref) < 1e-2
    # the LQ output holds exactly the same bf16 values as the plain output
    assert torch.equal(lq.cpu(), lq_reference(out))
    # a capped grid (co-resident launch) gives the same result
    out2 = torch.empty_like(out)
    ops.stream_project(x if f32 else xb, W, out=out2, max_ctas=3)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("B,P,Q,tr,cs", [(1000, 32, 2048, False, True), (64, 32, 768, True, False), (5, 8, 256, False, True),
                                         (4097, 64, 2048, True, False), (333, 16, 520, False, True), (32768, 32, 2048, False, True),
                                         (20000, 32, 768, True, False)])
def test_stream_reduce(ops, B, P, Q, tr, cs):
    g = torch.Generator(device="cuda").manual_seed(B + P + Q)
    L = torch.randn(B, P, device="cuda", generator=g).to(torch.bfloat16)
    R = torch.randn(B, Q + 8, device="cuda", generator=g).to(torch.bfloat16)[:, :Q]            # strided view
    G = torch.ones((Q, P) if tr else (P, Q), device="cuda")
    colsum = torch.ones(Q, device="cuda") if cs else None
    ops.stream_reduce(ops.lq_pack(L), P, R, G, transpose_out=tr, colsum=colsum, scale=0.5)
    ref = 0.5 * (L.double().T @ R.double()).float()
    ref = (ref.T if tr else ref) + 1.0
    assert rel(G, ref) < 1e-4, rel(G, ref)
    if cs:
        assert rel(colsum, 0.5 * R.double().sum(0).float() + 1.0) < 1e-4
    G2 = torch.ones_like(G)
    ops.stream_reduce(ops.lq_pack(L), P, R, G2, transpose_out=tr, scale=0.5, max_ctas=5)
    assert rel(G2, ref) < 1e-4


def make_problem(B, D, H, r, seed):
    g = torch.Generator().manual_seed(seed)
    w1 = torch.randn(H, D, generator=g) / math.sqrt(D)
    b1 = torch.randn(H, generator=g) * 0.1
    w2 = torch.randn(H, H, generator=g) / math.sqrt(H)
    b2 = torch.randn(H, generator=g) * 0.1
    x = torch.randn(B, D, generator=g)
    x = x / x.norm(dim=1, keepdim=True) * math.sqrt(D) * 0.5
    a0 = torch.randn(D * r, generator=g) / math.sqrt(D)
    b0 = torch.randn(r * H, generator=g) * (0.5 / math.sqrt(r))
    a1 = torch.randn(H * r, generator=g) / math.sqrt(H)
    bb1 = torch.randn(r * H, generator=g) * (0.5 / math.sqrt(r))
    beta0 = torch.randn(H, generator=g) * 0.1
    beta1 = torch.randn(H, generator=g) * 0.1
    dy = torch.randn(B, H, generator=g) / math.sqrt(H)
    return dict(w1=w1, b1=b1, w2=w2, b2=b2, x=x, a=[a0, a1], b=[b0, bb1], beta=[beta0, beta1], dy=dy)


def run_merged(ops, p, B, D, H, r, grad_scale=1.0):
    dev = "cuda"
    c = lambda t: t.to(dev)
    if True:
        pk = ops.PackedProjector(D, H, r, dev, merged=True)
        pk.pack_adapter_merged(c(p["w1"]), c(p["w2"]), c(p["a"][0]), c(p["b"][0]), c(p["beta"][0]), c(p["a"][1]), c(p["b"][1]), c(p["beta"][1]),
                               c(p["b1"]), c(p["b2"]))
        st = ops.MlpStash(B, D, H, r, dev, full=True, merged=True)
        y = torch.full((B, H), float("nan"), device=dev)
        ops.adapted_mlp_fwd(pk, st, c(p["x"]), y)
        grads = dict(dA0=torch.zeros(D, r, device=dev), dB0=torch.zeros(r, H, device=dev), dbeta0=torch.zeros(H, device=dev),
                     dA1=torch.zeros(H, r, device=dev), dB1=torch.zeros(r, H, device=dev), dbeta1=torch.zeros(H, device=dev))
        ops.adapted_mlp_bwd(pk, st, c(p["dy"]), grads, grad_scale=grad_scale)
        torch.cuda.synchronize()
    return y, grads


@pytest.mark.parametrize("B,D,H,r", [(300, 768, 2048, 32), (4, 768, 2048, 32), (130, 64, 128, 8), (1111, 512, 2048, 64),
                                     (256, 1024, 2048, 16), (4096, 768, 2048, 32), (9000, 768, 2048, 32)])
def test_merged_schedule_matches_oracle(ops, B, D, H, r):
    """merged weights + rank-r side products computed inside the GEMMs (r <= 32) or by the row-panel pass (r = 64)"""
    p = make_problem(B, D, H, r, seed=B + D + r)
    y, g = run_merged(ops, p, B, D, H, r)
    y_ref, gr = O.adapted_mlp_full_grads(p["w1"], p["b1"], p["w2"], p["b2"], p["x"], p["a"], p["b"], p["beta"], p["dy"])
    dA0, dA1, dB0, dB1, dbeta0, dbeta1 = gr
    assert rel(y, y_ref) < TOL, ("y", rel(y, y_ref))
    for name, got, ref in [("dA0", g["dA0"], dA0.view(D, r)), ("dB0", g["dB0"], dB0.view(r, H)), ("dbeta0", g["dbeta0"], dbeta0),
                           ("dA1", g["dA1"], dA1.view(H, r)), ("dB1", g["dB1"], dB1.view(r, H)), ("dbeta1", g["dbeta1"], dbeta1)]:
        assert rel(got, ref) < TOL, (name, rel(got, ref))


def test_merged_schedule_repeatable_and_scaled(ops):
    """back-to-back steps on the same buffers and grad_scale applied to every gradient"""
    B, D, H, r = 2048, 768, 2048, 32
    p = make_problem(B, D, H, r, seed=5)
    y1, g1 = run_merged(ops, p, B, D, H, r)
    for _ in range(3):
        y2, g2 = run_merged(ops, p, B, D, H, r, grad_scale=0.25)
    assert torch.equal(y1, y2)
    for k in g1:
        assert rel(g2[k], 0.25 * g1[k]) < 1e-4, k
