"""Fused projection + batch-reduction pass (csrc/panel.cu dmi_panel_fused, csrc/panel_tc.cu dmi_panel_fused_tc) -- GPU parity.

One sweep over an activation gradient replaces dmi_skinny_rows + dmi_outer_reduce over the same matrix: (dv, dB1, dbeta1) from
dY and (du, dB0, dbeta0) from dpre, the autograd of the bmm pair and bias add of the reference's Projector.lora_forward
(dmi/model/projector.py:146-157).  Checked against torch fp32 on the bf16-rounded input (the kernels round the streamed matrix
to bf16 exactly once, so the bf16 copy is bit-exact and the fp32 reductions agree to accumulation order), and through the
adapted-MLP backward with the option on against the separate-pass schedule."""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

# Kernel variants written after the round's GPU budget was spent: compiled and wired behind dmi_set_option bits, never in the
# default schedule, and their tests only run on request until they have met a GPU.
experimental = pytest.mark.skipif(os.environ.get("DMI_EXPERIMENTAL") != "1", reason="unvalidated kernel variant: set DMI_EXPERIMENTAL=1")


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("M,K,R,f32,pad", [
    (64, 2048, 32, False, 0),      # exactly one panel
    (1, 1024, 16, True, 0),        # a single row: 63 zero-filled rows in the panel
    (1000, 2048, 32, False, 8),    # ragged last panel, padded leading dimension
    (1000, 2048, 32, True, 4),
    (4097, 1024, 16, False, 0),    # more panels than clusters, one row into the last panel
    (9000, 2048, 16, True, 0),
])
def test_panel_fused_matches_torch(M, K, R, f32, pad):
    from dmi_b200 import ops
    dev, bf = "cuda", torch.bfloat16
    g = torch.Generator(device=dev).manual_seed(M + K + R)
    base = torch.randn(M, K + pad, device=dev, generator=g) / 8
    inp = (base if f32 else base.to(bf))[:, :K]
    W = (torch.randn(R, K, device=dev, generator=g) / math.sqrt(K)).to(bf)
    L = torch.randn(M, R, device=dev, generator=g).to(bf)
    out = torch.full((M, R), 7.0, device=dev, dtype=bf)
    G0 = torch.randn(R, K, device=dev, generator=g)
    G, cs = G0.clone(), torch.ones(K, device=dev)          # both are accumulated into
    copy = torch.empty(M, K, device=dev, dtype=bf) if f32 else None
    ops.panel_fused(inp, W, L, out, G, colsum=cs, copy=copy, scale=0.5)
    xb = inp.to(bf).float()
    assert _rel(out, xb @ W.float().t()) < 6e-3              # bf16 output rounding
    assert _rel(G - G0, 0.5 * (L.float().t() @ xb)) < 1e-5   # fp32, accumulation order only
    assert _rel(cs - 1.0, 0.5 * xb.sum(0)) < 1e-5
    if f32:
        assert torch.equal(copy, inp.to(bf))                 # the bf16 operand copy is bit-exact


def test_panel_fused_without_colsum():
    from dmi_b200 import ops
    dev, bf = "cuda", torch.bfloat16
    g = torch.Generator(device=dev).manual_seed(3)
    M, K, R = 300, 2048, 32
    inp = (torch.randn(M, K, device=dev, generator=g) / 8).to(bf)
    W = (torch.randn(R, K, device=dev, generator=g) / math.sqrt(K)).to(bf)
    L = torch.randn(M, R, device=dev, generator=g).to(bf)
    out = torch.empty(M, R, device=dev, dtype=bf)
    G = torch.zeros(R, K, device=dev)
    ops.panel_fused(inp, W, L, out, G)
    assert _rel(G, L.float().t() @ inp.float()) < 1e-5


def test_panel_fused_rejects_unsupported_shapes():
    from dmi_b200 import ops
    dev, bf = "cuda", torch.bfloat16
    M, K, R = 64, 768, 32                                    # K must be 1024 or 2048
    z = lambda *s, dt=bf: torch.zeros(*s, device=dev, dtype=dt)
    with pytest.raises(RuntimeError):
        ops.panel_fused(z(M, K), z(R, K), z(M, R), z(M, R), z(R, K, dt=torch.float32))


@pytest.mark.parametrize("M,K,pad,colsum", [
    (128, 2048, 0, True),          # exactly one panel
    (1, 2048, 0, True),            # a single row: the rest of the TMA box is zero-filled
    (1000, 2048, 8, True),         # ragged last panel, padded leading dimension
    (4097, 1024, 0, True),         # K = 1024: 4 column tiles per CTA; one row into the last panel
    (300, 2048, 0, False),         # no column sum
    (20000, 2048, 0, True),        # more panels than clusters: accumulators persist over several panels per CTA
])
def test_panel_fused_tc_matches_torch(M, K, pad, colsum):
    """tcgen05 form: the TMA-swizzled tile is read as the K-major A operand (projection) and as the MN-major A operand (reductions)."""
    from dmi_b200 import ops
    dev, bf = "cuda", torch.bfloat16
    R = 32
    g = torch.Generator(device=dev).manual_seed(M + K)
    inp = (torch.randn(M, K + pad, device=dev, generator=g) / 8).to(bf)[:, :K]
    W = (torch.randn(R, K, device=dev, generator=g) / math.sqrt(K)).to(bf)
    L = torch.randn(M, R + 8, device=dev, generator=g).to(bf)[:, :R]      # row stride R + 8, like u inside [x | u]
    out = torch.full((M, R), 7.0, device=dev, dtype=bf)
    G0 = torch.randn(R, K, device=dev, generator=g)
    G, cs = G0.clone(), torch.ones(K, device=dev)
    ops.panel_fused_tc(inp, W, L, out, G, colsum=cs if colsum else None, scale=0.5)
    xb = inp.float()
    assert _rel(out, xb @ W.float().t()) < 6e-3
    assert _rel(G - G0, 0.5 * (L.float().t() @ xb)) < 1e-5
    if colsum:
        assert _rel(cs - 1.0, 0.5 * xb.sum(0)) < 1e-5
    else:
        assert torch.equal(cs, torch.ones(K, device=dev))


@experimental
@pytest.mark.parametrize("M,K", [(128, 2048), (1, 2048), (1000, 2048), (4097, 1024), (20000, 2048)])
def test_panel_fused_tc_merged_colsum_matches_torch(M, K):
    """variant with the column sum folded into the batch-reduction MMAs (ones column written into the staged L panel)"""
    from dmi_b200 import ops
    dev, bf = "cuda", torch.bfloat16
    R = 32
    g = torch.Generator(device=dev).manual_seed(M + K + 3)
    inp = (torch.randn(M, K, device=dev, generator=g) / 8).to(bf)
    W = (torch.randn(R, K, device=dev, generator=g) / math.sqrt(K)).to(bf)
    L = torch.randn(M, R + 8, device=dev, generator=g).to(bf)[:, :R]
    out = torch.full((M, R), 7.0, device=dev, dtype=bf)
    G, cs = torch.zeros(R, K, device=dev), torch.zeros(K, device=dev)
    ops.panel_fused_tc(inp, W, L, out, G, colsum=cs, scale=0.5, merged_colsum=True)
    xb = inp.float()
    assert _rel(out, xb @ W.float().t()) < 6e-3
    assert _rel(G, 0.5 * (L.float().t() @ xb)) < 1e-5
    assert _rel(cs, 0.5 * xb.sum(0)) < 1e-5


def test_panel_fused_tc_rejects_unsupported_shapes():
    from dmi_b200 import ops
    dev, bf = "cuda", torch.bfloat16
    z = lambda *s, dt=bf: torch.zeros(*s, device=dev, dtype=dt)
    with pytest.raises(RuntimeError):                        # rank 16 is not compiled for the tcgen05 form
        ops.panel_fused_tc(z(64, 2048), z(16, 2048), z(64, 16), z(64, 16), z(16, 2048, dt=torch.float32))
    with pytest.raises(RuntimeError):                        # K must be 1024 or 2048
        ops.panel_fused_tc(z(64, 768), z(32, 768), z(64, 32), z(64, 32), z(32, 768, dt=torch.float32))


@experimental
@pytest.mark.parametrize("M,K", [(128, 2048), (1, 768), (1000, 768), (4097, 1024), (20000, 2048)])
def test_panel_tc_project_matches_torch(M, K):
    """projection-only mode of the tcgen05 panel kernel (v = h A1, u = x A0), output into a column slice like [h | v]"""
    from dmi_b200 import ops
    dev, bf = "cuda", torch.bfloat16
    R = 32
    g = torch.Generator(device=dev).manual_seed(M + K)
    ext = (torch.randn(M, K + R, device=dev, generator=g) / 8).to(bf)
    inp, out = ext[:, :K], ext[:, K:]
    ref_in = inp.float().clone()
    W = (torch.randn(R, K, device=dev, generator=g) / math.sqrt(K)).to(bf)
    ops.panel_tc_project(inp, W, out)
    assert torch.equal(inp.float(), ref_in)
    assert _rel(out, ref_in @ W.float().t()) < 6e-3


@experimental
@pytest.mark.parametrize("M,K,transpose,colsum", [(128, 2048, False, True), (1000, 768, True, False), (4097, 1024, True, False),
                                                  (20000, 2048, False, True), (20000, 2048, True, False), (7, 768, False, False)])
def test_panel_tc_reduce_matches_torch(M, K, transpose, colsum):
    """reduction-only mode (dB1 + dbeta1, dA1, dA0): same contract as outer_reduce"""
    from dmi_b200 import ops
    dev, bf = "cuda", torch.bfloat16
    R = 32
    g = torch.Generator(device=dev).manual_seed(M + K + 1)
    inp = (torch.randn(M, K + 32, device=dev, generator=g) / 8).to(bf)[:, :K]
    L = torch.randn(M, R, device=dev, generator=g).to(bf)
    G0 = torch.randn((K, R) if transpose else (R, K), device=dev, generator=g)
    G, cs = G0.clone(), torch.ones(K, device=dev)
    ops.panel_tc_reduce(L, inp, G, transpose_out=transpose, colsum=cs if colsum else None, scale=0.25)
    ref = 0.25 * (L.float().t() @ inp.float())
    assert _rel(G - G0, ref.t() if transpose else ref) < 1e-5
    if colsum:
        assert _rel(cs - 1.0, 0.25 * inp.float().sum(0)) < 1e-5


@experimental
@pytest.mark.parametrize("M,K,pad", [(128, 2048, 0), (1, 2048, 0), (1000, 2048, 4), (4097, 1024, 0), (20000, 2048, 0)])
def test_panel_fused_tc32_matches_torch(M, K, pad):
    """fp32-input form (the dY pass): TMA-staged fp32 quarters converted in-kernel into the swizzled bf16 MMA tile + the bf16 copy"""
    from dmi_b200 import ops
    dev, bf = "cuda", torch.bfloat16
    R = 32
    g = torch.Generator(device=dev).manual_seed(M + K + 2)
    inp = (torch.randn(M, K + pad, device=dev, generator=g) / 8)[:, :K]
    W = (torch.randn(R, K, device=dev, generator=g) / math.sqrt(K)).to(bf)
    L = torch.randn(M, R + 8, device=dev, generator=g).to(bf)[:, :R]
    ext = torch.full((M, K + R), 7.0, device=dev, dtype=bf)          # [copy | out], like dyext = [dY | dv]
    copy, out = ext[:, :K], ext[:, K:]
    G0 = torch.randn(R, K, device=dev, generator=g)
    G, cs = G0.clone(), torch.ones(K, device=dev)
    ops.panel_fused_tc32(inp, W, L, out, G, colsum=cs, copy=copy, scale=0.5)
    xb = inp.to(bf).float()
    assert torch.equal(copy, inp.to(bf))
    assert _rel(out, xb @ W.float().t()) < 6e-3
    assert _rel(G - G0, 0.5 * (L.float().t() @ xb)) < 1e-5
    assert _rel(cs - 1.0, 0.5 * xb.sum(0)) < 1e-5


@pytest.mark.parametrize("B,opt", [(200, 1), (4096, 1), (200, 2), (4096, 2), (8192, -1),
                                   pytest.param(200, 18, marks=experimental), pytest.param(4096, 30, marks=experimental), pytest.param(4096, 34, marks=experimental),
                                   pytest.param(200, 6, marks=experimental), pytest.param(4096, 6, marks=experimental),
                                   pytest.param(200, 10, marks=experimental), pytest.param(4096, 14, marks=experimental)])
def test_adapted_mlp_backward_fused_schedule_matches_separate(B, opt):
    """dmi_set_option("fused_panel", v) swaps pairs of launches of the backward for a fused pass (1: mma.sync form over dY and dpre,
    2: tcgen05 form over dpre, -1: the default, which picks the tcgen05 form from 8192 rows up); gradients must agree."""
    from dmi_b200 import ops
    dev = "cuda"
    D, H, r = 768, 2048, 32
    g = torch.Generator(device=dev).manual_seed(B)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)
    z = lambda *s: torch.zeros(*s, device=dev)
    w1, w2, b1, b2 = rn(H, D) / math.sqrt(D), rn(H, H) / math.sqrt(H), rn(H) * 0.1, rn(H) * 0.1
    A0, B0, A1, B1 = rn(D * r) / math.sqrt(D), rn(r * H) * 0.1, rn(H * r) / math.sqrt(H), rn(r * H) * 0.1
    be0, be1 = rn(H) * 0.1, rn(H) * 0.1
    x, dy = rn(B, D), rn(B, H) / math.sqrt(H)
    y = torch.empty(B, H, device=dev)
    pk = ops.PackedProjector(D, H, r, dev)
    pk.pack_base(w1, w2)
    pk.pack_adapter(A0, B0, be0, A1, B1, be1, b1, b2)
    st = ops.MlpStash(B, D, H, r, dev, full=True)
    res = {}
    try:
        for o in (0, opt):
            ops.set_option("fused_panel", o)
            grads = dict(dA0=z(D, r), dB0=z(r, H), dbeta0=z(H), dA1=z(H, r), dB1=z(r, H), dbeta1=z(H))
            ops.adapted_mlp_fwd(pk, st, x, y)
            ops.adapted_mlp_bwd(pk, st, dy, grads)
            res[o] = grads
    finally:
        ops.set_option("fused_panel", -1)
    for k in res[0]:
        assert _rel(res[opt][k], res[0][k]) < 2e-3, k            # both are bf16-operand paths; du/dv round identically up to summation order
