"""tcgen05 row-panel kernels for the rank-r side products of the adapted-MLP step (csrc/panel_tc.cu, csrc/panel_tc32.cu) -- GPU parity.

One sweep over an activation gradient replaces dmi_skinny_rows + dmi_outer_reduce over the same matrix: (dv, dB1, dbeta1) and the bf16
copy from the fp32 dY, (du, dB0, dbeta0) from dpre -- the autograd of the bmm pair and bias add of the reference's
Projector.lora_forward (dmi/model/projector.py:146-157).  Checked against torch fp32 on the bf16-rounded input (the kernels round the
streamed matrix to bf16 exactly once, so the bf16 copy is bit-exact and the fp32 reductions agree to accumulation order), row by row
for the projection, and through the adapted-MLP backward against the separate-pass (mma.sync) schedule."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


def _rows_rel(a, b):
    d = (a.float() - b.float()).norm(dim=1)
    return (d / b.float().norm(dim=1).clamp_min(1e-3)).max().item()


@pytest.mark.parametrize("M,K,pad,colsum", [
    (128, 2048, 0, True),          # exactly one panel
    (1, 2048, 0, True),            # a single row: the rest of the TMA box is zero-filled
    (1000, 2048, 8, True),         # ragged last panel, padded leading dimension
    (4097, 1024, 0, True),         # K = 1024: 4 column tiles per CTA; one row into the last panel
    (300, 2048, 0, False),         # no column sum
    (20000, 2048, 0, True),        # more panels than clusters: accumulators persist over several panels per CTA
    (32768, 2048, 0, True),        # the benchmark shape
])
def test_panel_fused_tc_matches_torch(M, K, pad, colsum):
    """the TMA-swizzled tile is read as the K-major A operand (projection) and as the MN-major A operand (reductions); the column sum
    rides in the batch-reduction MMAs through a ones column written into the staged L panel"""
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
    ref = xb @ W.float().t()
    assert _rel(out, ref) < 6e-3 and _rows_rel(out, ref) < 3e-2     # bf16 output rounding; worst row
    assert _rel(G - G0, 0.5 * (L.float().t() @ xb)) < 1e-5          # fp32, accumulation order only
    if colsum:
        assert _rel(cs - 1.0, 0.5 * xb.sum(0)) < 1e-5
    else:
        assert torch.equal(cs, torch.ones(K, device=dev))


def test_panel_fused_tc_rejects_unsupported_shapes():
    from dmi_b200 import ops
    dev, bf = "cuda", torch.bfloat16
    z = lambda *s, dt=bf: torch.zeros(*s, device=dev, dtype=dt)
    with pytest.raises(RuntimeError):                        # rank 16 is not compiled for the tcgen05 form
        ops.panel_fused_tc(z(64, 2048), z(16, 2048), z(64, 16), z(64, 16), z(16, 2048, dt=torch.float32))
    with pytest.raises(RuntimeError):                        # K must be 1024 or 2048
        ops.panel_fused_tc(z(64, 768), z(32, 768), z(64, 32), z(64, 32), z(32, 768, dt=torch.float32))


@pytest.mark.parametrize("M,K", [(128, 2048), (1, 768), (1000, 768), (4097, 1024), (20000, 2048), (32768, 768)])
def test_panel_tc_project_matches_torch(M, K):
    """projection-only mode (v = h A1, u = x A0), output into a column slice like [h | v]"""
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
    ref = ref_in @ W.float().t()
    assert _rel(out, ref) < 6e-3 and _rows_rel(out, ref) < 3e-2


@pytest.mark.parametrize("M,K,pad", [(128, 2048, 0), (1, 2048, 0), (1000, 2048, 4), (4097, 1024, 0), (20000, 2048, 0), (32768, 2048, 0)])
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
    ref = xb @ W.float().t()
    assert _rel(out, ref) < 6e-3 and _rows_rel(out, ref) < 3e-2
    assert _rel(G - G0, 0.5 * (L.float().t() @ xb)) < 1e-5
    assert _rel(cs - 1.0, 0.5 * xb.sum(0)) < 1e-5


@pytest.mark.parametrize("B,D", [(200, 768), (4096, 768), (8192, 768), (8192, 640), (20001, 1024)])
def test_adapted_mlp_panel_schedule_matches_separate_passes(B, D):
    """dmi_set_option("fused_panel", v): 0 = separate mma.sync side passes (skinny_rows + outer_reduce, the small-batch schedule),
    1 = tcgen05 panel passes at any size, -1 = auto (panel passes from 8192 rows).  Outputs and gradients must agree."""
    from dmi_b200 import ops
    dev = "cuda"
    H, r = 2048, 32
    g = torch.Generator(device=dev).manual_seed(B)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)
    z = lambda *s: torch.zeros(*s, device=dev)
    w1, w2, b1, b2 = rn(H, D) / math.sqrt(D), rn(H, H) / math.sqrt(H), rn(H) * 0.1, rn(H) * 0.1
    A0, B0, A1, B1 = rn(D * r) / math.sqrt(D), rn(r * H) * 0.1, rn(H * r) / math.sqrt(H), rn(r * H) * 0.1
    be0, be1 = rn(H) * 0.1, rn(H) * 0.1
    x, dy = rn(B, D), rn(B, H) / math.sqrt(H)
    pk = ops.PackedProjector(D, H, r, dev)
    pk.pack_base(w1, w2)
    pk.pack_adapter(A0, B0, be0, A1, B1, be1, b1, b2)
    st = ops.MlpStash(B, D, H, r, dev, full=True)
    res = {}
    try:
        for o in (0, 1, -1):
            ops.set_option("fused_panel", o)
            y = torch.empty(B, H, device=dev)
            grads = dict(dA0=z(D, r), dB0=z(r, H), dbeta0=z(H), dA1=z(H, r), dB1=z(r, H), dbeta1=z(H))
            ops.adapted_mlp_fwd(pk, st, x, y)
            ops.adapted_mlp_bwd(pk, st, dy, grads)
            res[o] = (y, grads)
    finally:
        ops.set_option("fused_panel", -1)
    for o in (1, -1):
        assert _rel(res[o][0], res[0][0]) < 2e-3
        for k in res[0][1]:
            assert _rel(res[o][1][k], res[0][1][k]) < 2e-3, (o, k)      # both are bf16-operand paths; u/v/du/dv round identically up to summation order
