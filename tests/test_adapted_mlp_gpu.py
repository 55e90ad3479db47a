"""Adapted MLP2 projector (C ABI dmi_adapted_mlp_fwd/bwd) against the CPU oracle on the same seeded inputs.

Tolerance: bf16 operands with fp32 accumulation -> 1e-2 relative (north_star) on outputs and gradients, measured as
||a-b|| / ||b|| per tensor."""
import math

import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-2


def rel(a, b):
    return ((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm().clamp_min(1e-30)).item()


def rows_rel(a, b):
    """worst row of a [B, *] matrix: a few wrong rows in a ragged last tile must not hide inside a Frobenius norm"""
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm(dim=1) / b.norm(dim=1).clamp_min(1e-6)).max().item()


def make_problem(B, D, H, r, seed):
    g = torch.Generator().manual_seed(seed)
    w1 = torch.randn(H, D, generator=g) / math.sqrt(D)
    b1 = torch.randn(H, generator=g) * 0.1
    w2 = torch.randn(H, H, generator=g) / math.sqrt(H)
    b2 = torch.randn(H, generator=g) * 0.1
    x = torch.randn(B, D, generator=g)
    x = x / x.norm(dim=1, keepdim=True) * math.sqrt(D) * 0.5     # O(1) pre-activations
    a0 = torch.randn(D * r, generator=g) / math.sqrt(D)
    b0 = torch.randn(r * H, generator=g) * (0.5 / math.sqrt(r))
    a1 = torch.randn(H * r, generator=g) / math.sqrt(H)
    bb1 = torch.randn(r * H, generator=g) * (0.5 / math.sqrt(r))
    beta0 = torch.randn(H, generator=g) * 0.1
    beta1 = torch.randn(H, generator=g) * 0.1
    dy = torch.randn(B, H, generator=g) / math.sqrt(H)
    return dict(w1=w1, b1=b1, w2=w2, b2=b2, x=x, a=[a0, a1], b=[b0, bb1], beta=[beta0, beta1], dy=dy)


def run_cuda(p, B, D, H, r, flags, full):
    from dmi_b200 import ops
    dev = "cuda"
    pk = ops.PackedProjector(D, H, r, dev)
    pk.pack_base(p["w1"].to(dev), p["w2"].to(dev))
    c = lambda t: t.to(dev)
    pk.pack_adapter(c(p["a"][0]), c(p["b"][0]), c(p["beta"][0]), c(p["a"][1]), c(p["b"][1]), c(p["beta"][1]), c(p["b1"]), c(p["b2"]))
    st = ops.MlpStash(B, D, H, r, dev, full=True)
    y = torch.full((B, H), float("nan"), device=dev)
    ops.adapted_mlp_fwd(pk, st, c(p["x"]), y, flags=flags)
    grads = dict(dA0=torch.zeros(D, r, device=dev), dB0=torch.zeros(r, H, device=dev), dbeta0=torch.zeros(H, device=dev))
    if full:
        grads.update(dA1=torch.zeros(H, r, device=dev), dB1=torch.zeros(r, H, device=dev), dbeta1=torch.zeros(H, device=dev))
    ops.adapted_mlp_bwd(pk, st, c(p["dy"]), grads, flags=flags)
    torch.cuda.synchronize()
    return y, grads


@pytest.mark.parametrize("B,D,H,r", [(300, 768, 2048, 32), (4, 768, 2048, 32), (130, 64, 128, 8), (1111, 512, 2048, 64),
                                     (256, 1024, 2048, 16)])
def test_full_adapted_mlp_fwd_bwd(B, D, H, r):
    p = make_problem(B, D, H, r, seed=B + D + r)
    y, g = run_cuda(p, B, D, H, r, flags=0, full=True)
    y_ref, gr = O.adapted_mlp_full_grads(p["w1"], p["b1"], p["w2"], p["b2"], p["x"], p["a"], p["b"], p["beta"], p["dy"])
    dA0, dA1, dB0, dB1, dbeta0, dbeta1 = gr
    assert rel(y, y_ref) < TOL, ("y", rel(y, y_ref))
    for name, got, ref in [("dA0", g["dA0"], dA0.view(D, r)), ("dB0", g["dB0"], dB0.view(r, H)), ("dbeta0", g["dbeta0"], dbeta0),
                           ("dA1", g["dA1"], dA1.view(H, r)), ("dB1", g["dB1"], dB1.view(r, H)), ("dbeta1", g["dbeta1"], dbeta1)]:
        assert rel(got, ref) < TOL, (name, rel(got, ref))


@pytest.mark.parametrize("B,D", [(8192, 768), (20000, 768), (32768, 768), (8192, 640), (20001, 640), (9000, 1024)])
def test_full_adapted_mlp_benchmark_schedule_vs_oracle(B, D):
    """The schedule bench.py times (from 8192 rows: CTA-pair GEMMs with TMA-store epilogues, tcgen05 projections, the fp32-input dY
    pass and the merged-column-sum dpre pass) against the oracle DIRECTLY (reference math: dmi/model/projector.py:61-116 + autograd),
    including ragged batches (20000 = 78.125 pair tiles, 20001 ends one row into a tile) and mm_dim 640 (8 reference configs).
    Frobenius error per tensor plus the worst ROW of y and the worst row/column of every gradient."""
    H, r = 2048, 32
    p = make_problem(B, D, H, r, seed=B + D)
    y, g = run_cuda(p, B, D, H, r, flags=0, full=True)
    y_ref, gr = O.adapted_mlp_full_grads(p["w1"], p["b1"], p["w2"], p["b2"], p["x"], p["a"], p["b"], p["beta"], p["dy"])
    dA0, dA1, dB0, dB1, dbeta0, dbeta1 = gr
    assert torch.isfinite(y).all()
    assert rel(y, y_ref) < TOL, ("y", rel(y, y_ref))
    assert rows_rel(y, y_ref) < 2 * TOL, ("y worst row", rows_rel(y, y_ref))
    for name, got, ref in [("dA0", g["dA0"], dA0.view(D, r)), ("dB0", g["dB0"], dB0.view(r, H)), ("dA1", g["dA1"], dA1.view(H, r)),
                           ("dB1", g["dB1"], dB1.view(r, H))]:
        assert rel(got, ref) < TOL, (name, rel(got, ref))
        assert rows_rel(got, ref) < 3 * TOL and rows_rel(got.t(), ref.t()) < 3 * TOL, (name, rows_rel(got, ref), rows_rel(got.t(), ref.t()))
    for name, got, ref in [("dbeta0", g["dbeta0"], dbeta0), ("dbeta1", g["dbeta1"], dbeta1)]:
        assert rel(got, ref) < TOL, (name, rel(got, ref))


@pytest.mark.parametrize("B,D,H,r", [(300, 768, 2048, 32), (4, 768, 2048, 32), (32, 64, 128, 8)])
def test_h1_mode_matches_lora_forward_as_written(B, D, H, r):
    """DMI_MLP_STOP_AFTER_FIRST_ACT reproduces Projector.lora_forward as written: y = gelu(pre)."""
    from dmi_b200 import _lib
    p = make_problem(B, D, H, r, seed=7 + B)
    y, g = run_cuda(p, B, D, H, r, flags=_lib.MLP_STOP_AFTER_FIRST_ACT, full=False)
    leaves = [t.clone().requires_grad_(True) for t in (p["a"][0], p["b"][0], p["beta"][0])]
    params = {"projector.net.0.weight": p["w1"], "projector.net.0.bias": p["b1"]}
    y_ref = O.lora_forward_as_written(params, p["x"], [leaves[0], p["a"][1]], [leaves[1], p["b"][1]], [leaves[2], p["beta"][1]])
    gA, gB, gbeta = torch.autograd.grad((y_ref * p["dy"]).sum(), leaves)
    assert rel(y, y_ref.detach()) < TOL
    assert rel(g["dA0"], gA.view(D, r)) < TOL
    assert rel(g["dB0"], gB.view(r, H)) < TOL
    assert rel(g["dbeta0"], gbeta) < TOL


def test_linearity_in_dy_at_full_size():
    """size-independent property at the benchmark shape: gradients are linear in dy (bwd(2*dy) == 2*bwd(dy))."""
    from dmi_b200 import ops
    B, D, H, r = 8192, 768, 2048, 32
    p = make_problem(B, D, H, r, seed=11)
    _, g1 = run_cuda(p, B, D, H, r, flags=0, full=True)
    p2 = dict(p)
    p2["dy"] = p["dy"] * 2
    _, g2 = run_cuda(p2, B, D, H, r, flags=0, full=True)
    for k in g1:
        assert rel(g2[k], 2 * g1[k]) < 2e-3, k
