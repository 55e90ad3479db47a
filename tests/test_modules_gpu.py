"""The nn.Module mirrors (dmi_b200.model) on the GPU against the golden vectors produced by the reference and against
the CPU oracle.  bf16 contractions: 1e-2 relative; fp32 kernels (merge): 1e-5; index work bit-exact."""
import math
import os
import tempfile

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-2


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def sd_of(d, prefix="sd/"):
    return {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}


def test_gemm_mn_weight_gradient_shape():
    from dmi_b200 import ops
    for (K, M, N) in [(1024, 2048, 768), (300, 256, 2048), (64, 128, 64), (5000, 2048, 2048)]:
        g = torch.Generator(device="cuda").manual_seed(K + M + N)
        a = torch.randn(K, M, device="cuda", generator=g).to(torch.bfloat16)
        b = torch.randn(K, N, device="cuda", generator=g).to(torch.bfloat16)
        out = torch.ones(M, N, device="cuda")
        ops.gemm_mn(a, b, out, alpha=0.5, accumulate=True)
        ref = 0.5 * (a.float().T @ b.float()) + 1.0
        assert rel(out, ref) < 2e-3, (K, M, N, rel(out, ref))


def test_merge_adapter_exact_fp32(golden_dir):
    from dmi_b200 import ops
    d = load(golden_dir, "fewshot_merged")
    D, _, H, r, alpha, n_tokens, K, B, N = [int(v) for v in d["meta"]]
    sd = sd_of(d)
    kw = dict(n_tokens=n_tokens, rank=r, alpha=float(alpha), lm_dim=H, mm_dim=D)
    a, b, bias = O.average_adapters([O.hypernetwork_forward(sd, z, **kw) for z in d["zs"]])
    c = lambda t: t.cuda()
    for i, name in enumerate(["0", "3"]):
        w, bb = ops.merge_adapter(c(sd[f"projector.net.{name}.weight"]), c(sd[f"projector.net.{name}.bias"]), c(a[i]), c(b[i]), c(bias[i]))
        assert rel(w, sd[f"generated_projector.{name}.weight"]) < 1e-5
        assert rel(bb, sd[f"generated_projector.{name}.bias"]) < 1e-5


def _projector(D, H, sd=None, dropout=0.1):
    from dmi_b200.model.projector import Projector
    from dmi_b200.utils.args import ProjectorArgs
    p = Projector(ProjectorArgs(proj_dropout=dropout), H, D, "cuda")
    if sd is not None:
        p.load_state_dict({k[len("projector."):]: v for k, v in sd.items() if k.startswith("projector.")})
    return p


def test_projector_mlp2_train_with_injected_dropout_mask(golden_dir):
    """Projector.forward (train_projector path) vs the reference run: same weights, same keep mask -> out and dW/db."""
    from dmi_b200.model.mlp2 import plain_mlp2
    d = load(golden_dir, "projector_mlp2")
    D, H, B = [int(v) for v in d["meta"]]
    p = _projector(D, H, sd_of(d))
    assert list(p.state_dict().keys()) == ["net.0.weight", "net.0.bias", "net.3.weight", "net.3.bias"]
    x, dy, keep = d["x"].cuda(), d["dy"].cuda(), d["keep"].cuda()
    out = plain_mlp2(x, p.net[0].weight, p.net[0].bias, p.net[3].weight, p.net[3].bias, dropout_p=0.1, keep=keep)
    assert rel(out, d["out"]) < TOL
    (out * dy).sum().backward()
    for k in ("net.0.weight", "net.0.bias", "net.3.weight", "net.3.bias"):
        got = dict(p.named_parameters())[k].grad
        assert rel(got, d["grad/" + k]) < TOL, (k, rel(got, d["grad/" + k]))
    p.eval()
    with torch.no_grad():
        assert rel(p(x), d["out_eval"]) < TOL


@pytest.mark.parametrize("B", [1024, 8192, 9001])
def test_projector_mlp2_large_vs_oracle(B):
    """plain MLP2 training step with an injected dropout mask against the oracle; from 8192 rows the two forward / backward-data GEMMs run
    on the CTA-pair kernel whose epilogue applies the keep mask (9001: ragged last tile), the weight gradients on the MN-major GEMM"""
    D, H = 768, 2048
    torch.manual_seed(0)
    p = _projector(D, H)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, D, generator=g)
    x = x / x.norm(dim=1, keepdim=True)
    dy = torch.randn(B, H, generator=g) / math.sqrt(H)
    keep = torch.rand(B, H, generator=g) >= 0.1
    from dmi_b200.model.mlp2 import plain_mlp2
    out = plain_mlp2(x.cuda(), p.net[0].weight, p.net[0].bias, p.net[3].weight, p.net[3].bias, dropout_p=0.1, keep=keep.cuda())
    (out * dy.cuda()).sum().backward()
    sd = {"projector." + k: v.detach().cpu().clone().requires_grad_(True) for k, v in p.state_dict().items()}
    ref = O.projector_forward(sd, x, drop_keep=keep, p_drop=0.1)
    (ref * dy).sum().backward()
    assert rel(out, ref) < TOL
    for k, v in p.named_parameters():
        assert rel(v.grad, sd["projector." + k].grad) < TOL, (k, rel(v.grad, sd["projector." + k].grad))


def test_plain_mlp2_in_place_gradient_accumulation_matches_autograd():
    """grad_in_place: dW, db accumulated straight into existing .grad tensors over two micro-steps == autograd's AccumulateGrad;
    the callback is told about every parameter once per backward; without existing .grad tensors the flag falls back to autograd"""
    from dmi_b200.model.mlp2 import plain_mlp2
    D, H, B = 768, 2048, 96
    torch.manual_seed(1)
    pa, pb = _projector(D, H), _projector(D, H)
    pb.load_state_dict(pa.state_dict())
    g = torch.Generator(device="cuda").manual_seed(2)
    xs = [torch.randn(B, D, device="cuda", generator=g) for _ in range(2)]
    dys = [torch.randn(B, H, device="cuda", generator=g) / math.sqrt(H) for _ in range(2)]
    keeps = [torch.rand(B, H, device="cuda", generator=g) >= 0.1 for _ in range(2)]
    told = []
    for q in pb.parameters():
        q.grad = torch.zeros_like(q)
    for x, dy, keep in zip(xs, dys, keeps):
        plain_mlp2(x, pa.net[0].weight, pa.net[0].bias, pa.net[3].weight, pa.net[3].bias, dropout_p=0.1, keep=keep).backward(dy)
        plain_mlp2(x, pb.net[0].weight, pb.net[0].bias, pb.net[3].weight, pb.net[3].bias, dropout_p=0.1, keep=keep,
                   grad_in_place=lambda q: told.append(q)).backward(dy)
    assert len(told) == 8 and {id(q) for q in told} == {id(q) for q in pb.parameters()}
    for (k, qa), qb in zip(pa.named_parameters(), pb.parameters()):
        assert rel(qb.grad, qa.grad) < 1e-5, (k, rel(qb.grad, qa.grad))
    pc = _projector(D, H)
    pc.load_state_dict(pa.state_dict())
    pc.grad_in_place = True
    pc.train()
    pc.net[2].p = 0.0
    pc(xs[0]).backward(dys[0])                    # no .grad yet -> ordinary autograd path
    assert all(q.grad is not None for q in pc.parameters())


def test_lora_wrapper_matches_reference(golden_dir):
    from dmi_b200.model.lora import LoraWrapper
    from dmi_b200.utils.args import LoraArgs, ProjectorArgs
    d = load(golden_dir, "lora_full")
    D, H, r, alpha, B = [int(v) for v in d["meta"]]
    sd = sd_of(d)
    with tempfile.NamedTemporaryFile(suffix=".pt") as f:
        torch.save({"projector_state_dict": {k[len("projector."):]: v for k, v in sd.items() if k.startswith("projector.")}}, f.name)
        w = LoraWrapper(LoraArgs(lora_rank=r, lora_alpha=alpha, lora_n_proj_layers=2), ProjectorArgs(proj_name_or_path=f.name), H, D, "cuda")
    assert sorted(w.state_dict().keys()) == sorted(sd.keys())
    w.load_state_dict(sd)
    w.train()
    out = w(d["x"].cuda())
    assert rel(out, d["out"]) < TOL
    (out * d["dy"].cuda()).sum().backward()
    for i in range(2):
        for n in ("A", "B"):
            got = getattr(w.lora_adapters.loras[i], n).grad
            ref = d[f"grad/lora_adapters.loras.{i}.{n}"]
            assert rel(got, ref) < TOL, (i, n, rel(got, ref))
    assert all(p.grad is None for p in w.projector.parameters())


def test_combine_lora_returns_merged_sequential(golden_dir):
    """combine_lora -> nn.Sequential with the reference's child indices; forward/backward through the fused MLP2 kernels."""
    d = load(golden_dir, "fewshot_merged")
    D, _, H, r, alpha, n_tokens, K, B, N = [int(v) for v in d["meta"]]
    sd = sd_of(d)
    kw = dict(n_tokens=n_tokens, rank=r, alpha=float(alpha), lm_dim=H, mm_dim=D)
    a, b, bias = O.average_adapters([O.hypernetwork_forward(sd, z, **kw) for z in d["zs"]])
    p = _projector(D, H, sd)
    p.eval()
    c = lambda ts: [t.cuda() for t in ts]
    merged = p.combine_lora(c(a), c(b), c(bias))
    assert isinstance(merged, torch.nn.Sequential)
    assert list(merged.state_dict().keys()) == ["0.weight", "0.bias", "3.weight", "3.bias"]
    assert merged[1] is p.net[1] and merged[2] is p.net[2]          # shared GELU / Dropout instances
    out = merged(d["x"].cuda())
    assert rel(out, d["out"]) < TOL
    (out * d["dy"].cuda()).sum().backward()
    for k, v in merged.named_parameters():
        assert rel(v.grad, d["grad/" + k]) < TOL, (k, rel(v.grad, d["grad/" + k]))
    with pytest.raises(ValueError):
        p.combine_lora(c(a)[:1], c(b)[:1], c(bias)[:1])
    with pytest.raises(ValueError):
        p.combine_lora(c(a) + c(a)[:1], c(b) + c(b)[:1], c(bias) + c(bias)[:1])


def test_two_adapters_in_flight_before_backward():
    """two lora_forward calls with different adapters, backward afterwards: each backward must use its own adapter's factors
    (the bf16 operand buffers are shared per projector and are re-packed when another forward has overwritten them)"""
    from dmi_b200.model.projector import Projector
    from dmi_b200.utils.args import ProjectorArgs
    D, H, r, B = 64, 128, 8, 96
    torch.manual_seed(0)
    proj = Projector(ProjectorArgs(), H, D, "cuda")
    proj.eval()
    proj.lora_forward_mode = "full"
    g = torch.Generator(device="cuda").manual_seed(1)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
    x = rn(B, D)
    dy = rn(B, H) / math.sqrt(H)

    def adapter(scale):
        return [(rn(D * r) * scale / math.sqrt(D)).requires_grad_(True), (rn(r * H) * 0.3).requires_grad_(True), (rn(H) * 0.1).requires_grad_(True),
                (rn(H * r) * scale / math.sqrt(H)).requires_grad_(True), (rn(r * H) * 0.3).requires_grad_(True), (rn(H) * 0.1).requires_grad_(True)]

    ad1, ad2 = adapter(1.0), adapter(3.0)
    fwd = lambda ad: proj.lora_forward(x, [ad[0], ad[3]], [ad[1], ad[4]], [ad[2], ad[5]])
    # reference gradients: one adapter at a time
    refs = []
    for ad in (ad1, ad2):
        refs.append([t.clone() for t in torch.autograd.grad((fwd(ad) * dy).sum(), ad)])
    # both forwards first, then both backwards (in the opposite order)
    y1, y2 = fwd(ad1), fwd(ad2)
    g2 = torch.autograd.grad((y2 * dy).sum(), ad2)
    g1 = torch.autograd.grad((y1 * dy).sum(), ad1)
    for got, ref in ((g1, refs[0]), (g2, refs[1])):
        for a, b in zip(got, ref):
            assert rel(a, b) < 1e-3, rel(a, b)


def test_lora_layer_standalone_forward_backward():
    """LoRALayer.forward = (alpha/r) x A B (reference dmi/model/lora.py:15-17) on the kernels, with gradients to A and B"""
    from dmi_b200.model.lora import LoRALayer
    torch.manual_seed(0)
    D, H, r, alpha, B = 768, 2048, 32, 16, 300
    layer = LoRALayer(D, H, r, alpha).cuda()
    with torch.no_grad():
        layer.B.normal_(0, 0.05)
    x = torch.randn(B, D, device="cuda")
    dy = torch.randn(B, H, device="cuda") / math.sqrt(H)
    y = layer(x)
    (y * dy).sum().backward()
    A, Bm = layer.A.detach().cpu().double().requires_grad_(True), layer.B.detach().cpu().double().requires_grad_(True)
    ref = (alpha / r) * (x.cpu().double() @ A @ Bm)
    (ref * dy.cpu().double()).sum().backward()
    assert rel(y, ref.detach()) < TOL
    assert rel(layer.A.grad, A.grad) < TOL and rel(layer.B.grad, Bm.grad) < TOL


def test_h6_clip_includes_frozen_projector_matches_reference_autograd(golden_dir):
    """the reference clips over the whole HyperNetWrapper (train_hypernet.py:148), whose frozen-by-omission projector keeps
    accumulating net.0.{weight,bias}.grad because only the hypernet's gradients are ever zeroed.  With
    clip_includes_frozen_projector the kernels reproduce that: three identical micro-steps with the hypernet gradients zeroed in
    between -> projector.net.0 gradients = 1x, 2x, 3x the oracle's autograd gradient, and the clip norm over clip_parameters() grows
    exactly like the reference's; with the flag off (default) the projector receives nothing."""
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from test_hypernet_gpu import build_wrapper
    d = load(golden_dir, "hypernet_h1_full_ctx")
    D_hyp, D_mm, H, r, alpha, n_tokens, K, B, prune = [int(v) for v in d["meta"][:9]]
    w = build_wrapper(d)
    w.eval()
    x, z, dy = d["x"].cuda(), d["z"].cuda(), d["dy"].cuda()
    # oracle: the reference's autograd with the projector's first layer requiring grad
    sd = {k: v.clone() for k, v in sd_of(d).items()}
    names = ["projector.net.0.weight", "projector.net.0.bias"]
    hyper = [k for k in sd if k.startswith("hypernet.") and "generators.1" not in k and "pos_encs" not in k]
    leaves = {k: sd[k].clone().requires_grad_(True) for k in names + hyper}
    full = dict(sd)
    full.update(leaves)
    out = O.hypernet_wrapper_forward(full, d["x"], d["z"], n_tokens=n_tokens, rank=r, alpha=float(alpha), lm_dim=H, mm_dim=D_mm)
    (out * d["dy"]).sum().backward()
    g_proj = {k: leaves[k].grad for k in names}
    hyper_sq = sum(float(leaves[k].grad.double().pow(2).sum()) for k in hyper)
    proj_sq = sum(float(g.double().pow(2).sum()) for g in g_proj.values())
    # default: nothing lands in the projector
    (w(x, z) * dy).sum().backward()
    assert all(p.grad is None for p in w.projector.parameters())
    assert {id(p) for p in w.clip_parameters()} == {id(p) for p in w.hypernet.parameters()}
    for p in w.hypernet.parameters():
        p.grad = None
    w.clip_includes_frozen_projector = True
    assert {id(p) for p in w.clip_parameters()} == {id(p) for p in w.parameters()}
    for step in (1, 2, 3):
        for p in w.hypernet.parameters():          # optimizer.zero_grad() of the reference: hypernet parameters only
            p.grad = None
        (w(x, z) * dy).sum().backward()
        for k in names:
            got = dict(w.named_parameters())[k].grad
            assert rel(got, step * g_proj[k]) < TOL, (step, k)
        total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(p.grad) for p in w.clip_parameters() if p.grad is not None]))
        want = math.sqrt(hyper_sq + step * step * proj_sq)
        assert abs(float(total) - want) / want < TOL, (step, float(total), want)
    assert w.projector.net[3].weight.grad is None      # as-written path: the second Linear is never applied (H1)
