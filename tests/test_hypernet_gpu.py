"""Hypernetwork side of the path on the GPU (augmentation, pooling, generators, splice, wrapper) against the golden
vectors produced by running the reference, and against the CPU oracle at the real widths.

fp32 kernels (normalise, pooling, generator GEMV, merge, splice): 1e-5 relative; the 3xTF32 rotation: 1e-5;
anything that goes through the bf16 projector GEMMs: 1e-2; index / data-movement work: bit-exact."""
import math
import os
import tempfile

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
F32 = 1e-5
BF16 = 1e-2


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def sd_of(d):
    return {k[3:]: v for k, v in d.items() if k.startswith("sd/")}


def build_wrapper(d, full_mode=False):
    from dmi_b200.model.hypernet import HyperNetWrapper
    from dmi_b200.utils.args import HypnetArgs, ProjectorArgs
    D_hyp, D_mm, H, r, alpha, n_tokens, K, B, prune = [int(v) for v in d["meta"][:9]]
    sd = sd_of(d)
    proj_sd = {k[len("projector."):]: v for k, v in sd.items() if k.startswith("projector.")}
    with tempfile.NamedTemporaryFile(suffix=".pt") as f:
        # the reference loads the FULL-width projector checkpoint and prunes its first layer on load (projector.py:49-52)
        if prune >= 0 and proj_sd["net.0.weight"].shape[1] == prune:
            proj_sd = dict(proj_sd)
            proj_sd["net.0.weight"] = torch.nn.functional.pad(proj_sd["net.0.weight"], (0, D_hyp - prune))
        torch.save({"projector_state_dict": proj_sd}, f.name)
        w = HyperNetWrapper(HypnetArgs(hn_arch="attention", hn_hypnet_dim=D_hyp, hn_rank=r, hn_alpha=alpha, hn_predict_bias=True,
                                       hn_n_proj_layers=2, hn_use_pos_encs=True),
                            ProjectorArgs(proj_name_or_path=f.name, proj_prune=None if prune < 0 else prune), H, D_mm, n_tokens, "cuda")
    missing = w.load_state_dict({k: v for k, v in sd.items() if not k.startswith("generated_projector")}, strict=True)
    return w


@pytest.mark.parametrize("name", ["hypernet_h1_full_ctx", "hypernet_h1_masked", "hypernet_h1_pruned", "hypernet_h1_dropout"])
def test_hypernet_wrapper_matches_reference_run(golden_dir, name):
    """HyperNetWrapper.forward(x, z) + backward: same state-dict, same inputs as the reference run that made the fixture."""
    d = load(golden_dir, name)
    w = build_wrapper(d)
    assert sorted(w.state_dict().keys()) == sorted(k for k in sd_of(d))
    keep = d.get("keep_mask")
    x, z, dy = d["x"].cuda(), d["z"].cuda(), d["dy"].cuda()
    if keep is not None:
        w.train()
        a_w, b_w, biases = w.hypernet(z, keep_mask=keep[0, :2].cuda())       # rows 0..1 of the [S,S] mask are the ones that matter
        out = w.projector.lora_forward(x, a_w, b_w, biases)
    else:
        w.eval()
        out = w(x, z)
        a_w, b_w, biases = w.hypernet(z)
        for i in range(2):                                                    # the generated adapter itself is fp32-exact
            assert rel(a_w[i], d[f"adapter/a{i}"]) < F32 and rel(b_w[i], d[f"adapter/b{i}"]) < F32 and rel(biases[i], d[f"adapter/bias{i}"]) < F32
    assert rel(out, d["out"]) < BF16, rel(out, d["out"])
    (out * dy).sum().backward()
    params = dict(w.named_parameters())
    for k, ref in d.items():
        if not k.startswith("grad/"):
            continue
        p = params[k[5:]]
        if ref.numel() == 0:
            assert p.grad is None, f"{k} must not receive a gradient (SURVEY H1)"
            continue
        if k.endswith("projector.net.0.weight") or k.endswith("projector.net.0.bias"):
            continue              # the reference's autograd also fills the frozen projector's .grad; the kernels do not
        if ref.double().norm() < 1e-7:
            assert p.grad is None or p.grad.double().norm().item() < 1e-5, k
            continue
        assert p.grad is not None, k
        assert rel(p.grad, ref) < BF16, (k, rel(p.grad, ref))


def test_hypernetwork_forward_fp32_exact_at_real_width():
    """D=768, r=32, H=2048, K=128 (v4 shape): generated adapter vs the oracle, 1e-5; and the key-masked short-support case."""
    from dmi_b200.model.hypernet import HyperNetwork
    from dmi_b200.utils.args import HypnetArgs
    torch.manual_seed(0)
    D, H, r, n_tokens = 768, 2048, 32, 128
    hn = HyperNetwork(HypnetArgs(hn_arch="attention", hn_hypnet_dim=D, hn_rank=r, hn_alpha=32, hn_n_proj_layers=2, hn_use_pos_encs=True), H, D, n_tokens, "cuda")
    hn.eval()
    with torch.no_grad():
        for gnr in hn.generators:
            gnr.bias.normal_(0, 0.02)
    sd = {"hypernet." + k: v.detach().cpu() for k, v in hn.state_dict().items()}
    g = torch.Generator().manual_seed(1)
    for K in (128, 32):
        z = torch.randn(1 + 2 * K, D, generator=g)
        z = z / z.norm(dim=1, keepdim=True)
        a_w, b_w, biases = hn(z.cuda())
        ra, rb, rbias = O.hypernetwork_forward(sd, z, n_tokens=n_tokens, rank=r, alpha=32.0, lm_dim=H, mm_dim=D)
        for i in range(2):
            assert rel(a_w[i], ra[i]) < F32 and rel(b_w[i], rb[i]) < F32 and rel(biases[i], rbias[i]) < F32, (K, i)


@pytest.mark.parametrize("D,H,r,n_tokens,K,pos,p_keep", [(520, 256, 8, 8, 5, False, None), (1024, 512, 16, 40, 40, True, 0.9),
                                                          (768, 512, 32, 128, 17, True, None), (64, 128, 8, 4, 4, True, 0.8)])
def test_hypernetwork_shapes_forward_backward_vs_oracle_autograd(D, H, r, n_tokens, K, pos, p_keep):
    """the cooperative pooling kernels + generators at other widths / support sizes (D not a multiple of 128, short key-masked
    supports, no positional encodings, injected attention-dropout mask): outputs AND every parameter gradient against the oracle's
    autograd, fp32 tolerance"""
    from dmi_b200.model.hypernet import HyperNetwork
    from dmi_b200.utils.args import HypnetArgs
    torch.manual_seed(D + K)
    hn = HyperNetwork(HypnetArgs(hn_arch="attention", hn_hypnet_dim=D, hn_rank=r, hn_alpha=16, hn_n_proj_layers=2, hn_use_pos_encs=pos), H, D, n_tokens, "cuda")
    hn.train()
    with torch.no_grad():
        for gnr in hn.generators:
            gnr.bias.normal_(0, 0.02)
        for lin in (hn.hypnet.q, hn.hypnet.k, hn.hypnet.v):
            lin.bias.normal_(0, 0.1)
    sd = {"hypernet." + k: v.detach().cpu().clone().requires_grad_(v.dtype.is_floating_point and "pos_encs" not in k) for k, v in hn.state_dict().items()}
    g = torch.Generator().manual_seed(7)
    z = torch.randn(1 + 2 * K, D, generator=g)
    z = z / z.norm(dim=1, keepdim=True)
    S = 2 + z.shape[0]
    # the oracle scales kept weights by 1 / (1 - 0.05) (nn.Dropout(0.05), hypernet.py:54); the mask content is free
    keep2 = (torch.rand(2, S, generator=g) < p_keep) if p_keep is not None else None
    Sp = max(S, 2 * n_tokens + 3)
    full = None
    if keep2 is not None:
        full = torch.ones(1, Sp, Sp, dtype=torch.bool)
        full[0, :2, :S] = keep2
        a_w, b_w, biases = hn(z.cuda(), keep_mask=keep2.cuda())
    else:
        hn.eval()                                   # no mask given: eval-mode forward (train mode would draw one)
        a_w, b_w, biases = hn(z.cuda())
    ra, rb, rbias = O.hypernetwork_forward(sd, z, n_tokens=n_tokens, rank=r, alpha=16.0, lm_dim=H, mm_dim=D, use_pos_encs=pos, keep_mask=full)
    outs = list(a_w) + list(b_w) + list(biases)
    refs = list(ra) + list(rb) + list(rbias)
    for i, (o, rf) in enumerate(zip(outs, refs)):
        assert rel(o, rf) < F32, (i, rel(o, rf))
    cot = [torch.randn(o.shape, generator=g) / math.sqrt(o.numel()) for o in outs]
    torch.autograd.backward(outs, [c.cuda() for c in cot])
    torch.autograd.backward(refs, cot)
    for k, q in hn.named_parameters():
        ref = sd["hypernet." + k].grad
        assert q.grad is not None and ref is not None, k
        assert ((q.grad.cpu().double() - ref.double()).norm() / ref.double().norm().clamp_min(1e-3)).item() < 2e-5, (k, rel(q.grad, ref))


def test_augment_matches_oracle_and_is_bit_exact_on_data_movement():
    from dmi_b200 import augment as A
    g = torch.Generator().manual_seed(3)
    B, K, D = 37, 19, 768
    mm, m, t, p = torch.randn(B, D, generator=g), torch.randn(K, D, generator=g), torch.randn(K, D, generator=g), torch.randn(1, D, generator=g)
    R = O.get_rotation_matrix(D, np.random.RandomState(42))
    c = lambda x: x.cuda()
    # 1. no rotation, already-normalised inputs: pure interleave -> bit-exact
    mm_n, m_n, t_n, p_n = (O.l2_normalize(x) for x in (mm, m, t, p))
    x2, z2 = A.process_embeddings(c(mm_n), (c(m_n), c(t_n), c(p_n)), R=None)
    xo, zo = O.process_embeddings(mm_n, (m_n, t_n, p_n), None)
    assert torch.equal(x2.cpu(), xo) and torch.equal(z2.cpu(), zo)
    assert torch.equal(A.interleave_embeddings(c(m_n), c(t_n)).cpu(), O.interleave_embeddings(m_n, t_n))
    # 2. L2 normalisation kernel vs x / x.norm()
    assert rel(A.l2_normalize(c(mm)), mm_n) < 1e-6
    # 3. rotation on tf32 tensor cores with the 3-term split: fp32-accurate, text / prefix rows untouched
    x3, z3 = A.process_embeddings(c(mm_n), (c(m_n), c(t_n), c(p_n)), R=c(R))
    xo, zo = O.process_embeddings(mm_n, (m_n, t_n, p_n), R)
    assert rel(x3, xo) < F32 and rel(z3, zo) < F32, (rel(x3, xo), rel(z3, zo))
    assert torch.equal(z3[0].cpu(), p_n[0]) and torch.equal(z3[2::2].cpu(), t_n)
    # more than 512 batch rows: the batch and the support rows are two GEMM launches again (<= 512: one launch with a row split)
    big = O.l2_normalize(torch.randn(700, D, generator=g))
    x3b, z3b = A.process_embeddings(c(big), (c(m_n), c(t_n), c(p_n)), R=c(R))
    assert rel(x3b, O.process_embeddings(big, (m_n, t_n, p_n), R)[0]) < F32 and rel(z3b, z3.cpu()) < 1e-6
    # isometry: row norms and the Gram matrix are preserved
    assert (x3.norm(dim=1) - 1).abs().max().item() < 1e-5
    assert rel(x3.double() @ x3.double().T, mm_n.double() @ mm_n.double().T) < 1e-4
    # 4. fused normalise + rotate == normalise then rotate; bf16 copy for the projector operand
    xb = torch.zeros(B, D + 32, dtype=torch.bfloat16, device="cuda")
    x4, z4 = A.process_embeddings(c(mm), (c(m), c(t), c(p)), R=c(R), normalize=True, mm_out_bf16=xb[:, :D])
    assert rel(x4, xo) < F32 and rel(z4, zo) < F32
    assert torch.equal(xb[:, :D], x4.to(torch.bfloat16)) and xb[:, D:].abs().max().item() == 0
    # 5. permutation + sign flip are exact special cases of the isometry: bit-exact data movement
    perm = torch.randperm(D, generator=g)
    sign = (torch.randint(0, 2, (D,), generator=g) * 2 - 1).float()
    x5, z5 = A.process_embeddings(c(mm_n), (c(m_n), c(t_n), c(p_n)), R=None, perm=c(perm), sign=c(sign))
    assert torch.equal(x5.cpu(), mm_n[:, perm] * sign)
    assert torch.equal(z5[1::2].cpu(), m_n[:, perm] * sign)
    # 6. pruned projector: encoder narrower than the hypernet -> support rows zero-padded on the right (train_hypernet.py:99-100)
    d = 512
    Rd = O.get_rotation_matrix(d, np.random.RandomState(7))
    mmd, md = O.l2_normalize(torch.randn(B, d, generator=g)), O.l2_normalize(torch.randn(K, d, generator=g))
    x6, z6 = A.process_embeddings(c(mmd), (c(md), c(t_n), c(p_n)), R=c(Rd), prune=d, finetune_mm_dim=D)
    xo, zo = O.process_embeddings(mmd, (md, t_n, p_n), Rd, prune=d, finetune_mm_dim=D)
    assert rel(x6, xo) < F32 and rel(z6, zo) < F32 and z6[1::2, d:].abs().max().item() == 0


def test_splice_bit_exact(golden_dir):
    from dmi_b200.model.mmmodel import splice_prefix
    d = load(golden_dir, "splice")
    table = d["table"].to(torch.bfloat16).cuda()
    emb, mask, lab = splice_prefix(d["projected"].cuda(), table, d["ids"].cuda(), d["attn"].cuda(), d["labels"].cuda())
    assert emb.dtype == torch.float32 and torch.equal(emb.cpu(), d["embeds"])
    assert torch.equal(lab.cpu(), d["labels_out"])
    assert torch.equal(mask.cpu(), torch.cat((torch.ones(d["attn"].shape[0], 1), d["attn"]), -1))
    # bf16 mode (north_star): same indices, values rounded once
    emb16, _, _ = splice_prefix(d["projected"].cuda(), table, d["ids"].cuda(), None, None, embeds_dtype=torch.bfloat16)
    assert torch.equal(emb16.cpu(), d["embeds"].to(torch.bfloat16))
    # larger, with gradient to the prefix slot only
    g = torch.Generator().manual_seed(0)
    B, T, H, V = 5, 77, 2048, 1000
    proj = torch.randn(B, H, generator=g).cuda().requires_grad_(True)
    tab = torch.randn(V, H, generator=g).to(torch.bfloat16).cuda()
    ids = torch.randint(0, V, (B, T), generator=g).cuda()
    e, _, _ = splice_prefix(proj, tab, ids)
    ref = torch.cat((proj.detach().unsqueeze(1), tab[ids].float()), 1)
    assert torch.equal(e, ref)
    wgt = torch.randn(B, 1 + T, H, generator=g).cuda()
    (e * wgt).sum().backward()
    assert torch.equal(proj.grad, wgt[:, 0, :])


def test_fewshot_pipeline_generate_merge_finetune(golden_dir):
    """generate_projector_from_multiple_adapters -> merged projector (fp32 exact) -> forward/backward of the merged MLP2"""
    d = load(golden_dir, "fewshot_merged")
    meta = torch.cat([d["meta"][:8], torch.tensor([-1])])
    d2 = dict(d)
    d2["meta"] = meta
    w = build_wrapper(d2)
    w.eval()
    w.generate_projector_from_multiple_adapters([z.cuda() for z in d["zs"]])
    sd = sd_of(d)
    for k in ("0.weight", "0.bias", "3.weight", "3.bias"):
        assert rel(w.generated_projector.state_dict()[k], sd["generated_projector." + k]) < F32, k
    out = w(d["x"].cuda(), None)
    assert rel(out, d["out"]) < BF16
    (out * d["dy"].cuda()).sum().backward()
    for k, v in w.generated_projector.named_parameters():
        assert rel(v.grad, d["grad/" + k]) < BF16, k
    assert [p.shape for p in w.trainable_parameters()] == [v.shape for v in w.generated_projector.parameters()]


def test_mean_adapter_equals_mean_of_adapters_at_real_width():
    """8f-2: one generator pass over the mean modality code == element-wise mean of N separately generated adapters
    (linearity of the generators), support sets of different lengths, real widths (D=768, H=2048, r=32)."""
    import tempfile
    from dmi_b200.model.hypernet import HyperNetWrapper
    from dmi_b200.model.projector import Projector
    from dmi_b200.utils.args import HypnetArgs, ProjectorArgs
    D, H, r = 768, 2048, 32
    torch.manual_seed(3)
    base = Projector(ProjectorArgs(), H, D, "cuda")
    with tempfile.NamedTemporaryFile(suffix=".pt") as f:
        torch.save({"projector_state_dict": base.state_dict()}, f.name)
        w = HyperNetWrapper(HypnetArgs(hn_arch="attention", hn_hypnet_dim=D, hn_rank=r, hn_alpha=32, hn_n_proj_layers=2, hn_use_pos_encs=True),
                            ProjectorArgs(proj_name_or_path=f.name), H, D, 128, "cuda")
    w.eval()
    g = torch.Generator(device="cuda").manual_seed(0)
    zs = [torch.randn(1 + 2 * k, D, device="cuda", generator=g) for k in (32, 32, 7, 128, 1)]
    with torch.no_grad():
        singles = [w.hypernet(z) for z in zs]
        a_m, b_m, bias_m = w.hypernet.mean_adapter(zs)
    for l in range(2):
        for got, idx in ((a_m, 0), (b_m, 1), (bias_m, 2)):
            ref = torch.stack([s[idx][l] for s in singles]).mean(0)
            assert rel(got[l], ref) < F32, (l, idx, rel(got[l], ref))


def _real_width_hypernet():
    from dmi_b200.model.hypernet import HyperNetwork
    from dmi_b200.utils.args import HypnetArgs
    torch.manual_seed(3)
    return HyperNetwork(HypnetArgs(hn_arch="attention", hn_hypnet_dim=768, hn_rank=32, hn_alpha=32, hn_predict_bias=True, hn_n_proj_layers=2,
                                   hn_use_pos_encs=True), 2048, 768, 128, "cuda")


def test_generator_gradient_factor_mode_and_in_place_accumulation_match_dense():
    """Three ways to the same generator gradients at the real widths (92160 x 768 and 133120 x 768), two micro-steps accumulated:
    (a) autograd's dense accumulation, (b) the kernel accumulating into .grad in place with GradSync told through the ready callback,
    (c) rank-1 factor mode (the backward only READS G; (alpha/r dw, e) pairs are turned into the dense gradient once)."""
    from dmi_b200.parallel import GradSync, Rank1FactorSync
    hn = _real_width_hypernet()
    hn.eval()                                                         # no attention dropout: the three runs must see the same function

    def rel(a, b):          # hypnet.k.bias has an exactly-zero gradient in exact arithmetic (softmax shift invariance): absolute floor
        a, b = a.detach().double(), b.detach().double()
        return ((a - b).norm() / b.norm().clamp_min(1e-3)).item()
    g = torch.Generator(device="cuda").manual_seed(0)
    zs = [torch.nn.functional.normalize(torch.randn(257, 768, device="cuda", generator=g), dim=1) for _ in range(2)]

    def run():
        for z in zs:
            a_w, b_w, biases = hn(z)
            loss = sum((t * torch.linspace(-1, 1, t.numel(), device="cuda")).sum() for t in (*a_w, *b_w, *biases))
            loss.backward()
    run()                                                             # (a)
    ref = {n: p.grad.clone() for n, p in hn.named_parameters()}
    for p in hn.parameters():
        p.grad = None
    # (b) flat-bucket views + fused in-place accumulation
    params = list(hn.generators.parameters()) + [p for n, p in hn.named_parameters() if not n.startswith("generators")]
    sync = GradSync(params, bucket_bytes=256 << 20)
    sync.zero_grad()
    hn.fuse_generator_grad_accumulation = True
    hn.grad_ready_callback = sync.notify
    run()
    sync.finish()
    for n, p in hn.named_parameters():
        assert rel(p.grad, ref[n]) < F32, n
    assert all(sync._touched)                                         # every bucket was seen, including the in-place generator buckets
    # (c) factor mode on both generators
    sync.zero_grad()
    hn.fuse_generator_grad_accumulation = False
    hn.grad_ready_callback = None
    hn.factor_sinks = {l: Rank1FactorSync(gen.weight.shape[0], 768, "cuda", max_terms=4) for l, gen in enumerate(hn.generators)}
    run()
    for l, gen in enumerate(hn.generators):
        assert gen.weight.grad.abs().max().item() == 0.0              # nothing dense was written during the micro-steps
        hn.factor_sinks[l].apply_(gen.weight.grad, gen.bias.grad)
    for n, p in hn.named_parameters():
        assert rel(p.grad, ref[n]) < F32, n
    sync.remove()
