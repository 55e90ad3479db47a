"""Pins oracle/oracle.py against vectors produced by RUNNING THE REFERENCE (oracle/make_golden.py).

fp32 tolerance: 1e-5 relative (north_star) on outputs and gradients; index work bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O

RTOL = 1e-5


def load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def sd_of(d):
    return {k[3:]: v for k, v in d.items() if k.startswith("sd/")}


def rel_err(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def assert_close(a, b, what, rtol=RTOL):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if b.double().norm() < 1e-7:     # mathematically-zero gradients (e.g. d/d k.bias: softmax shift invariance)
        assert a.double().norm() < 1e-6, f"{what}: expected ~0, got norm {a.double().norm():.3e}"
        return
    e = rel_err(a.double(), b.double())
    assert e < rtol, f"{what}: rel err {e:.3e}"


def hyper_kwargs(meta):
    D_hyp, D_mm, H, r, alpha, n_tokens, K, B, prune = [int(v) for v in meta]
    return dict(n_tokens=n_tokens, rank=r, alpha=float(alpha), lm_dim=H, mm_dim=D_mm), prune


@pytest.mark.parametrize("name", ["hypernet_h1_full_ctx", "hypernet_h1_masked", "hypernet_h1_pruned",
                                  "hypernet_h1_dropout"])
def test_hypernet_wrapper_matches_reference(golden_dir, name):
    d = load(golden_dir, name)
    kw, prune = hyper_kwargs(d["meta"])
    sd = sd_of(d)
    params = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "pos_encs" not in k) for k, v in sd.items()}
    keep = d.get("keep_mask")
    out = O.hypernet_wrapper_forward(params, d["x"], d["z"], keep_mask=keep, **kw)
    assert_close(out.detach(), d["out"], name + " out")
    (out * d["dy"]).sum().backward()
    # z assembled by the oracle's process_embeddings equals the reference's z bit-exactly
    _, z2 = O.process_embeddings(d["x"], (d["m"], d["t"], d["p"]), None,
                                 prune=None if prune < 0 else prune, finetune_mm_dim=int(d["meta"][0]))
    assert torch.equal(z2, d["z"])
    none_grads = []
    for k, v in d.items():
        if not k.startswith("grad/"):
            continue
        p = params[k[5:]]
        if v.numel() == 0:
            none_grads.append(k[5:])
            assert p.grad is None or p.grad.abs().max() == 0, k
            continue
        assert_close(p.grad, v, k)
    # generators.1.* and projector.net.3.* never receive a gradient
    assert sorted(none_grads) == ["hypernet.generators.1.bias", "hypernet.generators.1.weight",
                                  "projector.net.3.bias", "projector.net.3.weight"]
    for i in range(2):
        if f"adapter/a{i}" in d:
            a_w, b_w, biases = O.hypernetwork_forward(sd, d["z"], **kw)
            assert_close(a_w[i], d[f"adapter/a{i}"], f"A{i}")
            assert_close(b_w[i], d[f"adapter/b{i}"], f"B{i}")
            assert_close(biases[i], d[f"adapter/bias{i}"], f"bias{i}")


def test_pe_buffer_matches_reference(golden_dir):
    d = load(golden_dir, "hypernet_h1_full_ctx")
    pe = d["sd/hypernet.pos_encs.pe"]
    assert_close(O.sinusoidal_pe(pe.shape[2], pe.shape[1]), pe, "pe", rtol=1e-6)


def test_fewshot_merged_projector(golden_dir):
    d = load(golden_dir, "fewshot_merged")
    D, _, H, r, alpha, n_tokens, K, B, N = [int(v) for v in d["meta"]]
    sd = sd_of(d)
    kw = dict(n_tokens=n_tokens, rank=r, alpha=float(alpha), lm_dim=H, mm_dim=D)
    adapters = [O.hypernetwork_forward(sd, z, **kw) for z in d["zs"]]
    a, b, bias = O.average_adapters(adapters)
    merged = O.combine_lora(sd, a, b, bias)
    for k in ("0.weight", "0.bias", "3.weight", "3.bias"):
        assert_close(merged[k], sd["generated_projector." + k], "merged " + k)
    leaves = {k: v.clone().requires_grad_(True) for k, v in merged.items()}
    out = O.merged_forward(leaves, d["x"])
    assert_close(out.detach(), d["out"], "merged out")
    (out * d["dy"]).sum().backward()
    for k in leaves:
        assert_close(leaves[k].grad, d["grad/" + k], "grad " + k)
    # H2: merged MLP == un-merged full adapted MLP
    y2 = O.adapted_mlp_full(sd["projector.net.0.weight"], sd["projector.net.0.bias"],
                            sd["projector.net.3.weight"], sd["projector.net.3.bias"], d["x"], a, b, bias)
    assert_close(y2, d["out"], "full adapted == merged")


def test_projector_mlp2_with_dropout(golden_dir):
    d = load(golden_dir, "projector_mlp2")
    sd = {k: v.clone().requires_grad_(True) for k, v in sd_of(d).items()}
    out = O.projector_forward(sd, d["x"], drop_keep=d["keep"], p_drop=0.1)
    assert_close(out.detach(), d["out"], "train out")
    (out * d["dy"]).sum().backward()
    for k, v in d.items():
        if k.startswith("grad/"):
            assert_close(sd["projector." + k[5:]].grad, v, k)
    assert_close(O.projector_forward(sd_of(d), d["x"]), d["out_eval"], "eval out")


def test_lora_only_forward(golden_dir):
    d = load(golden_dir, "lora_full")
    D, H, r, alpha, B = [int(v) for v in d["meta"]]
    sd = sd_of(d)
    loras = [(sd[f"lora_adapters.loras.{i}.A"].clone().requires_grad_(True),
              sd[f"lora_adapters.loras.{i}.B"].clone().requires_grad_(True)) for i in range(2)]
    out = O.only_lora_forward(sd, d["x"], loras, alpha=alpha, rank=r)
    assert_close(out.detach(), d["out"], "lora out")
    (out * d["dy"]).sum().backward()
    for i in range(2):
        assert_close(loras[i][0].grad, d[f"grad/lora_adapters.loras.{i}.A"], f"dA{i}")
        assert_close(loras[i][1].grad, d[f"grad/lora_adapters.loras.{i}.B"], f"dB{i}")


def test_splice_bit_exact(golden_dir):
    d = load(golden_dir, "splice")
    table = d["table"].to(torch.bfloat16)
    emb, mask, lab = O.splice_prefix(d["projected"], table, d["ids"], d["attn"], d["labels"])
    assert bool(d["embeds_is_fp32"]) and emb.dtype == torch.float32      # torch.cat promotion
    assert torch.equal(emb, d["embeds"])
    assert torch.equal(lab, d["labels_out"])
    assert lab[:, 0].eq(-100).all() and mask[:, 0].eq(1).all()


def test_rotation_restatement_bit_exact(golden_dir):
    z = np.load(os.path.join(golden_dir, "rotation.npz"))
    for key in z.files:
        dim = int(key[1:])
        R = O.ortho_group_rvs(dim, np.random.RandomState(42))
        assert np.array_equal(R, z[key]), key
        np.testing.assert_allclose(R @ R.T, np.eye(dim), atol=1e-12)


def test_optimizer_restatement_matches_torch_adamw_and_clip(golden_dir):
    """oracle.clip_grad_norm + oracle.adamw_step against the fixture produced by the reference's own call sequence
    (torch.nn.utils.clip_grad_norm_ then optim.AdamW.step(), train_hypernet.py:148-149) over 3 steps."""
    d = load(golden_dir, "optimizer_adamw_clip")
    n, steps, skip = int(d["n_params"]), int(d["steps"]), int(d["no_grad_index"])
    hp = dict(lr=float(d["lr"]), betas=(float(d["beta1"]), float(d["beta2"])), eps=float(d["eps"]), weight_decay=float(d["weight_decay"]))
    idx = [i for i in range(n) if i != skip]
    p = {i: d[f"p0/{i}"] for i in range(n)}
    m = {i: torch.zeros_like(p[i]) for i in idx}
    v = {i: torch.zeros_like(p[i]) for i in idx}
    for t in range(steps):
        clipped, total = O.clip_grad_norm([d[f"g{t}/{i}"] for i in idx], float(d["max_grad_norm"]))
        assert torch.allclose(total, d[f"norm{t}"], rtol=1e-6)
        for i, g in zip(idx, clipped):
            assert torch.allclose(g, d[f"gclip{t}/{i}"], rtol=1e-6, atol=1e-12)
            p[i], m[i], v[i] = O.adamw_step(p[i], g, m[i], v[i], t + 1, **hp)
            assert torch.allclose(p[i], d[f"p{t + 1}/{i}"], rtol=1e-6, atol=1e-7), (t, i)
            assert torch.allclose(m[i], d[f"m{t + 1}/{i}"], rtol=1e-5, atol=1e-9)
            assert torch.allclose(v[i], d[f"v{t + 1}/{i}"], rtol=1e-5, atol=1e-12)
        assert torch.equal(p[skip], d[f"p{t + 1}/{skip}"])            # no gradient -> untouched (no weight decay either)
