"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (build container only).

Usage:  python oracle/make_golden.py           (needs /root/reference; writes tests/golden/)

The reference has no tests or golden vectors, so the oracle is pinned against outputs of the
reference's own modules: this script imports ``dmi.model.{projector,hypernet,lora,mmmodel}`` from
``/root/reference`` (import recipe from SURVEY.md section 8c), runs them on seeded CPU inputs at reduced
widths (so the fixtures stay small) and stores the state-dict, the inputs, the outputs and the
autograd gradients.  ``/root/reference`` does not exist on the GPU box; the committed ``.npz`` files do.
Test infrastructure only.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference():
    import transformers  # noqa: F401  (must be imported before timm is stubbed)
    sys.modules.setdefault("timm", types.ModuleType("timm"))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import dmi.model.hypernet as hn
    import dmi.model.lora as lora
    import dmi.model.mmmodel as mm
    import dmi.model.projector as proj
    import dmi.utils.args as args
    return proj, hn, lora, mm, args


def npify(d):
    out = {}
    for k, v in d.items():
        if isinstance(v, torch.Tensor):
            v = v.detach()
            out[k] = v.to(torch.float32).numpy() if v.dtype == torch.bfloat16 else v.numpy()
        else:
            out[k] = np.asarray(v)
    return out


def unit_rows(g, n, d):
    x = torch.randn(n, d, generator=g)
    return x / x.norm(dim=1, keepdim=True)


def build_wrapper(proj, hn, args, *, D_hyp, D_mm, H, r, alpha, n_tokens, prune=None, seed=0):
    torch.manual_seed(seed)
    pa = args.ProjectorArgs(proj_dropout=0.1)
    base = proj.Projector(pa, H, D_hyp, "cpu")
    # make the biases non-trivial
    tmp = tempfile.NamedTemporaryFile(suffix=".pt", delete=False)
    torch.save({"projector_state_dict": base.state_dict()}, tmp.name)
    pa2 = args.ProjectorArgs(proj_name_or_path=tmp.name, proj_dropout=0.1, proj_prune=prune)
    ha = args.HypnetArgs(hn_arch="attention", hn_hypnet_dim=D_hyp, hn_rank=r, hn_alpha=alpha,
                         hn_predict_bias=True, hn_n_proj_layers=2, hn_use_pos_encs=True)
    w = hn.HyperNetWrapper(ha, pa2, H, D_mm, n_tokens, "cpu")
    # the reference zero-initialises the generator bias; randomise it so the bias path is exercised
    with torch.no_grad():
        for gen in w.hypernet.generators:
            gen.bias.normal_(0, 0.02)
    os.unlink(tmp.name)
    return w


def case_hypernet(proj, hn, args, name, *, D_hyp=64, D_mm=64, H=128, r=4, alpha=8, n_tokens=6, K=6, B=5,
                  prune=None, train_dropout=False, seed=1):
    """HyperNetWrapper.forward(x, z) as written (H1) + grads of sum(out*dy) wrt all hypernet params."""
    w = build_wrapper(proj, hn, args, D_hyp=D_hyp, D_mm=D_mm, H=H, r=r, alpha=alpha, n_tokens=n_tokens,
                      prune=prune, seed=seed)
    g = torch.Generator().manual_seed(seed + 100)
    x = unit_rows(g, B, D_mm)
    m = unit_rows(g, K, D_mm)
    t = unit_rows(g, K, D_hyp)
    p = unit_rows(g, 1, D_hyp)
    if prune is not None:
        m_pad = torch.nn.functional.pad(m, (0, D_hyp - prune, 0, 0))
    else:
        m_pad = m
    z = torch.cat([p, torch.stack((m_pad, t), 0).transpose(0, 1).reshape(-1, D_hyp)], 0)
    dy = torch.randn(B, H, generator=g) / H ** 0.5
    extra = {}
    if train_dropout:
        w.train()
        captured = {}

        def hook(mod, inp, out):
            captured["keep"] = (out != 0).to(torch.float32)
        h = w.hypernet.hypnet.dropout.register_forward_hook(hook)
        torch.manual_seed(seed + 7)
        out = w(x, z)
        h.remove()
        extra["keep_mask"] = captured["keep"][0]          # [heads, S, S]
    else:
        w.eval()
        out = w(x, z)
    (out * dy).sum().backward()
    a_w, b_w, biases = [], [], []
    with torch.no_grad():
        if not train_dropout:
            a_w, b_w, biases = w.hypernet(z)
    d = {"x": x, "m": m, "t": t, "p": p, "z": z, "dy": dy, "out": out,
         "meta": np.array([D_hyp, D_mm, H, r, alpha, n_tokens, K, B, -1 if prune is None else prune])}
    for k, v in w.state_dict().items():
        d["sd/" + k] = v
    for k, v in w.named_parameters():
        d["grad/" + k] = v.grad if v.grad is not None else torch.zeros(0)
    for i, (a, b, c) in enumerate(zip(a_w, b_w, biases)):
        d[f"adapter/a{i}"], d[f"adapter/b{i}"], d[f"adapter/bias{i}"] = a, b, c
    d.update(extra)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **npify(d))
    return w, x, z


def case_fewshot(proj, hn, args, name, *, D=64, H=128, r=4, alpha=8, n_tokens=6, K=3, N=3, B=7, seed=3):
    """generate_projector_from_multiple_adapters (mask path: K < n_tokens) -> merged MLP2 fwd + grads."""
    w = build_wrapper(proj, hn, args, D_hyp=D, D_mm=D, H=H, r=r, alpha=alpha, n_tokens=n_tokens, seed=seed)
    w.eval()
    g = torch.Generator().manual_seed(seed + 100)
    zs = []
    for _ in range(N):
        m, t, p = unit_rows(g, K, D), unit_rows(g, K, D), unit_rows(g, 1, D)
        zs.append(torch.cat([p, torch.stack((m, t), 0).transpose(0, 1).reshape(-1, D)], 0))
    w.generate_projector_from_multiple_adapters(zs)
    x = unit_rows(g, B, D)
    dy = torch.randn(B, H, generator=g) / H ** 0.5
    out = w(x, None)
    (out * dy).sum().backward()
    d = {"x": x, "dy": dy, "out": out, "zs": torch.stack(zs),
         "meta": np.array([D, D, H, r, alpha, n_tokens, K, B, N])}
    for k, v in w.state_dict().items():
        d["sd/" + k] = v
    for k, v in w.generated_projector.named_parameters():
        d["grad/" + k] = v.grad
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **npify(d))


def case_projector(proj, args, name, *, D=64, H=128, B=9, seed=5):
    """Projector.forward with dropout ACTIVE (train_projector path) + grads; keep-mask captured."""
    torch.manual_seed(seed)
    p = proj.Projector(args.ProjectorArgs(proj_dropout=0.1), H, D, "cpu")
    p.train()
    g = torch.Generator().manual_seed(seed + 100)
    x = unit_rows(g, B, D)
    dy = torch.randn(B, H, generator=g) / H ** 0.5
    cap = {}
    hk = p.net[2].register_forward_hook(lambda m, i, o: cap.__setitem__("keep", (o != 0) | (i[0] == 0)))
    torch.manual_seed(seed + 7)
    out = p(x)
    hk.remove()
    (out * dy).sum().backward()
    p.eval()
    out_eval = p(x)
    d = {"x": x, "dy": dy, "out": out, "out_eval": out_eval, "keep": cap["keep"].to(torch.float32),
         "meta": np.array([D, H, B])}
    for k, v in p.state_dict().items():
        d["sd/projector." + k] = v
    for k, v in p.named_parameters():
        d["grad/" + k] = v.grad
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **npify(d))


def case_lora(proj, lora, args, name, *, D=64, H=128, r=4, alpha=8, B=6, seed=9):
    """LoraWrapper.forward (only_lora_forward) + grads to A,B (B randomised: reference inits B=0)."""
    torch.manual_seed(seed)
    base = proj.Projector(args.ProjectorArgs(), H, D, "cpu")
    tmp = tempfile.NamedTemporaryFile(suffix=".pt", delete=False)
    torch.save({"projector_state_dict": base.state_dict()}, tmp.name)
    la = args.LoraArgs(lora_rank=r, lora_alpha=alpha, lora_n_proj_layers=2)
    w = lora.LoraWrapper(la, args.ProjectorArgs(proj_name_or_path=tmp.name), H, D, "cpu")
    os.unlink(tmp.name)
    with torch.no_grad():
        for l in w.lora_adapters.loras:
            l.B.normal_(0, 0.05)
    w.train()
    g = torch.Generator().manual_seed(seed + 100)
    x = unit_rows(g, B, D)
    dy = torch.randn(B, H, generator=g) / H ** 0.5
    out = w(x)
    (out * dy).sum().backward()
    d = {"x": x, "dy": dy, "out": out, "meta": np.array([D, H, r, alpha, B])}
    for k, v in w.state_dict().items():
        d["sd/" + k] = v
    for k, v in w.named_parameters():
        d["grad/" + k] = v.grad if v.grad is not None else torch.zeros(0)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **npify(d))


def case_splice(mm, name, *, B=3, T=5, H=16, V=11, seed=11):
    """mmmodel.py:36-48 executed through HypernetMMModel.forward with a stub LLM that records its inputs."""
    g = torch.Generator().manual_seed(seed)
    table = torch.randn(V, H, generator=g).to(torch.bfloat16)
    rec = {}

    class StubLLM(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.emb = torch.nn.Embedding(V, H).to(torch.bfloat16)
            with torch.no_grad():
                self.emb.weight.copy_(table)
            self.config = types.SimpleNamespace(hidden_size=H)

        def get_input_embeddings(self):
            return self.emb

        def forward(self, inputs_embeds=None, labels=None):
            rec["embeds"], rec["labels"] = inputs_embeds, labels
            return types.SimpleNamespace(loss=inputs_embeds.float().sum() * 0)

    proj_out = torch.randn(B, H, generator=g)

    class StubHyper(torch.nn.Module):
        def forward(self, x, z):
            return proj_out

    model = mm.HypernetMMModel(StubLLM(), StubHyper(), "cpu", H, "t", 0)
    ids = torch.randint(0, V, (B, T), generator=g)
    am = torch.ones(B, T)
    am[:, -2:] = 0
    labels = torch.randint(0, V, (B, T), generator=g)
    model(torch.zeros(B, H), None, ids, am, labels)
    d = {"projected": proj_out, "table": table, "ids": ids, "attn": am, "labels": labels,
         "embeds": rec["embeds"], "labels_out": rec["labels"],
         "embeds_is_fp32": np.array(rec["embeds"].dtype == torch.float32)}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **npify(d))


def case_rotation(name, dims=(8, 48), seed=42):
    """scipy.stats.ortho_group.rvs with a seeded RandomState, as _get_rotation_matrix draws it."""
    from scipy.stats import ortho_group
    d = {}
    for dim in dims:
        rs = np.random.RandomState(seed)
        d[f"R{dim}"] = ortho_group.rvs(dim, random_state=rs)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)


def case_optimizer(name, *, steps=3, seed=13):
    """The optimizer step exactly as the reference trainer issues it (train_hypernet.py:148-149 with the optimizer of :526-532 and
    the v4 config's hyper-parameters): torch.nn.utils.clip_grad_norm_(params, max_grad_norm) then optim.AdamW.step().
    One parameter never receives a gradient (generators.1.* in the reference's H1 training) and must stay untouched."""
    g = torch.Generator().manual_seed(seed)
    shapes = [(13, 7), (5,), (129, 33), (2, 96), (1031,), (6, 6)]
    hp = dict(lr=1e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=5e-6)
    max_grad_norm = 1.0
    params = [torch.nn.Parameter(torch.randn(*s, generator=g)) for s in shapes]
    opt = torch.optim.AdamW(params=params, **hp)
    out = {"n_params": torch.tensor(len(shapes)), "steps": torch.tensor(steps), "max_grad_norm": torch.tensor(max_grad_norm),
           "lr": torch.tensor(hp["lr"]), "beta1": torch.tensor(hp["betas"][0]), "beta2": torch.tensor(hp["betas"][1]),
           "eps": torch.tensor(hp["eps"]), "weight_decay": torch.tensor(hp["weight_decay"]), "no_grad_index": torch.tensor(5)}
    for i, p_ in enumerate(params):
        out[f"p0/{i}"] = p_.detach().clone()
    for t in range(steps):
        scale = [3.0, 0.02, 1.0][t % 3]                      # step 0 clips hard, step 1 does not clip, step 2 clips
        for i, p_ in enumerate(params):
            if i == 5:
                p_.grad = None
                continue
            p_.grad = torch.randn(*shapes[i], generator=g) * scale
            out[f"g{t}/{i}"] = p_.grad.detach().clone()
        total = torch.nn.utils.clip_grad_norm_(params, max_grad_norm)
        opt.step()
        out[f"norm{t}"] = total.detach().clone()
        for i, p_ in enumerate(params):
            out[f"p{t + 1}/{i}"] = p_.detach().clone()
            if i != 5:
                out[f"gclip{t}/{i}"] = p_.grad.detach().clone()
                out[f"m{t + 1}/{i}"] = opt.state[p_]["exp_avg"].detach().clone()
                out[f"v{t + 1}/{i}"] = opt.state[p_]["exp_avg_sq"].detach().clone()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **npify(out))


def main():
    os.makedirs(OUT, exist_ok=True)
    proj, hn, lora, mm, args = import_reference()
    torch.set_num_threads(1)
    case_hypernet(proj, hn, args, "hypernet_h1_full_ctx")                                    # K == n_tokens
    case_hypernet(proj, hn, args, "hypernet_h1_masked", K=3, seed=2)                         # key-mask path
    case_hypernet(proj, hn, args, "hypernet_h1_pruned", D_mm=40, prune=40, seed=4)           # A0 truncation
    case_hypernet(proj, hn, args, "hypernet_h1_dropout", train_dropout=True, seed=6)         # attention dropout
    case_fewshot(proj, hn, args, "fewshot_merged")
    case_projector(proj, args, "projector_mlp2")
    case_lora(proj, lora, args, "lora_full")
    case_splice(mm, "splice")
    case_rotation("rotation")
    case_optimizer("optimizer_adamw_clip")
    print("golden vectors written to", os.path.abspath(OUT))
    for f in sorted(os.listdir(OUT)):
        print(f"  {f}: {os.path.getsize(os.path.join(OUT, f))} bytes")


if __name__ == "__main__":
    main()
