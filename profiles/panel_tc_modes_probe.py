"""ROUND-2 FIRST RUN.  Single-mode launches of the tcgen05 panel kernel (project-only, reduce-only; written after round 1's GPU budget
was spent, never run on a GPU yet) against the kernels they would replace, then the bench step with fused_panel = -1 (default: tcgen05
dpre pass) / 6 (+ tcgen05 reductions) / 14 (+ tcgen05 projections) / 30 (+ the fp32-input dY pass).  Correctness is checked first; timings only print if it holds.
  gpurun: DMI_EXPERIMENTAL=1 python -m pytest tests/test_panel_gpu.py -x -q && python profiles/panel_tc_modes_probe.py
  ncu   : ncu --set full --import-source on --clock-control none -k regex:panel_tc -c 3 -o gpurun_out/panel_tc python profiles/panel_tc_modes_probe.py
"""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sample-efficient-multimodality_b200"))
import torch
from dmi_b200 import ops
dev, bf = "cuda", torch.bfloat16
B, D, H, r = int(os.environ.get("ROWS", 32768)), 768, 2048, 32
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)
z = lambda *s: torch.zeros(*s, device=dev)


def timeit(fn, reps=30, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


ok = True
for (M, K) in [(128, 2048), (1000, 768), (32768, 2048), (32768, 768)]:
    inp = (rn(M, K) / 8).to(bf)
    W = (rn(r, K) / math.sqrt(K)).to(bf)
    L = rn(M, r).to(bf)
    out = torch.empty(M, r, device=dev, dtype=bf)
    ops.panel_tc_project(inp, W, out)
    e0 = rel(out, inp.float() @ W.float().t())
    G, Gt, cs = z(r, K), z(K, r), z(K)
    ops.panel_tc_reduce(L, inp, G, colsum=cs)
    ops.panel_tc_reduce(L, inp, Gt, transpose_out=True)
    ref = L.float().t() @ inp.float()
    e1, e2, e3 = rel(G, ref), rel(Gt, ref.t()), rel(cs, inp.float().sum(0))
    good = e0 < 6e-3 and max(e1, e2, e3) < 1e-4
    ok = ok and good
    print(f"check M={M:6d} K={K:5d}: project {e0:.2e}  reduce {e1:.2e}  reduce^T {e2:.2e}  colsum {e3:.2e}  {'OK' if good else 'MISMATCH'}", flush=True)
for (M, K) in [(1000, 2048), (32768, 2048)]:          # merged-column-sum variant of the fused bf16 pass
    inp = (rn(M, K) / 8).to(bf)
    W = (rn(r, K) / math.sqrt(K)).to(bf)
    L = rn(M, r).to(bf)
    out = torch.empty(M, r, device=dev, dtype=bf)
    G, cs = z(r, K), z(K)
    ops.panel_fused_tc(inp, W, L, out, G, colsum=cs, merged_colsum=True)
    e = (rel(out, inp.float() @ W.float().t()), rel(G, L.float().t() @ inp.float()), rel(cs, inp.float().sum(0)))
    good = e[0] < 6e-3 and max(e[1:]) < 1e-4
    ok = ok and good
    print(f"check mcs  M={M:6d} K={K:5d}: out {e[0]:.2e}  G {e[1]:.2e}  colsum {e[2]:.2e}  {'OK' if good else 'MISMATCH'}", flush=True)
for (M, K) in [(1000, 2048), (32768, 2048)]:          # fp32-input form (the dY pass)
    inp = rn(M, K) / 8
    W = (rn(r, K) / math.sqrt(K)).to(bf)
    L = rn(M, r).to(bf)
    out = torch.empty(M, r, device=dev, dtype=bf)
    cp = torch.empty(M, K, device=dev, dtype=bf)
    G, cs = z(r, K), z(K)
    ops.panel_fused_tc32(inp, W, L, out, G, colsum=cs, copy=cp)
    xb_ = inp.to(bf).float()
    e = (rel(out, xb_ @ W.float().t()), rel(G, L.float().t() @ xb_), rel(cs, xb_.sum(0)))
    good = e[0] < 6e-3 and max(e[1:]) < 1e-4 and torch.equal(cp, inp.to(bf))
    ok = ok and good
    print(f"check fp32 M={M:6d} K={K:5d}: out {e[0]:.2e}  G {e[1]:.2e}  colsum {e[2]:.2e}  copy_exact={torch.equal(cp, inp.to(bf))}  {'OK' if good else 'MISMATCH'}", flush=True)
print("ALL OK" if ok else "FAILED", flush=True)
if not ok:
    sys.exit(1)

hb = [(rn(B, H) / 8).to(bf) for _ in range(3)]
xb = [(rn(B, D) / 8).to(bf) for _ in range(3)]
Wh, Wd = (rn(r, H) / math.sqrt(H)).to(bf), (rn(r, D) / math.sqrt(D)).to(bf)
Lp = rn(B, r).to(bf)
out = torch.empty(B, r, device=dev, dtype=bf)
hf = [rn(B, H) / 8 for _ in range(3)]
cpb = torch.empty(B, H, device=dev, dtype=bf)
G, Gt, Gd, cs = z(r, H), z(H, r), z(D, r), z(H)
for name, fn, nbytes in [
    ("project [B,2048]: skinny_rows", lambda i: ops.skinny_rows(hb[i % 3], Wh, out), B * H * 2),
    ("project [B,2048]: tcgen05", lambda i: ops.panel_tc_project(hb[i % 3], Wh, out), B * H * 2),
    ("project [B,768]: skinny_rows", lambda i: ops.skinny_rows(xb[i % 3], Wd, out), B * D * 2),
    ("project [B,768]: tcgen05", lambda i: ops.panel_tc_project(xb[i % 3], Wd, out), B * D * 2),
    ("reduce+colsum [B,2048]: outer_reduce", lambda i: ops.outer_reduce(Lp, hb[i % 3], G, colsum=cs), B * H * 2),
    ("reduce+colsum [B,2048]: tcgen05", lambda i: ops.panel_tc_reduce(Lp, hb[i % 3], G, colsum=cs), B * H * 2),
    ("reduce^T [B,2048]: outer_reduce", lambda i: ops.outer_reduce(Lp, hb[i % 3], Gt, transpose_out=True), B * H * 2),
    ("reduce^T [B,2048]: tcgen05", lambda i: ops.panel_tc_reduce(Lp, hb[i % 3], Gt, transpose_out=True), B * H * 2),
    ("reduce^T [B,768]: outer_reduce", lambda i: ops.outer_reduce(Lp, xb[i % 3], Gd, transpose_out=True), B * D * 2),
    ("reduce^T [B,768]: tcgen05", lambda i: ops.panel_tc_reduce(Lp, xb[i % 3], Gd, transpose_out=True), B * D * 2),
    ("fused dpre pass: tcgen05 (round-1 default)", lambda i: ops.panel_fused_tc(hb[i % 3], Wh, Lp, out, G, colsum=cs), B * H * 2),
    ("fused dpre pass: tcgen05, merged colsum", lambda i: ops.panel_fused_tc(hb[i % 3], Wh, Lp, out, G, colsum=cs, merged_colsum=True), B * H * 2),
    ("fp32 dY pass: skinny_rows + outer_reduce", lambda i: (ops.skinny_rows(hf[i % 3], Wh, out, copy=cpb), ops.outer_reduce(Lp, cpb, G, colsum=cs)), B * H * 6),
    ("fp32 dY pass: tcgen05", lambda i: ops.panel_fused_tc32(hf[i % 3], Wh, Lp, out, G, colsum=cs, copy=cpb), B * H * 6),
]:
    ms = timeit(fn)
    print(f"{name:40s}: {ms*1e3:7.1f} us   {nbytes/ms/1e9:5.2f} TB/s", flush=True)

w1, w2 = rn(H, D) / math.sqrt(D), rn(H, H) / math.sqrt(H)
b1, b2 = z(H), z(H)
A0, B0, A1, B1 = rn(D * r) / math.sqrt(D), rn(r * H) * 0.1, rn(H * r) / math.sqrt(H), rn(r * H) * 0.1
be0, be1 = z(H), z(H)
xs = [rn(B, D) for _ in range(3)]
dys = [rn(B, H) / math.sqrt(H) for _ in range(3)]
y = torch.empty(B, H, device=dev)
grads = dict(dA0=z(D, r), dB0=z(r, H), dbeta0=z(H), dA1=z(H, r), dB1=z(r, H), dbeta1=z(H))
pk = ops.PackedProjector(D, H, r, dev)
pk.pack_base(w1, w2)
st = ops.MlpStash(B, D, H, r, dev, full=True)


def step(i):
    pk.pack_adapter(A0, B0, be0, A1, B1, be1, b1, b2)
    ops.adapted_mlp_fwd(pk, st, xs[i % 3], y)
    ops.adapted_mlp_bwd(pk, st, dys[i % 3], grads)


F = 2 * D * H + 4 * H * H + 4 * r * D + 18 * r * H
res = {}
for name, opt in (("separate passes", 0), ("default (tcgen05 dpre pass)", -1), ("+ tcgen05 reductions", 6), ("+ tcgen05 projections", 14), ("+ tcgen05 fp32 dY pass", 30), ("+ merged colsum in the dpre pass", 62), ("separate passes (again)", 0)):
    ops.set_option("fused_panel", opt)
    for k in grads:
        grads[k].zero_()
    step(0)
    res[opt] = {k: v.clone() for k, v in grads.items()}
    ms = timeit(step, reps=60, warm=5)
    print(f"step, {name:30s}: {ms*1e3:8.1f} us/step  {B/ms/1e3:6.2f} M samples/s  {B*F/ms/1e9:6.0f} TFLOP/s", flush=True)
for opt in (-1, 6, 14, 30, 62):
    print(f"fused_panel={opt:3d} vs separate: " + "  ".join(f"{k} {rel(res[opt][k], res[0][k]):.1e}" for k in res[0]))
ops.set_option("fused_panel", -1)
