"""Fused projection + batch-reduction pass (panel.cu) against the skinny_rows + outer_reduce pair it replaces:
results (vs the two kernels and vs torch fp32 on the bf16-rounded input), time per pass, and the bench step with the option on/off.
CUDA events, rotating inputs larger than L2."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sample-efficient-multimodality_b200"))
import torch
from dmi_b200 import ops
dev = "cuda"
B, D, H, r = int(os.environ.get("ROWS", 32768)), 768, 2048, 32
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)
z = lambda *s: torch.zeros(*s, device=dev)
bf = torch.bfloat16


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


def check(M, K, R, f32, ld_pad=0):
    base = rn(M, K + ld_pad) / 8                    # bf16 pad must keep 16-byte rows: ld_pad % 8 == 0 (bf16) / % 4 == 0 (fp32)
    inp = (base if f32 else base.to(bf))[:, :K]     # row stride K + ld_pad
    W = (rn(R, K) / math.sqrt(K)).to(bf)
    L = rn(M, R).to(bf)
    out = torch.full((M, R), 7.0, device=dev, dtype=bf)
    G, cs = z(R, K), z(K)
    copy = torch.empty(M, K, device=dev, dtype=bf) if f32 else None
    ops.panel_fused(inp, W, L, out, G, colsum=cs, copy=copy, scale=0.5)
    torch.cuda.synchronize()
    xb = inp.to(bf).float()
    ref_out = xb @ W.float().t()
    ref_G = 0.5 * (L.float().t() @ xb)
    ref_cs = 0.5 * xb.sum(0)
    e = (rel(out, ref_out), rel(G, ref_G), rel(cs, ref_cs))
    ok = e[0] < 6e-3 and e[1] < 1e-4 and e[2] < 1e-4
    if f32:
        ok = ok and torch.equal(copy, inp.to(bf))
    print(f"check M={M:6d} K={K} R={R} f32={int(f32)} pad={ld_pad}: out {e[0]:.2e}  G {e[1]:.2e}  colsum {e[2]:.2e}  {'OK' if ok else 'MISMATCH'}", flush=True)
    return ok


ok = True
for (M, K, R, f32, pad) in [(64, 2048, 32, False, 0), (1000, 2048, 32, False, 8), (1000, 2048, 32, True, 4), (4097, 1024, 16, False, 0),
                            (4097, 1024, 16, True, 0), (32768, 2048, 32, False, 0), (32768, 2048, 32, True, 0), (9000, 2048, 16, True, 0)]:
    ok = check(M, K, R, f32, pad) and ok
print("ALL OK" if ok else "FAILED", flush=True)

# ---- time per pass -------------------------------------------------------------------------------------------------------
hb = [(rn(B, H) / 8).to(bf) for _ in range(3)]
hf = [rn(B, H) / 8 for _ in range(3)]
Wr = (rn(r, H) / math.sqrt(H)).to(bf)
Lp = rn(B, r).to(bf)
out = torch.empty(B, r, device=dev, dtype=bf)
cp = torch.empty(B, H, device=dev, dtype=bf)
G, cs = z(r, H), z(H)
rows = [
    ("bf16 pass: skinny_rows + outer_reduce", lambda i: (ops.skinny_rows(hb[i % 3], Wr, out), ops.outer_reduce(Lp, hb[i % 3], G, colsum=cs)), B * H * 2),
    ("bf16 pass: fused", lambda i: ops.panel_fused(hb[i % 3], Wr, Lp, out, G, colsum=cs), B * H * 2),
    ("fp32 pass: skinny_rows + outer_reduce", lambda i: (ops.skinny_rows(hf[i % 3], Wr, out, copy=cp), ops.outer_reduce(Lp, cp, G, colsum=cs)), B * H * 6),
    ("fp32 pass: fused", lambda i: ops.panel_fused(hf[i % 3], Wr, Lp, out, G, colsum=cs, copy=cp), B * H * 6),
]
for name, fn, nbytes in rows:
    ms = timeit(fn)
    print(f"{name:42s}: {ms*1e3:7.1f} us   {nbytes/ms/1e9:5.2f} TB/s (algorithmic bytes of one sweep)", flush=True)

# ---- the bench step with the option off / on ----------------------------------------------------------------------------
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
for name, opt in (("step, separate passes", 0), ("step, fused passes", 1), ("step, separate passes (again)", 0), ("step, fused passes (again)", 1)):
    ops.set_option("fused_panel", opt)
    for k in grads:
        grads[k].zero_()
    step(0)
    res[opt] = {k: v.clone() for k, v in grads.items()}
    ms = timeit(step, reps=60, warm=5)
    print(f"{name:34s}: {ms*1e3:8.1f} us/step  {B/ms/1e3:6.2f} M samples/s  {B*F/ms/1e9:6.0f} TFLOP/s", flush=True)
for k in res[0]:
    print(f"grad {k:7s} fused vs separate: rel {rel(res[1][k], res[0][k]):.2e}")
ops.set_option("fused_panel", -1)
