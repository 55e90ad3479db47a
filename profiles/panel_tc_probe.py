"""tcgen05 fused projection + batch-reduction pass (panel_tc.cu): results vs torch, time vs skinny_rows + outer_reduce and vs the
mma.sync fused kernel, and the bench step with fused_panel = 0 / 2.  CUDA events, rotating inputs larger than L2."""
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


def check(M, K, pad=0, colsum=True):
    R = 32
    base = (rn(M, K + pad) / 8).to(bf)
    inp = base[:, :K]
    W = (rn(R, K) / math.sqrt(K)).to(bf)
    Lb = rn(M, R + 8).to(bf)
    L = Lb[:, :R]                                  # row stride R + 8, like u inside [x | u]
    out = torch.full((M, R), 7.0, device=dev, dtype=bf)
    G, cs = z(R, K), z(K)
    ops.panel_fused_tc(inp, W, L, out, G, colsum=cs if colsum else None, scale=0.5)
    torch.cuda.synchronize()
    xb = inp.float()
    e = (rel(out, xb @ W.float().t()), rel(G, 0.5 * (L.float().t() @ xb)), rel(cs, 0.5 * xb.sum(0)) if colsum else 0.0)
    ok = e[0] < 6e-3 and e[1] < 1e-4 and e[2] < 1e-4
    print(f"check M={M:6d} K={K} pad={pad} colsum={int(colsum)}: out {e[0]:.2e}  G {e[1]:.2e}  colsum {e[2]:.2e}  {'OK' if ok else 'MISMATCH'}", flush=True)
    return ok


ok = True
for args in [(128, 2048), (1000, 2048, 8), (4097, 1024), (300, 2048, 0, False), (32768, 2048)]:
    ok = check(*args) and ok
print("ALL OK" if ok else "FAILED", flush=True)

hb = [(rn(B, H) / 8).to(bf) for _ in range(3)]
Wr = (rn(r, H) / math.sqrt(H)).to(bf)
Lp = rn(B, r).to(bf)
out = torch.empty(B, r, device=dev, dtype=bf)
G, cs = z(r, H), z(H)
for name, fn in [
    ("bf16 pass: skinny_rows + outer_reduce", lambda i: (ops.skinny_rows(hb[i % 3], Wr, out), ops.outer_reduce(Lp, hb[i % 3], G, colsum=cs))),
    ("bf16 pass: fused, mma.sync", lambda i: ops.panel_fused(hb[i % 3], Wr, Lp, out, G, colsum=cs)),
    ("bf16 pass: fused, tcgen05", lambda i: ops.panel_fused_tc(hb[i % 3], Wr, Lp, out, G, colsum=cs)),
]:
    ms = timeit(fn)
    print(f"{name:42s}: {ms*1e3:7.1f} us   {B*H*2/ms/1e9:5.2f} TB/s (one sweep of the bf16 matrix)", flush=True)

if ok:
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
    for name, opt in (("step, separate passes", 0), ("step, tcgen05 dpre pass", 2), ("step, separate passes (again)", 0), ("step, tcgen05 dpre pass (again)", 2)):
        ops.set_option("fused_panel", opt)
        for k in grads:
            grads[k].zero_()
        step(0)
        res[opt] = {k: v.clone() for k, v in grads.items()}
        ms = timeit(step, reps=60, warm=5)
        print(f"{name:34s}: {ms*1e3:8.1f} us/step  {B/ms/1e3:6.2f} M samples/s  {B*F/ms/1e9:6.0f} TFLOP/s", flush=True)
    for k in res[0]:
        print(f"grad {k:7s} tcgen05 vs separate: rel {rel(res[2][k], res[0][k]):.2e}")
    ops.set_option("fused_panel", -1)
