"""Step time of the K-extension schedule vs the merged-weight schedule at the bench shape, and the
streaming side kernels against the shared-memory kernels they replace.  CUDA events, rotating inputs larger than L2."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sample-efficient-multimodality_b200"))
import torch
from dmi_b200 import ops
dev = "cuda"
B, D, H, r = int(os.environ.get("ROWS", 32768)), 768, 2048, 32
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)
w1, w2 = rn(H, D) / math.sqrt(D), rn(H, H) / math.sqrt(H)
b1, b2 = torch.zeros(H, device=dev), torch.zeros(H, device=dev)
A0, B0, A1, B1 = rn(D * r) / math.sqrt(D), rn(r * H) * 0.1, rn(H * r) / math.sqrt(H), rn(r * H) * 0.1
be0, be1 = torch.zeros(H, device=dev), torch.zeros(H, device=dev)
xs = [rn(B, D) for _ in range(3)]
dys = [rn(B, H) / math.sqrt(H) for _ in range(3)]
y = torch.empty(B, H, device=dev)
z = lambda *s: torch.zeros(*s, device=dev)
grads = dict(dA0=z(D, r), dB0=z(r, H), dbeta0=z(H), dA1=z(H, r), dB1=z(r, H), dbeta1=z(H))


def timeit(fn, reps=50, warm=5):
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


pk_old = ops.PackedProjector(D, H, r, dev)
pk_old.pack_base(w1, w2)
st_old = ops.MlpStash(B, D, H, r, dev, full=True)
pk_m = ops.PackedProjector(D, H, r, dev, merged=True)
st_m = ops.MlpStash(B, D, H, r, dev, full=True, merged=True)


def step_old(i):
    pk_old.pack_adapter(A0, B0, be0, A1, B1, be1, b1, b2)
    ops.adapted_mlp_fwd(pk_old, st_old, xs[i % 3], y)
    ops.adapted_mlp_bwd(pk_old, st_old, dys[i % 3], grads)


def step_merged(i):
    pk_m.pack_adapter_merged(w1, w2, A0, B0, be0, A1, B1, be1, b1, b2)
    ops.adapted_mlp_fwd(pk_m, st_m, xs[i % 3], y)
    ops.adapted_mlp_bwd(pk_m, st_m, dys[i % 3], grads)


F = 2 * D * H + 4 * H * H + 4 * r * D + 18 * r * H
for name, fn in (("K-extension schedule", step_old), ("merged schedule (side products in the GEMMs)", step_merged), ("K-extension schedule (again)", step_old)):
    ms = timeit(fn)
    print(f"{name:34s}: {ms*1e3:8.1f} us/step  {B/ms/1e3:6.2f} M samples/s  {B*F/ms/1e9:6.0f} TFLOP/s", flush=True)
print("pack merged:", f"{timeit(lambda i: pk_m.pack_adapter_merged(w1, w2, A0, B0, be0, A1, B1, be1, b1, b2))*1e3:.1f} us",
      " pack K-ext:", f"{timeit(lambda i: pk_old.pack_adapter(A0, B0, be0, A1, B1, be1, b1, b2))*1e3:.1f} us")
# ---- side kernels alone ----
hb = [rn(B, H).to(torch.bfloat16) for _ in range(3)]
xf = xs
Wr = (rn(r, H) / math.sqrt(H)).to(torch.bfloat16)
Wd = (rn(r, D) / math.sqrt(D)).to(torch.bfloat16)
out = torch.empty(B, r, device=dev, dtype=torch.bfloat16)
lq = torch.empty(ops.lq_words(B, r), device=dev, dtype=torch.int32)
cp = torch.empty(B, H, device=dev, dtype=torch.bfloat16)
Lp = rn(B, r).to(torch.bfloat16)
ops.stream_project(hb[0], Wr, out_lq=lq)
G = z(r, H); Gt = z(H, r); cs = z(H)
rows = [
    ("project bf16 [B,2048]x32  stream", lambda i: ops.stream_project(hb[i % 3], Wr, out_lq=lq), B * H * 2),
    ("project bf16 [B,2048]x32  stream, 148 CTAs", lambda i: ops.stream_project(hb[i % 3], Wr, out_lq=lq, max_ctas=148), B * H * 2),
    ("project bf16 [B,2048]x32  smem", lambda i: ops.skinny_rows(hb[i % 3], Wr, out), B * H * 2),
    ("project f32+copy [B,2048]x32 stream", lambda i: ops.stream_project(dys[i % 3], Wr, out_lq=lq, copy=cp), B * H * 6),
    ("project f32+copy [B,2048]x32 smem", lambda i: ops.skinny_rows(dys[i % 3], Wr, out, copy=cp), B * H * 6),
    ("reduce [B,2048] colsum    stream", lambda i: ops.stream_reduce(lq, r, hb[i % 3], G, colsum=cs), B * H * 2),
    ("reduce [B,2048] colsum    stream, 148 CTAs", lambda i: ops.stream_reduce(lq, r, hb[i % 3], G, colsum=cs, max_ctas=148), B * H * 2),
    ("reduce [B,2048] colsum    smem", lambda i: ops.outer_reduce(Lp, hb[i % 3], G, colsum=cs), B * H * 2),
    ("reduce [B,2048] transposed stream", lambda i: ops.stream_reduce(lq, r, hb[i % 3], Gt, transpose_out=True), B * H * 2),
    ("reduce [B,2048] transposed smem", lambda i: ops.outer_reduce(Lp, hb[i % 3], Gt, transpose_out=True), B * H * 2),
    ("cvt f32->bf16 [B,2048] (torch)", lambda i: cp.copy_(dys[i % 3]), B * H * 6),
]
for name, fn, nbytes in rows:
    ms = timeit(fn, reps=30)
    print(f"{name:46s}: {ms*1e3:7.1f} us  {nbytes/ms/1e9:5.2f} TB/s", flush=True)
