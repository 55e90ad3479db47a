"""Tile-shape sweep of the bf16 row-panel projection (dmi_set_option("skinny_variant")): rows per CTA x cp.async stages."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sample-efficient-multimodality_b200"))
import torch
from dmi_b200 import ops, augment as A
B, H, r = 32768, 2048, 32
g = torch.Generator(device="cuda").manual_seed(0)
hb = [torch.randn(B, H, device="cuda", generator=g).to(torch.bfloat16) for _ in range(3)]
W = (torch.randn(r, H, device="cuda", generator=g) / math.sqrt(H)).to(torch.bfloat16)
out = torch.empty(B, r, device="cuda", dtype=torch.bfloat16)
names = {0: "64 rows x 4 stages (default)", 1: "64 x 2", 2: "32 x 4", 3: "128 x 2", 4: "64 x 3", 5: "32 x 2"}
ref = None
for v in range(6):
    ops.set_option("skinny_variant", v)
    ops.skinny_rows(hb[0], W, out)
    torch.cuda.synchronize()
    if ref is None:
        ref = out.clone()
    ok = torch.equal(out, ref)
    for _ in range(3):
        ops.skinny_rows(hb[1], W, out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(30):
        ops.skinny_rows(hb[i % 3], W, out)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 30
    print(f"variant {v} {names[v]:28s}: {ms*1e3:6.1f} us  {B*H*2/ms/1e9:5.2f} TB/s  same_result={ok}", flush=True)
ops.set_option("skinny_variant", 0)
for n in (768, 2048):
    A.get_rotation_matrix_device(n, "cuda"); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): A.get_rotation_matrix_device(n, "cuda")
    b.record(); torch.cuda.synchronize()
    print(f"device Haar n={n}: {a.elapsed_time(b)/5:.2f} ms")
