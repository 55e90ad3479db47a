"""One launch of each fused-panel variant (bf16 input, fp32 input) at the bench shape, for an ncu capture."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sample-efficient-multimodality_b200"))
import torch
from dmi_b200 import ops
dev, bf = "cuda", torch.bfloat16
B, H, r = 32768, 2048, 32
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)
hf = rn(B, H) / 8
hb = hf.to(bf)
Wr = (rn(r, H) / math.sqrt(H)).to(bf)
Lp = rn(B, r).to(bf)
out = torch.empty(B, r, device=dev, dtype=bf)
cp = torch.empty(B, H, device=dev, dtype=bf)
G, cs = torch.zeros(r, H, device=dev), torch.zeros(H, device=dev)
ops.panel_fused(hb, Wr, Lp, out, G, colsum=cs)
ops.panel_fused(hf, Wr, Lp, out, G, colsum=cs, copy=cp)
torch.cuda.synchronize()
print("done")
