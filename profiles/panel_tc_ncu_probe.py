"""One launch of the tcgen05 fused dpre pass and of its merged-column-sum variant at the bench shape, for an ncu capture."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sample-efficient-multimodality_b200"))
import torch
from dmi_b200 import ops
dev, bf = "cuda", torch.bfloat16
B, H, r = 32768, 2048, 32
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)
hb = (rn(B, H) / 8).to(bf)
Wr = (rn(r, H) / math.sqrt(H)).to(bf)
Lp = rn(B, r).to(bf)
out = torch.empty(B, r, device=dev, dtype=bf)
G, cs = torch.zeros(r, H, device=dev), torch.zeros(H, device=dev)
ops.panel_fused_tc(hb, Wr, Lp, out, G, colsum=cs)
if os.environ.get("DMI_EXPERIMENTAL") == "1":
    ops.panel_fused_tc(hb, Wr, Lp, out, G, colsum=cs, merged_colsum=True)
torch.cuda.synchronize()
print("done")
