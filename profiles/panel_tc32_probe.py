"""The fp32-input panel pass (dY sweep) alone at the bench shape: time (CUDA events) and, under ncu, one launch to capture."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sample-efficient-multimodality_b200"))
import torch
from dmi_b200 import ops
dev = "cuda"
M, K, R = 32768, 2048, 32
g = torch.Generator(device=dev).manual_seed(0)
inp32 = [torch.randn(M, K, device=dev, generator=g) for _ in range(2)]          # two inputs: 537 MB > L2
W = torch.randn(R, K, device=dev, generator=g).to(torch.bfloat16)
L = torch.randn(M, R, device=dev, generator=g).to(torch.bfloat16)
out = torch.empty(M, R, dtype=torch.bfloat16, device=dev)
G = torch.zeros(R, K, device=dev)
cs = torch.zeros(K, device=dev)
copy = torch.empty(M, K, dtype=torch.bfloat16, device=dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for i in range(3):
    ops.panel_fused_tc32(inp32[i & 1], W, L, out, G, colsum=cs, copy=copy)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(reps):
    ops.panel_fused_tc32(inp32[i & 1], W, L, out, G, colsum=cs, copy=copy)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / reps * 1e3
print(f"panel_fused_tc32 M={M} K={K}: {us:.1f} us  {M * K * 6 / us / 1e6:.2f} TB/s algorithmic", flush=True)
