"""Bisect of the CTA-pair kernel's MMA-side overhead (measurement modes of dmi_set_option("gemm_debug"))."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sample-efficient-multimodality_b200"))
import torch
from dmi_b200 import ops
M, N, K = 32768, 2048, 2080
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
Bm = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(torch.bfloat16)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
fl = 2.0 * M * N * K


def t(reps=30):
    for _ in range(3):
        ops.gemm_tn(A, Bm, out0=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        ops.gemm_tn(A, Bm, out0=out)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for rnd in range(2):
    for pair in (0, 1):
        for dbg in (0, 3, 3 + 32, 3 + 32 + 64, 3 + 32 + 128, 3 + 16 + 32 + 64):
            if pair == 0 and dbg > 3:
                continue
            ops.set_option("gemm_pair", pair)
            ops.set_option("gemm_debug", 0)
            ops.gemm_tn(A, Bm, out0=out)          # refresh the operand ring with real data before a no-TMA mode
            ops.set_option("gemm_debug", dbg)
            ms = t()
            print(f"round {rnd} pair={pair} debug={dbg:3d}: {ms*1e3:7.1f} us {fl/ms/1e9:6.0f} TFLOP/s", flush=True)
ops.set_option("gemm_debug", 0); ops.set_option("gemm_pair", -1)
