"""One launch each of: cuBLAS, 1-CTA kernel, CTA-pair kernel (full and pure-MMA debug mode) inside a profiler window, for
`ncu --profile-from-start off --set full`: compares tensor-pipe duty and SM clock between the variants."""
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


def all_variants():
    torch.matmul(A, Bm.t())
    for pair in (0, 1):
        for dbg in (0, 3, 1, 2):
            ops.set_option("gemm_pair", pair)
            ops.set_option("gemm_debug", dbg)
            ops.gemm_tn(A, Bm, out0=out)
    ops.set_option("gemm_debug", 0)
    ops.set_option("gemm_pair", -1)


for _ in range(2):
    all_variants()
torch.cuda.synchronize()
torch.cuda.profiler.start()
all_variants()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
