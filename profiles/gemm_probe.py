"""Small driver used under ncu: runs one big-GEMM shape of the adapted MLP a few times.
   python profiles/gemm_probe.py --mode 0|1|2 --cluster -1|0|1 [--M 16384 --N 2048 --K 2080]"""
import argparse
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sample-efficient-multimodality_b200"))
import torch  # noqa: E402

from dmi_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--cluster", type=int, default=-1)
ap.add_argument("--pair", type=int, default=-1)
ap.add_argument("--debug", type=int, default=0)
ap.add_argument("--M", type=int, default=16384)
ap.add_argument("--N", type=int, default=2048)
ap.add_argument("--K", type=int, default=2080)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
ops.set_option("gemm_cluster", a.cluster)
ops.set_option("gemm_pair", a.pair)
ops.set_option("gemm_debug", a.debug)
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(a.M, a.K, device="cuda", generator=g).to(torch.bfloat16)
B = (torch.randn(a.N, a.K, device="cuda", generator=g) / math.sqrt(a.K)).to(torch.bfloat16)
bias = torch.randn(a.N, device="cuda", generator=g)
if a.mode == 0:
    out = torch.empty(a.M, a.N, device="cuda")
    fn = lambda: ops.gemm_tn(A, B, bias=bias, out0=out)
elif a.mode == 1:
    h = torch.empty(a.M, a.N, device="cuda", dtype=torch.bfloat16)
    pre = torch.empty(a.M, a.N, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.gemm_tn(A, B, mode=ops.EPI_GELU, bias=bias, out0=h, out1=pre)
else:
    pre = torch.randn(a.M, a.N, device="cuda", generator=g).to(torch.bfloat16)
    out = torch.empty(a.M, a.N, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.gemm_tn(A, B, mode=ops.EPI_GELU_BWD, out0=out, aux=pre)
for _ in range(3):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    fn()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
print(f"mode={a.mode} debug={a.debug} pair={a.pair} cluster={a.cluster} M={a.M} N={a.N} K={a.K}: {ms*1e3:.1f} us  {2.0*a.M*a.N*a.K/ms/1e9:.0f} TFLOP/s")
