"""GEMM headroom probe: times the three big contractions of the adapted-MLP step with (a) cuBLAS (torch.matmul, bf16 out),
(b) this library's tcgen05 kernel in its default / CTA-pair / no-epilogue / no-TMA configurations, all in one process with
CUDA events, inputs larger than L2.   python profiles/gemm_headroom_probe.py [--M 32768]"""
import argparse
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sample-efficient-multimodality_b200"))
import torch  # noqa: E402

from dmi_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, default=32768)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
g = torch.Generator(device="cuda").manual_seed(0)


def timeit(fn, reps=a.reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for (N, K, mode, name) in [(2048, 800, 1, "layer0_gelu"), (2048, 2080, 0, "layer1_store"), (2048, 2080, 2, "dpre_gelugrad"),
                           (2048, 768, 1, "layer0_gelu_K768"), (2048, 2048, 0, "layer1_store_K2048")]:
    M = a.M
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    Bm = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    fl = 2.0 * M * N * K
    ms = timeit(lambda: torch.matmul(A, Bm.t()))
    print(f"{name:22s} M={M} N={N} K={K}  cuBLAS bf16-out        : {ms*1e3:8.1f} us {fl/ms/1e9:7.0f} TFLOP/s", flush=True)
    if mode == 0:
        out = torch.empty(M, N, device="cuda")
        fn = lambda: ops.gemm_tn(A, Bm, bias=bias, out0=out)
        outb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        fnb = lambda: ops.gemm_tn(A, Bm, bias=bias, out0=outb)
    elif mode == 1:
        h = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        pre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        fn = lambda: ops.gemm_tn(A, Bm, mode=ops.EPI_GELU, bias=bias, out0=h, out1=pre)
        fnb = None
    else:
        pre = torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        fn = lambda: ops.gemm_tn(A, Bm, mode=ops.EPI_GELU_BWD, out0=out, aux=pre)
        fnb = None
    for pair in (0, 1):
        for dbg in (0, 1, 2, 3):
            ops.set_option("gemm_pair", pair)
            ops.set_option("gemm_debug", dbg)
            ms = timeit(fn)
            print(f"{name:22s} M={M} N={N} K={K}  ours pair={pair} debug={dbg}        : {ms*1e3:8.1f} us {fl/ms/1e9:7.0f} TFLOP/s", flush=True)
    ops.set_option("gemm_debug", 0)
    if fnb is not None:
        for pair in (0, 1):
            ops.set_option("gemm_pair", pair)
            ms = timeit(fnb)
            print(f"{name:22s} M={M} N={N} K={K}  ours pair={pair} bf16-out        : {ms*1e3:8.1f} us {fl/ms/1e9:7.0f} TFLOP/s", flush=True)
    ops.set_option("gemm_pair", -1)
    del A, Bm
# a big square for calibration against MEASURED_PEAKS (8192^3)
A = torch.randn(8192, 8192, device="cuda", generator=g).to(torch.bfloat16)
Bm = torch.randn(8192, 8192, device="cuda", generator=g).to(torch.bfloat16)
ms = timeit(lambda: torch.matmul(A, Bm.t()))
print(f"cuBLAS 8192^3: {ms*1e3:.1f} us {2.0*8192**3/ms/1e9:.0f} TFLOP/s")
out = torch.empty(8192, 8192, device="cuda", dtype=torch.bfloat16)
for pair in (0, 1):
    ops.set_option("gemm_pair", pair)
    ms = timeit(lambda: ops.gemm_tn(A, Bm, out0=out))
    print(f"ours pair={pair} 8192^3 bf16-out: {ms*1e3:.1f} us {2.0*8192**3/ms/1e9:.0f} TFLOP/s")
