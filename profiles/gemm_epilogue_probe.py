"""Where does the pair GEMM's time go beyond its main loop?  Times the three GEMMs of the step with the epilogue switched off piece by
piece (dmi_set_option("gemm_debug", bits)): 0 = full kernel, 16 = everything but the TMA stores, 64 = stores with an L2 evict-first
hint, 1 = no epilogue at all, 3 = no epilogue and no TMA loads.   python profiles/gemm_epilogue_probe.py"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sample-efficient-multimodality_b200"))
import torch  # noqa: E402

from dmi_b200 import ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
M = 32768


def timeit(fn, reps=20):
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


bf = torch.bfloat16
for (N, K, mode, name) in [(2048, 800, 1, "layer0_gelu"), (2048, 2080, 0, "layer1_store_f32"), (2048, 2080, 2, "dpre_gelugrad")]:
    A = torch.randn(M, K, device="cuda", generator=g).to(bf)
    Bm = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(bf)
    bias = torch.randn(N, device="cuda", generator=g)
    fl = 2.0 * M * N * K
    if mode == 0:
        out = torch.empty(M, N, device="cuda")
        fn = lambda: ops.gemm_tn(A, Bm, bias=bias, out0=out)
    elif mode == 1:
        h, pre = torch.empty(M, N, device="cuda", dtype=bf), torch.empty(M, N, device="cuda", dtype=bf)
        fn = lambda: ops.gemm_tn(A, Bm, mode=ops.EPI_GELU, bias=bias, out0=h, out1=pre)
    else:
        pre = torch.randn(M, N, device="cuda", generator=g).to(bf)
        out = torch.empty(M, N, device="cuda", dtype=bf)
        fn = lambda: ops.gemm_tn(A, Bm, mode=ops.EPI_GELU_BWD, out0=out, aux=pre)
    for rnd in range(2):
        for dbg in (0, 16, 64, 1, 3):
            ops.set_option("gemm_debug", dbg)
            ms = timeit(fn)
            print(f"round {rnd} {name:18s} debug={dbg:3d}: {ms*1e3:8.1f} us {fl/ms/1e9:7.0f} TFLOP/s", flush=True)
    ops.set_option("gemm_debug", 0)
    del A, Bm
