"""Sustained (power-capped) throughput of each GEMM variant on the layer-1 shape: every variant runs alone for ~100 ms after a
0.7 s idle pause, with the SM clock sampled through NVML during the run."""
import math, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sample-efficient-multimodality_b200"))
import torch
import pynvml
from dmi_b200 import ops
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
M, N, K = 32768, 2048, 2080
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
Bm = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(torch.bfloat16)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
fl = 2.0 * M * N * K


def run(name, fn, reps=400):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    time.sleep(0.7)
    clocks, power, stop = [], [], threading.Event()

    def poll():
        while not stop.is_set():
            clocks.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
            time.sleep(0.005)
    th = threading.Thread(target=poll)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    th.start()
    evs[0].record()
    for q in range(4):
        for _ in range(reps // 4):
            fn()
        evs[q + 1].record()
    torch.cuda.synchronize()
    stop.set()
    th.join()
    qs = [evs[i].elapsed_time(evs[i + 1]) / (reps // 4) for i in range(4)]
    clocks.sort()
    print(f"{name:28s} quarters us: " + " ".join(f"{q*1e3:7.1f}" for q in qs) + f"  last-quarter {fl/qs[3]/1e9:6.0f} TFLOP/s  clk med {clocks[len(clocks)//2]} min {clocks[0]} max {clocks[-1]} MHz  P max {max(power):.0f} W", flush=True)


import sys as _s
DBGS = [int(a) for a in _s.argv[1:]] or [0, 3]
run("cuBLAS", lambda: torch.matmul(A, Bm.t()))
for pair, cl in ((0, 0),) if len(DBGS) > 2 else ((0, 0), (1, 0), (0, 1)):
    for dbg in DBGS:
        ops.set_option("gemm_pair", pair)
        ops.set_option("gemm_cluster", cl)
        ops.set_option("gemm_debug", dbg)
        run(f"ours pair={pair} cluster={cl} debug={dbg}", lambda: ops.gemm_tn(A, Bm, out0=out))
ops.set_option("gemm_debug", 0); ops.set_option("gemm_pair", -1); ops.set_option("gemm_cluster", -1)
run("cuBLAS again", lambda: torch.matmul(A, Bm.t()))
