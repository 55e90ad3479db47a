#!/usr/bin/env python
"""Benchmark of the adapted-projector hot path (BASELINE.json metric: adapted-projector samples/s, fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" = one pass of the hot path over one batch: adapter -> operand layout, fused adapted MLP2 forward, backward to
the adapter factors (dA, dB, dbeta of both layers), and -- for N > 1 -- the bucketed NCCL all-reduce of those gradients,
overlapped with the layer-0 backward.  Workload (config.workload): the centre point of BASELINE.json configs[4]
(dim sweep) D=768, H=2048, r=32 at a per-GPU batch large enough to fill the chip; weak scaling (fixed rows per GPU).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "sample-efficient-multimodality_b200")
for _p in (ROOT, PKG_DIR):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "adapted_projector_fwd_bwd_samples_per_sec"
UNIT = "samples/s"


def flops_per_sample(D, H, r):
    """algorithmic FLOPs of the full adapted projector fwd+bwd with the base frozen"""
    return 2 * D * H + 4 * H * H + 4 * r * D + 18 * r * H


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(bf16_burst=p["bf16_tflops"], bf16_sustained=p["bf16_tflops_sustained"], hbm=p["hbm_gbs"], source="measured")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback")


def bind_to_gpu_numa_node(index: int):
    """Best effort: run this process on the CPUs of the NUMA node the GPU hangs off, so that pinned host buffers (first touch inside
    cudaHostAlloc) and the copy-issuing thread are local to the GPU's PCIe root.  At N = 8 round 1 measured 23 GB/s of H2D per GPU
    against 42 GB/s at N = 1 with all ranks on the default node.  Returns a description for the JSON line."""
    try:
        pr = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return {"pci": bdf, "numa_node": node, "bound": False, "why": "no NUMA information for the device"}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if not use:
            return {"pci": bdf, "numa_node": node, "bound": False, "why": "none of the node's CPUs is in this process's affinity mask"}
        os.sched_setaffinity(0, use)
        return {"pci": bdf, "numa_node": node, "bound": True, "cpus": len(use)}
    except Exception as e:
        return {"bound": False, "why": repr(e)[:120]}


class ClockSampler(threading.Thread):
    """polls NVML for SM clock / throttle reasons while the timed region runs"""

    def __init__(self, index: int, period_s: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
                 "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80)}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# -------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port (torch CPU fp32, autograd) of the same computation on the host cores
# -------------------------------------------------------------------------------------------------------------------
def cpu_problem(B, D, H, r, seed=0):
    g = torch.Generator().manual_seed(seed)
    w1 = torch.randn(H, D, generator=g) / math.sqrt(D)
    b1 = torch.zeros(H)
    w2 = torch.randn(H, H, generator=g) / math.sqrt(H)
    b2 = torch.zeros(H)
    x = torch.randn(B, D, generator=g)
    x = x / x.norm(dim=1, keepdim=True)
    a = [torch.randn(D * r, generator=g) / math.sqrt(D), torch.randn(H * r, generator=g) / math.sqrt(H)]
    b = [torch.randn(r * H, generator=g) * 0.1, torch.randn(r * H, generator=g) * 0.1]
    beta = [torch.zeros(H), torch.zeros(H)]
    dy = torch.randn(B, H, generator=g) / math.sqrt(H)
    return w1, b1, w2, b2, x, a, b, beta, dy


def time_cpu_port(D, H, r, sample_rows, steps, warmup, best_of=False, budget_s=None):
    """samples/s of oracle.adapted_mlp_full_grads (the reference's op sequence: F.linear + skinny matmuls + autograd) on the CPU.
    best_of: report the fastest step (3 warm-ups, best of 10) instead of the mean.  budget_s bounds the
    timed steps (the first warm-up step is the estimate); returns (samples/s, seconds per step, steps actually timed)."""
    from oracle import oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    prob = cpu_problem(sample_rows, D, H, r)
    t_est = None
    for _ in range(warmup):
        t0 = time.perf_counter()
        O.adapted_mlp_full_grads(*prob)
        t_est = time.perf_counter() - t0
    if budget_s is not None and t_est is not None:
        steps = max(3, min(steps, int(budget_s / max(t_est, 1e-6))))
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        O.adapted_mlp_full_grads(*prob)
        ts.append(time.perf_counter() - t0)
    dt = min(ts) if best_of else sum(ts) / len(ts)
    return sample_rows / dt, dt, steps


def run_reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path on the box's host cores (the oracle port: the reference is pure Python /
    torch and /root/reference does not exist on the GPU box), all host threads, on the SAME config as the GPU arm: the same rows
    per step (args.batch) and the same warm-up count; the step count is the driver's unless the run would exceed ~150 s."""
    if rank != 0:
        return
    rows = args.batch
    warm = max(3, args.warmup)
    sps, dt, steps = time_cpu_port(args.D, args.H, args.r, rows, args.steps, warm, budget_s=150.0)
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": sps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "rows_per_gpu": rows, "D": args.D, "H": args.H, "r": args.r, "global_batch": rows,
                       "parallelism": "cpu", "flops_per_sample": flops_per_sample(args.D, args.H, args.r),
                       "note": "CPU arm: one process on the host cores whatever --gpus says; steps capped so that the run ends within ~150 s"},
            "cpu_baseline": {"value": sps, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{rows} rows per step (the GPU arm's per-GPU batch), oracle port (torch CPU fp32 autograd), mean of {steps} steps after {warm} warm-ups"},
            "e2e": {"value": sps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_name(args):
    return (f"BASELINE configs[4] dim-sweep centre point: full adapted MLP2 projector fwd+bwd, D={args.D} H={args.H} r={args.r}, "
            f"{args.batch} rows per GPU, frozen base, grads to A/B/beta of both layers")


# -------------------------------------------------------------------------------------------------------------------
# our arm
# -------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32768, help="rows per GPU (weak scaling: fixed per GPU)")
    ap.add_argument("--D", type=int, default=768)
    ap.add_argument("--H", type=int, default=2048)
    ap.add_argument("--r", type=int, default=32)
    ap.add_argument("--cpu-rows", type=int, default=2048, help="rows per step of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-eager", action="store_true", help="run the end-to-end arm eagerly instead of replaying a captured CUDA graph")
    ap.add_argument("--no-kernel-breakdown", action="store_true")
    ap.add_argument("--dp-buckets", type=int, default=2, choices=[1, 2], help="NCCL reducer: gradient all-reduce buckets per step (2: layer-1 bucket overlaps the layer-0 backward)")
    ap.add_argument("--dp-reduce", default="symm", choices=["symm", "nccl"],
                    help="N > 1 gradient all-reduce: 'symm' = this library's one-shot kernel over NVSwitch peer memory, in the step's stream "
                         "(falls back to 'nccl' if symmetric memory cannot be set up); 'nccl' = torch.distributed all-reduce on a side stream")
    ap.add_argument("--dp-stream", default="side", choices=["side", "inline"],
                    help="'symm' reducer: run the one-shot all-reduce kernel on a side stream underneath the next step (default) or in the step's stream")
    ap.add_argument("--dp-no-multicast", action="store_true", help="'symm' reducer: peer loads over NVLink instead of the NVLS multicast address (the fallback path)")
    ap.add_argument("--no-llm", action="store_true", help="skip the configs[1] micro-step with a random-init Llama-3.2-1B")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary BASELINE configs (hypernet micro-step, few-shot, plain projector)")
    ap.add_argument("--sweep", action="store_true", help="full BASELINE configs[4] sweep (D 512..4096 x r 8..64) instead of the default reduced one")
    ap.add_argument("--no-sweep", action="store_true", help="skip the dimension sweep")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the PyTorch-eager-on-the-same-GPU baseline of the reference's op sequence")
    ap.add_argument("--fused-panel", type=int, default=-1, help="dmi_set_option('fused_panel', v): 0 = mma.sync side passes only, 1 = tcgen05 panel kernels at any size")
    ap.add_argument("--no-pdl", action="store_true", help="launch without programmatic dependent launch (A/B of the launch-gap overlap)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"            # keep stdout to the one JSON line
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))

    from dmi_b200 import ops
    from dmi_b200._lib import MlpArgs  # noqa: F401
    from dmi_b200.parallel import BucketAllReducer, FlatGrads
    if args.no_pdl:
        ops.set_option("pdl", 0)
    if args.fused_panel >= 0:
        ops.set_option("fused_panel", args.fused_panel)

    B, D, H, r = args.batch, args.D, args.H, args.r
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    gw = torch.Generator(device=dev).manual_seed(42)          # replicated parameters: same seed on every rank
    randn = lambda *s, gen=g: torch.randn(*s, device=dev, generator=gen)
    w1 = randn(H, D, gen=gw) / math.sqrt(D)
    b1 = torch.zeros(H, device=dev)
    w2 = randn(H, H, gen=gw) / math.sqrt(H)
    b2 = torch.zeros(H, device=dev)
    A0 = randn(D * r, gen=gw) / math.sqrt(D)
    B0 = randn(r * H, gen=gw) * 0.1
    A1 = randn(H * r, gen=gw) / math.sqrt(H)
    B1 = randn(r * H, gen=gw) * 0.1
    beta0 = torch.zeros(H, device=dev)
    beta1 = torch.zeros(H, device=dev)
    NBUF = 3                                                   # rotating inputs (each step reads x + dy = 369 MB at the default batch, far beyond the 126 MB L2)
    xs, dys = [], []
    for _ in range(NBUF):
        x = randn(B, D)
        xs.append(x / x.norm(dim=1, keepdim=True))
        dys.append(randn(B, H) / math.sqrt(H))

    pk = ops.PackedProjector(D, H, r, dev)
    pk.pack_base(w1, w2)
    st = ops.MlpStash(B, D, H, r, dev, full=True)
    y = torch.empty(B, H, device=dev)
    shapes = dict(dA1=(H, r), dB1=(r, H), dbeta1=(H,), dA0=(D, r), dB0=(r, H), dbeta0=(H,))
    buckets = [["dA1", "dB1", "dbeta1"], ["dA0", "dB0", "dbeta0"]]
    symm, dp_info = None, None
    if world > 1 and args.dp_reduce == "symm":
        try:
            from dmi_b200.parallel import SymmAllReducer
            n_flat = FlatGrads(shapes, buckets, "cpu").flat.numel()
            symm = SymmAllReducer(n_flat, dev, n_slots=2, use_multicast=not args.dp_no_multicast)
        except Exception as e:          # symmetric memory unavailable on this box / torch build: say so and use NCCL
            symm = None
            dp_info = {"kind": "nccl", "symm_error": repr(e)[:300]}
    if symm is not None:
        grads = [FlatGrads(shapes, buckets, dev, storage=symm.inputs[k]) for k in range(2)]   # gradients accumulate straight into peer-mapped memory
        reducer = None
        # one-off check of the kernel against NCCL on random data (both slots)
        chk = torch.randn(symm.numel, device=dev, generator=g)
        ref = chk.clone()
        dist.all_reduce(ref)
        errs = []
        for k in range(2):
            symm.inputs[k].copy_(chk)
            errs.append(float((symm.reduce(k) - ref).abs().max().item()))
            symm.inputs[k].zero_()
        dp_info = {"kind": "symm_oneshot_multimem" if symm.multicast else "symm_oneshot_peer_loads", "max_abs_diff_vs_nccl": max(errs), "stream": args.dp_stream,
                   "how": "dmi_allreduce_oneshot: barrier + multimem.ld_reduce (NVLS) / peer loads + barrier; 'side': on a side stream after the last gradient "
                          "kernel, its 128-thread CTAs co-resident with the next step's persistent GEMM CTAs, waited for at the gradient buffer's next use; "
                          "'inline': in the step's own stream"}
    else:
        grads = [FlatGrads(shapes, buckets, dev) for _ in range(2)]           # double-buffered so the all-reduce of step i overlaps step i+1
        reducer = BucketAllReducer(average=False) if world > 1 else None     # the 1/world factor is folded into grad_scale
        if world > 1 and dp_info is None:
            dp_info = {"kind": "nccl"}
    ev_l1 = [torch.cuda.Event() for _ in range(2)]

    done_ev = [None, None]          # all-reduce of the step that last used gradient buffer k has finished

    def step(i, comm=True):
        k = i & 1
        gbuf = grads[k]
        if done_ev[k] is not None:
            # point of USE of the double-buffered gradient buffer: the all-reduce issued two steps ago must be done before the buffer
            # is zeroed again.  (An optimizer consuming step i's reduced gradients waits on the same event, one step later.)
            torch.cuda.current_stream().wait_event(done_ev[k])
            done_ev[k] = None
        gbuf.zero_()
        pk.pack_adapter(A0, B0, beta0, A1, B1, beta1, b1, b2)
        ops.adapted_mlp_fwd(pk, st, xs[i % NBUF], y)
        ops.adapted_mlp_bwd(pk, st, dys[i % NBUF], gbuf.views, grad_scale=1.0 / world, layer1_event=ev_l1[k] if reducer is not None else None)
        if reducer is not None and comm:
            if args.dp_buckets == 2:
                reducer.reduce_bucket(gbuf.buckets[0], ev_l1[k])           # layer-1 grads: overlaps the layer-0 backward
                reducer.reduce_bucket(gbuf.buckets[1], None)               # layer-0 grads: overlaps the next step's forward
            else:
                reducer.reduce_bucket(gbuf.flat, None)                     # one all-reduce of the whole flat gradient buffer
            done_ev[k] = reducer.done_event()
        if symm is not None and comm:
            if args.dp_stream == "inline":
                symm.reduce(k)                                             # in this stream: reduced gradients land in symm.outputs[k]
            else:
                done_ev[k] = symm.reduce_async(k)                          # side stream, small co-resident CTAs: the next step runs underneath

    def drain():
        """every all-reduce issued so far completes on the compute stream (inside the timed region)"""
        if reducer is not None:
            reducer.wait()
        for k in range(2):
            if done_ev[k] is not None:
                torch.cuda.current_stream().wait_event(done_ev[k])
                done_ev[k] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)      # NVML init BEFORE the barrier: it takes milliseconds and differs per rank, and a rank that enters the
    for i in range(args.warmup):             # timed loop late makes every other rank wait at the first all-reduce (measured: +0.23 ms/step over 20 steps)
        step(i)
    drain()
    if rank == 0:
        sampler.start()
    barrier()
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i)
    drain()
    e1.record()
    barrier()
    launches = ops.launch_count() - l0
    ms = e0.elapsed_time(e1) / args.steps
    # the same loop without the gradient all-reduce: the difference is the communication time the overlap did NOT hide
    ms_nocomm = None
    if world > 1:
        n2 = max(5, min(args.steps, 50))
        for i in range(2):
            step(i, comm=False)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for i in range(n2):
            step(i, comm=False)
        c1.record()
        barrier()
        ms_nocomm = c0.elapsed_time(c1) / n2
    # keep the GPU under the same load a little longer if the timed region was too short for the clock sampler
    if rank == 0 and len(sampler.samples) < 5:
        t_end = time.time() + 0.4
        i = 0
        while time.time() < t_end:
            step(i, comm=False)          # rank-0-only load for the clock sampler: must not issue collectives
            i += 1
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms, ms_nocomm], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_nocomm = float(t[0].item()), float(t[1].item())
    value = world * B / (ms * 1e-3)

    peaks = load_peaks()
    F = flops_per_sample(D, H, r)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(args), "rows_per_gpu": B, "D": D, "H": H, "r": r, "global_batch": world * B,
                       "parallelism": f"dp{world}", "l2_policy": "inputs larger than L2 (x+dy = %.0f MB per step, 3 rotating buffers)" % ((B * D + B * H) * 4 / 1e6),
                       "flops_per_sample": F},
            "gpu_launches": int(launches),
            "step_tensor_tflops": value / world * F / 1e12,
            "step_tensor_frac_of_sustained": value / world * F / 1e12 / peaks["bf16_sustained"],
            "peaks": peaks}
    if clocks is not None:
        line["clocks"] = clocks
    if world > 1 and symm is not None:
        # diagnostics of the one-shot kernel: its latency with every rank in lockstep and nothing else running, and the time the step's
        # stream spends in it inside the loop (waiting for the slowest rank at the barrier + the kernel), max over ranks
        dist.barrier()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for i in range(50):
            symm.reduce(i & 1)
        a1.record()
        torch.cuda.synchronize()
        ms_alone = a0.elapsed_time(a1) / 50
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
        for i in range(3):
            step(i)
        dist.barrier()
        for i in range(20):
            k = i & 1
            grads[k].zero_()
            pk.pack_adapter(A0, B0, beta0, A1, B1, beta1, b1, b2)
            ops.adapted_mlp_fwd(pk, st, xs[i % NBUF], y)
            ops.adapted_mlp_bwd(pk, st, dys[i % NBUF], grads[k].views, grad_scale=1.0 / world)
            evs[i][0].record()
            symm.reduce(k)
            evs[i][1].record()
        torch.cuda.synchronize()
        ms_in = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
        t2 = torch.tensor([ms_alone, ms_in], device=dev)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        dp_info["ms_kernel_alone_lockstep"] = float(t2[0].item())
        dp_info["ms_in_step_wait_plus_kernel"] = float(t2[1].item())
    if world > 1:
        line["host_numa_binding_rank0"] = numa
        nbytes = grads[0].flat.numel() * 4
        line["comm"] = {"allreduce_bytes_per_step": nbytes, "reduce": dp_info, "ms_per_step_without_allreduce": ms_nocomm,
                        "ms_exposed_per_step": ms - ms_nocomm,
                        "how": "exposed = timed loop with minus without the gradient all-reduce (max over ranks; everything drained inside the timed region). "
                               "nccl reducer: side stream, layer-1 bucket overlaps the layer-0 backward, layer-0 bucket overlaps the next step, wait at the "
                               "double-buffered gradient buffer's next use"}

    # ---------------- sustained value: the same step for >= 200 back-to-back iterations (power-capped regime) ----------------
    if world == 1:
        n_s = max(200, args.steps)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(n_s):
            step(i)
        s1.record()
        torch.cuda.synchronize()
        ms_s = s0.elapsed_time(s1) / n_s
        line["sustained"] = {"steps": n_s, "ms_per_step": ms_s, "value": B / (ms_s * 1e-3), "unit": UNIT,
                             "step_tensor_tflops": B / (ms_s * 1e-3) * F / 1e12,
                             "frac_of_bf16_sustained_peak": B / (ms_s * 1e-3) * F / 1e12 / peaks["bf16_sustained"],
                             "frac_of_bf16_burst_peak": B / (ms_s * 1e-3) * F / 1e12 / peaks["bf16_burst"]}
    line["step_tensor_frac_of_burst"] = value / world * F / 1e12 / peaks["bf16_burst"]
    # ---------------- per-kernel breakdown + roofline of the dominant kernel (rank 0, N = 1 only) ----------------
    if rank == 0 and world == 1 and not args.no_kernel_breakdown:
        line.update(kernel_breakdown(ops, pk, st, y, xs, dys, grads[0], B, D, H, r, peaks, ms))
    # ---------------- end-to-end through the public module API with host inputs ----------------
    _trace("headline done")
    if not args.no_e2e:
        # with >= 4 ranks the graph-captured end-to-end arms failed intermittently with a launch failure while the
        # fp32 panel kernel had two converter groups skipping mbarrier phases.  That was removed, but the fix could not be re-verified
        # at >= 4 ranks within the round's GPU budget, so the multi-rank end-to-end arms -- bound by the host link (2 GPUs per PCIe
        # uplink), not by the side passes -- keep running the mma.sync side passes (--fused-panel 0 | 1 given explicitly overrides).
        e2e_guard = world > 1 and args.fused_panel < 0
        if e2e_guard:
            ops.set_option("fused_panel", 0)
        e2e = run_e2e(args, dev, rank, world, w1, b1, w2, b2, (A0, B0, beta0, A1, B1, beta1))
        _trace("e2e (bf16 host) done")
        if e2e is not None:
            if e2e_guard:
                e2e["side_passes"] = "mma.sync kernels (fused_panel = 0) in the multi-rank end-to-end arms, see DESIGN section 8"
            line["e2e"] = e2e
        try:
            e2f = run_e2e(args, dev, rank, world, w1, b1, w2, b2, (A0, B0, beta0, A1, B1, beta1), host_dtype=torch.float32)
            if e2f is not None:
                line["e2e_f32_host"] = e2f
            _trace("e2e (f32 host) done")
        except Exception as e:
            if rank == 0:
                line["e2e_f32_host"] = {"error": repr(e)[:300]}
        try:
            e2s = run_e2e_store(args, dev, rank, world, w1, b1, w2, b2, (A0, B0, beta0, A1, B1, beta1))
            if e2s is not None:
                line["e2e_store"] = e2s
            _trace("e2e (store) done")
        except Exception as e:
            if rank == 0:
                line["e2e_store"] = {"error": repr(e)[:300]}
        if e2e_guard:
            torch.cuda.synchronize()
            ops.set_option("fused_panel", -1)
    if rank == 0 and world == 1 and not args.no_extras:
        try:
            ms_h1 = _time_adapted_step(dev, B, D, H, r, steps=20, flags=1)
            F1 = 2 * D * H + 4 * r * D + 6 * r * H
            line["as_written_h1"] = {"ms_per_step": ms_h1, "value": B / ms_h1 * 1e3, "unit": UNIT, "flops_per_sample": F1, "tflops": B / ms_h1 * 1e3 * F1 / 1e12,
                                     "what": "the shipped default lora_forward_mode='as_written' (reference Projector.lora_forward stops after the first GELU, "
                                             "projector.py:124 / SURVEY H1): layer-0 adapted GEMM + GELU forward, gradients to A0/B0/beta0; same rows, D, H, r"}
        except Exception as e:
            line["as_written_h1"] = {"error": repr(e)[:300]}
    if rank == 0 and world == 1 and not args.no_gpu_eager:
        try:
            line["gpu_eager_baseline"] = gpu_eager_baseline(dev, B, D, H, r, w1, b1, w2, b2, (A0, B0, beta0, A1, B1, beta1), xs[0], dys[0])
        except Exception as e:
            line["gpu_eager_baseline"] = {"error": repr(e)[:300]}
    if rank == 0 and world == 1 and not args.no_extras:
        try:
            line["other_configs"] = bench_other_configs(dev, peaks, with_llm=not args.no_llm)
        except Exception as e:          # secondary measurements must never cost the headline line
            import traceback
            traceback.print_exc(file=sys.stderr)
            line["other_configs"] = {"error": repr(e)[:300]}
    _trace("headline + e2e done")
    if not args.no_extras and not args.no_sweep:
        try:
            sw = bench_sweep(dev, rank, world, args, peaks)
            _trace("sweep done")
            if rank == 0:
                line.setdefault("other_configs", {})["configs4_dim_sweep"] = sw
        except Exception as e:
            if rank == 0:
                line.setdefault("other_configs", {})["configs4_dim_sweep"] = {"error": repr(e)[:300]}
    if world > 1 and not args.no_extras:
        try:
            dp = bench_dp_configs(dev, rank, world)
            if rank == 0:
                line.setdefault("other_configs", {}).update(dp)
        except Exception as e:
            if rank == 0:
                line.setdefault("other_configs", {})["dp_configs_error"] = repr(e)[:300]
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sps, dt, n_cpu = time_cpu_port(D, H, r, args.cpu_rows, 10, 3, best_of=True)
        line["cpu_baseline"] = {"value": sps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{args.cpu_rows} rows per step of the same workload, oracle port (torch CPU fp32 autograd), best of {n_cpu} steps after 3 warm-ups"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _trace(msg):
    """BENCH_TRACE=1: synchronise and print a stage marker to stderr (localises an asynchronous CUDA fault)"""
    if os.environ.get("BENCH_TRACE"):
        torch.cuda.synchronize()
        print(f"[bench trace rank {os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


def _time_fn(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def bench_other_configs(dev, peaks, with_llm=True):
    """Latency of the other BASELINE.json configs' hot-path shapes on one GPU (random-init modules of the real sizes, synthetic
    inputs, the LLM itself excluded -- it is the stock HF module in both implementations).  Times are CUDA-event ms per call."""
    import tempfile

    import numpy as np

    from dmi_b200 import augment as A
    from dmi_b200.model.hypernet import HyperNetWrapper
    from dmi_b200.model.projector import Projector
    from dmi_b200.utils.args import HypnetArgs, ProjectorArgs
    out = {}
    D, H, r = 768, 2048, 32
    torch.manual_seed(0)
    base = Projector(ProjectorArgs(proj_dropout=0.1), H, D, dev)
    with tempfile.NamedTemporaryFile(suffix=".pt") as f:
        torch.save({"projector_state_dict": base.state_dict()}, f.name)
        w = HyperNetWrapper(HypnetArgs(hn_arch="attention", hn_hypnet_dim=D, hn_rank=r, hn_alpha=32, hn_n_proj_layers=2, hn_use_pos_encs=True),
                            ProjectorArgs(proj_name_or_path=f.name), H, D, 128, dev)
    w.train()
    g = torch.Generator(device=dev).manual_seed(1)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)
    # ---- configs[1] v4:llama1b_inst_all micro-step: B=4, K=128, rotation, hypernet + projector as written (H1), fwd+bwd ----
    B, K = 4, 128
    mm, m, t, p = rn(B, D), rn(K, D), rn(K, D), rn(1, D)
    R = A.get_rotation_matrix(D, dev, random_state=np.random.RandomState(0))
    dy = rn(B, H) / math.sqrt(H)
    keep = (torch.rand(2, 2 + 1 + 2 * K, device=dev, generator=g) >= 0.05)

    def micro_step():
        x2, z = A.process_embeddings(mm, (m, t, p), R=R, normalize=True)
        a_w, b_w, biases = w.hypernet(z, keep_mask=keep, n_layers=1)          # what HyperNetWrapper.forward does on the as-written path
        y = w.projector.lora_forward_first_layer(x2, a_w[0], b_w[0], biases[0])
        y.backward(dy)
        return y.detach()
    gen_bytes = 4.0 * D * sum(gen.weight.shape[0] for gen in w.hypernet.generators)
    entry = {"what": "augment(normalise+3xTF32 rotation+interleave) + hypernet fwd (pooling, generator 0 -- the as-written projector never reads the second "
                     "adapter, so HyperNetWrapper.forward does not run generator 1) + lora_forward as written + backward (generator-0 rank-1 gradient "
                     "accumulated into .grad in place); LLM excluded"}
    # (1) the whole micro-step replayed as ONE CUDA graph (captured before any eager autograd graph of these parameters exists)
    try:
        from dmi_b200.graphs import GraphedStep
        w.hypernet.fuse_generator_grad_accumulation = True
        gs = GraphedStep(micro_step, dict(mm=mm, R=R), params=list(w.hypernet.parameters()))
        ms_g = _time_fn(lambda: gs(), reps=30)
        entry.update({"ms_cuda_graph": ms_g, "samples_per_s_cuda_graph": B / ms_g * 1e3})
        del gs
    except Exception as e:
        entry["cuda_graph_error"] = repr(e)[:200]
    # (1b) gradient accumulation the way the reference trains (GA micro-steps per optimizer step, train_hypernet.py:119-149) with the
    # generator gradient kept as rank-1 factors: the backward only READS G; the dense gradient is formed once per optimizer step.
    # A second wrapper instance: autograd's AccumulateGrad nodes of the first one belong to another capture stream.
    try:
        import traceback
        from dmi_b200.parallel import Rank1FactorSync
        GA = 5
        with tempfile.NamedTemporaryFile(suffix=".pt") as f:
            torch.save({"projector_state_dict": base.state_dict()}, f.name)
            w2 = HyperNetWrapper(HypnetArgs(hn_arch="attention", hn_hypnet_dim=D, hn_rank=r, hn_alpha=32, hn_n_proj_layers=2, hn_use_pos_encs=True),
                                 ProjectorArgs(proj_name_or_path=f.name), H, D, 128, dev)
        w2.train()
        gen0 = w2.hypernet.generators[0]
        w2.hypernet.fuse_generator_grad_accumulation = True
        sink = Rank1FactorSync(gen0.weight.shape[0], D, dev, max_terms=GA)
        w2.hypernet.factor_sinks = {0: sink}

        def ga_loop():
            sink.n = 0
            for _ in range(GA):
                x2, z = A.process_embeddings(mm, (m, t, p), R=R, normalize=True)
                a_w, b_w, biases = w2.hypernet(z, keep_mask=keep, n_layers=1)
                w2.projector.lora_forward_first_layer(x2, a_w[0], b_w[0], biases[0]).backward(dy)
        small = [q for n_, q in w2.hypernet.named_parameters() if not n_.startswith("generators")]
        gs = GraphedStep(ga_loop, dict(mm=mm, R=R), params=small)
        ms_ga = _time_fn(lambda: gs(), reps=20)
        gen0.weight.grad, gen0.bias.grad = torch.zeros_like(gen0.weight), torch.zeros_like(gen0.bias)

        def apply():
            sink.n = GA
            sink.apply_(gen0.weight.grad, gen0.bias.grad)
        ms_ap = _time_fn(apply, reps=10)
        entry.update({"ms_per_micro_step_cuda_graph_ga5_rank1_factors": ms_ga / GA, "ms_apply_factors_per_optimizer_step": ms_ap,
                      "ms_per_micro_step_amortised_rank1": (ms_ga + ms_ap) / GA})
        del gs, w2, sink
        torch.cuda.empty_cache()
    except Exception as e:
        entry["rank1_graph_error"] = repr(e)[:200]
        traceback.print_exc(file=sys.stderr)
    w.hypernet.factor_sinks = None
    for q in w.hypernet.parameters():
        q.grad = None
    w.hypernet.fuse_generator_grad_accumulation = False
    # (2) eager (host-bound: ~60 short launches + autograd per micro-step)
    ms = _time_fn(micro_step, reps=10)
    entry.update({"ms_eager": ms, "samples_per_s_eager": B / ms * 1e3})
    out["hypernet_microstep_B4_K128"] = entry
    with torch.no_grad():
        z = A.process_embeddings(mm, (m, t, p), R=None, normalize=True)[1]
        ms_f = _time_fn(lambda: w.hypernet(z), reps=20)
    out["hypernet_forward_K128"] = {"ms": ms_f, "achieved_gbs": gen_bytes / ms_f / 1e6, "frac_of_hbm_peak": gen_bytes / ms_f / 1e6 / peaks["hbm"],
                                     "what": "HyperNetwork.forward: pooling + both generator GEMVs (692 MB of fp32 weights streamed once)"}
    # ---- configs[2] few-shot: N adapters -> mean -> merge; merged MLP2 fwd+bwd with dW/db at B=256 ----
    zs = [A.process_embeddings(None, (rn(32, D), rn(32, D), rn(1, D)), R=None, normalize=True)[1] for _ in range(4)]
    w.eval()
    ms_g = _time_fn(lambda: w.generate_projector_from_multiple_adapters(zs), reps=5)
    out["fewshot_generate_4_adapters_and_merge"] = {"ms": ms_g}
    xb, dyb = rn(256, D), rn(256, H) / math.sqrt(H)
    gp = w.generated_projector

    def ft_step():
        for q in gp.parameters():
            q.grad = None
        gp(xb).backward(dyb)
    ms_ft = _time_fn(ft_step, reps=20)
    out["fewshot_merged_projector_B256_fwd_bwd"] = {"ms": ms_ft, "samples_per_s": 256 / ms_ft * 1e3}
    w.generated_projector = None
    # ---- configs[3] train_projector: plain MLP2 with dropout, B=1024 (global batch of the 8-GPU config on one GPU) ----
    base.train()
    x1k, dy1k = rn(1024, D), rn(1024, H) / math.sqrt(H)

    def proj_step():
        for q in base.parameters():
            q.grad = None
        base(x1k).backward(dy1k)
    ms_p = _time_fn(proj_step, reps=20)
    flops = (4 * D * H + 6 * H * H) * 1024
    out["train_projector_B1024_fwd_bwd"] = {"ms": ms_p, "samples_per_s": 1024 / ms_p * 1e3, "tflops": flops / ms_p / 1e9,
                                            "what": "Projector.forward + backward through the module API, eager, gradients through autograd"}
    try:
        from dmi_b200.graphs import GraphedStep
        from dmi_b200.model.mlp2 import plain_mlp2
        from dmi_b200.model.projector import Projector as _P
        torch.manual_seed(0)
        pg = _P(ProjectorArgs(proj_dropout=0.1), H, D, dev)          # fresh parameters: their gradients are only ever written by this capture
        pg.train()
        for q in pg.parameters():
            q.grad = torch.zeros_like(q)
        l0, l1 = pg.net[0], pg.net[3]
        gsp = GraphedStep(lambda: plain_mlp2(x1k, l0.weight, l0.bias, l1.weight, l1.bias, dropout_p=0.1, cache=None, grad_in_place=True).backward(dy1k), {}, params=[])
        ms_pg = _time_fn(gsp, reps=50)
        out["train_projector_B1024_fwd_bwd"].update(ms_cuda_graph_in_place_grads=ms_pg, samples_per_s_cuda_graph=1024 / ms_pg * 1e3, tflops_cuda_graph=flops / ms_pg / 1e9)
        del gsp, pg
    except Exception as e:          # noqa: BLE001
        import traceback
        traceback.print_exc(file=sys.stderr)
        out["train_projector_B1024_fwd_bwd"]["graph_error"] = repr(e)
    # ---- optimizer step over the hypernet's 175 M parameters: fused clip-grad-norm + AdamW vs torch's own ----
    try:
        from dmi_b200.optim import FusedAdamW
        hp = dict(lr=1e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=5e-6)
        hparams = [q for q in w.hypernet.parameters()]
        for q in hparams:
            q.grad = torch.randn_like(q) * 1e-3
        n_par = sum(q.numel() for q in hparams)
        fo = FusedAdamW(hparams, **hp)
        ms_o = _time_fn(lambda: fo.step(max_grad_norm=1.0), reps=10)
        del fo
        to = torch.optim.AdamW(hparams, **hp)

        def torch_step():
            torch.nn.utils.clip_grad_norm_(hparams, 1.0)
            to.step()
        ms_t = _time_fn(torch_step, reps=5)
        del to
        for q in hparams:
            q.grad = None
        out["optimizer_step_hypernet"] = {"params": n_par, "ms_fused_clip_adamw": ms_o, "achieved_gbs": 32.0 * n_par / ms_o / 1e6,
                                          "frac_of_hbm_peak": 32.0 * n_par / ms_o / 1e6 / peaks["hbm"], "ms_torch_clip_plus_adamw": ms_t,
                                          "what": "clip_grad_norm_(1.0) + AdamW over all hypernet parameters: 4 B/param norm pass + 28 B/param update pass"}
    except Exception as e:
        out["optimizer_step_hypernet"] = {"error": repr(e)[:200]}
    # ---- configs[1] with the LLM: random-init Llama-3.2-1B (bf16), B=4, T=320: hot path vs whole micro-step ----
    if with_llm:
        try:
            out["v4_microstep_with_llama1b"] = bench_llm_microstep(dev, w, A, mm, (m, t, p), R)
        except Exception as e:
            out["v4_microstep_with_llama1b"] = {"error": repr(e)[:300]}
    # ---- few-shot: 16 support sets -> mean adapter through ONE generator pass (8f-2) vs 16 separate hypernet passes ----
    try:
        w.eval()
        zs16 = [A.process_embeddings(None, (rn(32, D), rn(32, D), rn(1, D)), R=None, normalize=True)[1] for _ in range(16)]
        with torch.no_grad():
            ms_mean = _time_fn(lambda: w.hypernet.mean_adapter(zs16), reps=5)
            ms_sep = _time_fn(lambda: [w.hypernet(zz) for zz in zs16], reps=3)
        out["fewshot_mean_adapter_16_sets"] = {"ms_one_generator_pass": ms_mean, "ms_16_hypernet_passes": ms_sep}
        w.train()
    except Exception as e:
        out["fewshot_mean_adapter_16_sets"] = {"error": repr(e)[:200]}
    # ---- isometry draw on the device (8f-3) and embedding-store gather (8f-4) ----
    try:
        gen = torch.Generator(device=dev).manual_seed(5)
        ms_h = _time_fn(lambda: A.get_rotation_matrix_device(D, dev, generator=gen), reps=10)
        out["isometry_draw_device_D768"] = {"ms": ms_h, "what": "Haar orthogonal 768x768 from Gaussian samples (compact-WY Householder, fp32)"}
        from dmi_b200.data import EmbeddingStore
        table = rn(400000, D)
        store = EmbeddingStore(table, mean=rn(D) * 0.01)
        sidx = torch.randint(0, 400000, (32768,), device=dev, generator=g)
        sout = torch.empty(32768, D, device=dev)
        sbf = torch.empty(32768, D, device=dev, dtype=torch.bfloat16)
        ms_g = _time_fn(lambda: store.gather(sidx, out=sout, out_bf16=sbf), reps=20)
        gb = 32768 * D * (4 + 4 + 2)
        out["embedding_store_gather_B32768"] = {"ms": ms_g, "achieved_gbs": gb / ms_g / 1e6, "frac_of_hbm_peak": gb / ms_g / 1e6 / peaks["hbm"],
                                                "what": "random row gather from a 1.2 GB fp32 table + mean subtraction + L2 normalise -> fp32 and bf16 batch"}
        del table, store
    except Exception as e:
        out["isometry_draw_device_D768"] = {"error": repr(e)[:200]}
    # ---- splice: B=32, T=320, fp32 out (reference promotion) and bf16 out ----
    from dmi_b200.model.mmmodel import splice_prefix
    table = rn(128256, H).to(torch.bfloat16)
    ids = torch.randint(0, 128256, (32, 320), device=dev, generator=g)
    proj = rn(32, H)
    for name, dt_, ob in (("splice_B32_T320_fp32", torch.float32, 4), ("splice_B32_T320_bf16", torch.bfloat16, 2)):
        ms_s = _time_fn(lambda: splice_prefix(proj, table, ids, None, None, dt_), reps=30)
        nbytes = 32 * 321 * H * (2 + ob)
        out[name] = {"ms": ms_s, "achieved_gbs": nbytes / ms_s / 1e6, "frac_of_hbm_peak": nbytes / ms_s / 1e6 / peaks["hbm"]}
    return out


def bench_llm_microstep(dev, w, A, mm, support, R):
    """BASELINE configs[1] (train_hypernet v4:llama1b_inst_all shape): one micro-step through HypernetMMModel with a random-init
    Llama-3.2-1B-shaped LLM in bf16 (no checkpoint available offline), B=4, K=128, T=320.  Reports the whole micro-step and the
    LLM-only forward+backward so that the hot path's share is visible."""
    from transformers import LlamaConfig, LlamaForCausalLM
    from dmi_b200.model.mmmodel import HypernetMMModel
    cfg = LlamaConfig(hidden_size=2048, num_hidden_layers=16, num_attention_heads=32, num_key_value_heads=8, intermediate_size=8192,
                      vocab_size=128256, tie_word_embeddings=True, rms_norm_eps=1e-5, rope_theta=500000.0, max_position_embeddings=4096)
    with torch.device(dev):
        llm = LlamaForCausalLM(cfg).to(torch.bfloat16)
    model = HypernetMMModel(llm, w, "cuda", 768, "bench", 0)      # device spelled like the reference configs: autocast is keyed on it (mmmodel.py:53)
    model.train()
    B, T = mm.shape[0], 320
    g = torch.Generator(device=dev).manual_seed(9)
    ids = torch.randint(0, 128256, (B, T), device=dev, generator=g)
    mask = torch.ones(B, T, device=dev, dtype=torch.int64)

    def micro_step():
        for q in w.hypernet.parameters():
            q.grad = None
        x2, z = A.process_embeddings(mm, support, R=R, normalize=True)
        loss, _ = model(x2, z, ids, mask, ids)
        loss.backward()

    emb = torch.randn(B, 1 + T, 2048, device=dev).requires_grad_(True)
    lab = torch.cat([torch.full((B, 1), -100, device=dev, dtype=torch.int64), ids], 1)

    def llm_only():
        emb.grad = None
        with torch.amp.autocast("cuda"):
            llm(inputs_embeds=emb, labels=lab).loss.backward()

    # the hot path's share is a difference of two ~25 ms numbers: interleave the two measurements and keep the fastest round of each
    tot, llm_t = [], []
    for _ in range(3):
        tot.append(_time_fn(micro_step, reps=3, warm=1))
        llm_t.append(_time_fn(llm_only, reps=3, warm=1))
    ms_total, ms_llm = min(tot), min(llm_t)
    del model, llm
    torch.cuda.empty_cache()
    return {"ms_micro_step_total": ms_total, "ms_llm_fwd_bwd_only": ms_llm, "ms_hot_path_and_glue": ms_total - ms_llm,
            "what": "HypernetMMModel.forward + backward, random-init Llama-3.2-1B shape (16 layers, hidden 2048, vocab 128256) bf16 autocast, "
                    "B=4 K=128 T=320; the LLM is the stock HF module in both implementations"}


def gpu_eager_baseline(dev, B, D, H, r, w1, b1, w2, b2, adapter, x, dy):
    """The reference's op sequence for this workload (Projector.only_lora_forward / combine_lora math: F.linear + two skinny matmuls
    per layer + GELU(tanh), gradients to the adapter factors by autograd; dmi/model/projector.py:61-74) run by PyTorch eager on THIS
    GPU -- the honest software baseline.  fp32 as the reference runs it (TF32 off), fp32 with
    TF32 matmuls allowed, and bf16 autocast.  Bench-side only: nothing of this is in the package."""
    import torch.nn.functional as Fn
    A0, B0, be0, A1, B1, be1 = [t.detach().clone().requires_grad_(True) for t in adapter]
    leaves = [A0, B0, be0, A1, B1, be1]

    def step(dtype):
        with torch.autocast("cuda", dtype=dtype, enabled=dtype is not None):
            pre = Fn.linear(x, w1, b1) + (x @ A0.view(D, r)) @ B0.view(r, H) + be0
            h = Fn.gelu(pre, approximate="tanh")
            yy = Fn.linear(h, w2, b2) + (h @ A1.view(H, r)) @ B1.view(r, H) + be1
        return torch.autograd.grad(yy, leaves, dy.to(yy.dtype))
    out = {"what": "torch eager on the same B200: F.linear + (x@A)@B + gelu(tanh) per layer, autograd to A/B/beta (frozen base); same shapes and inputs"}
    old = torch.backends.cuda.matmul.allow_tf32
    try:
        for name, tf32, dt in (("fp32", False, None), ("fp32_tf32_matmul", True, None), ("bf16_autocast", False, torch.bfloat16)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            ms = _time_fn(lambda: step(dt), reps=5 if name == "fp32" else 10, warm=2)
            out[name] = {"ms_per_step": ms, "samples_per_s": B / ms * 1e3}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    return out


def _time_adapted_step(dev, B, D, H, r, steps=20, flags=0, world=1, reducer=None):
    """one configs[4] point: full (or as-written) adapted MLP2 fwd+bwd at width D, rank r; returns ms per step (max over ranks)"""
    from dmi_b200 import ops
    from dmi_b200.parallel import FlatGrads
    g = torch.Generator(device=dev).manual_seed(D * 131 + r)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)
    w1, w2 = rn(H, D) / math.sqrt(D), rn(H, H) / math.sqrt(H)
    b1 = b2 = torch.zeros(H, device=dev)
    ad = (rn(D * r) / math.sqrt(D), rn(r * H) * 0.1, torch.zeros(H, device=dev), rn(H * r) / math.sqrt(H), rn(r * H) * 0.1, torch.zeros(H, device=dev))
    xs = [torch.nn.functional.normalize(rn(B, D), dim=1) for _ in range(2)]
    dys = [rn(B, H) / math.sqrt(H) for _ in range(2)]
    pk = ops.PackedProjector(D, H, r, dev)
    pk.pack_base(w1, w2)
    full = not (flags & 1)
    st = ops.MlpStash(B, D, H, r, dev, full=full)
    y = torch.empty(B, H, device=dev)
    shapes = dict(dA0=(D, r), dB0=(r, H), dbeta0=(H,))
    if full:
        shapes.update(dA1=(H, r), dB1=(r, H), dbeta1=(H,))
    gb = [FlatGrads(shapes, [list(shapes)], dev) for _ in range(2)]
    done = [None, None]

    def step(i):
        k = i & 1
        if done[k] is not None:
            torch.cuda.current_stream().wait_event(done[k])
            done[k] = None
        gb[k].zero_()
        pk.pack_adapter(*ad[:3], *(ad[3:] if full else (None, None, None)), b1, b2 if full else None)
        ops.adapted_mlp_fwd(pk, st, xs[k], y, flags=flags)
        ops.adapted_mlp_bwd(pk, st, dys[k], gb[k].views, flags=flags, grad_scale=1.0 / world)
        if reducer is not None:
            reducer.reduce_bucket(gb[k].flat, None)
            done[k] = reducer.done_event()
    for i in range(3):
        step(i)
    if reducer is not None:
        reducer.wait()
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    if reducer is not None:
        reducer.wait()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    del pk, st, gb, xs, dys
    torch.cuda.empty_cache()
    return ms


def bench_sweep(dev, rank, world, args, peaks):
    """BASELINE configs[4]: arbitrary encoder dim sweep (512..4096-d inputs, LoRA rank 8-64), adapted-projector fwd+bwd, args.batch rows
    per GPU.  Reading (i) kernel capacity: projector / adapter built natively at width D.  Reading (ii) reference-faithful: the
    hypernet stays 768 wide, so D < 768 is `proj_prune = D` (same kernels as (i) at that D: x is D wide, W1[:, :D], A0 = first D rows;
    train_hypernet.py:465-472, hypernet.py:187-188) and D > 768 is the static InfFS column gather emb[selected_features] down to 768
    (data/base.py:222-225) in the embedding-store prologue followed by the D = 768 step.
    N = 1: the whole grid (reduced unless --sweep); N > 1: the centre point's two corners with the gradient all-reduce."""
    from dmi_b200.parallel import BucketAllReducer
    H, B = args.H, args.batch
    if world > 1:
        pts = [(512, 8), (4096, 64)]
    elif args.sweep:
        pts = [(D, r) for D in (512, 640, 1024, 2048, 4096) for r in (8, 16, 32, 64)]
    else:
        pts = [(512, 8), (512, 32), (512, 64), (640, 32), (1024, 8), (1024, 32), (1024, 64), (2048, 32), (2048, 64), (4096, 8), (4096, 32), (4096, 64)]
    reducer = BucketAllReducer(average=False) if world > 1 else None
    out = {"rows_per_gpu": B, "H": H, "n_gpus": world, "points": []}
    for D, r in pts:
        ms = _time_adapted_step(dev, B, D, H, r, steps=10, world=world, reducer=reducer)
        F = flops_per_sample(D, H, r)
        sps = world * B / ms * 1e3
        out["points"].append({"D": D, "r": r, "ms_per_step": ms, "samples_per_s": sps, "tflops_per_gpu": sps / world * F / 1e12,
                              "frac_of_bf16_burst_peak": sps / world * F / 1e12 / peaks["bf16_burst"], "mflop_per_sample": F / 1e6})
    if world == 1:
        # reading (ii) for D > 768: InfFS column gather (static index gather + mean subtraction + L2 normalise) to 768, then the centre step
        from dmi_b200.data import EmbeddingStore
        ms768 = _time_adapted_step(dev, B, 768, H, 32, steps=10)
        g = torch.Generator(device=dev).manual_seed(3)
        faithful = []
        for D in (1024, 2048, 4096):
            table = torch.randn(65536, D, device=dev, generator=g)
            sel = torch.randperm(D, device=dev, generator=g)[:768].sort().values.cpu().numpy()
            store = EmbeddingStore(table, selected_features=sel)
            idx = torch.randint(0, 65536, (B,), device=dev, generator=g)
            buf = torch.empty(B, 768, device=dev, dtype=torch.bfloat16)
            ms_g = _time_fn(lambda: store.gather(idx, out_bf16=buf, want_f32=False), reps=10)
            faithful.append({"D_encoder": D, "ms_gather_to_768": ms_g, "ms_step_at_768": ms768, "samples_per_s": B / (ms_g + ms768) * 1e3})
            del table, store
        out["reference_faithful_D_gt_768"] = faithful
        # configs[4] batch axis at the centre point (D = 768, r = 32): 256 .. 65536 rows per GPU.  Below 8192 rows the schedule is the 1-CTA
        # GEMM + mma.sync side passes (api.cu: use_pair / use_panel_tc) and a step is a dozen launches of a few microseconds each.
        F = flops_per_sample(768, H, 32)
        rows = []
        for Bs in (256, 1024, 4096, 16384, 65536):
            ms = _time_adapted_step(dev, Bs, 768, H, 32, steps=20 if Bs <= 4096 else 10)
            rows.append({"rows": Bs, "ms_per_step": ms, "samples_per_s": Bs / ms * 1e3, "frac_of_bf16_burst_peak": Bs / ms * 1e3 * F / 1e12 / peaks["bf16_burst"]})
        out["batch_sweep_D768_r32"] = rows
        out["batch_sweep_note"] = "eager launches, 2 rotating input buffers: up to 16384 rows the working set fits the 126 MB L2 (L2-warm numbers)"
        out["reference_faithful_D_lt_768"] = "identical to the kernel-capacity points at D = 512 / 640 (proj_prune): see points"
    return out


def bench_dp_configs(dev, rank, world):
    """N > 1 only.  (1) BASELINE configs[3]: train_projector v1 shape, plain MLP2 768->2048->2048 with dropout 0.1, GLOBAL batch 1024
    (1024 / world rows per GPU), gradients of W1,b1,W2,b2 (23 MB) all-reduced bucket by bucket as autograd produces them (GradSync),
    then clip + AdamW on every rank (train_projector.py:51-73).  (2) the hypernet path of configs[1] under DP with the reference's
    gradient-accumulation semantics (train_hypernet.py:119-149): one micro-step per rank per GA slot, GA_local = 5 (v4: 40 = 8 x 5),
    hypernet gradients synchronised once per optimizer step either as a dense bucketed all-reduce overlapped with the backward or
    through the rank-1 factor all-gather (Rank1FactorSync) for the generator weights.  LLM excluded (stock HF module)."""
    import tempfile

    import numpy as np

    from dmi_b200 import augment as A
    from dmi_b200.model.hypernet import HyperNetWrapper
    from dmi_b200.model.projector import Projector
    from dmi_b200.optim import FusedAdamW
    from dmi_b200.parallel import GradSync, Rank1FactorSync
    from dmi_b200.utils.args import HypnetArgs, ProjectorArgs
    out = {}
    D, H, r = 768, 2048, 32

    def timed(fn, reps, warm=3):
        for _ in range(warm):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    # ---- (1) configs[3] ----
    from dmi_b200.graphs import GraphedStep
    rows = max(1, 1024 // world)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    x, dy = torch.randn(rows, D, device=dev, generator=g), torch.randn(rows, H, device=dev, generator=g) / math.sqrt(H) / world
    hp = dict(lr=1e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=5e-6)
    tp = {"rows_per_gpu": rows, "n_gpus": world, "grad_bytes": 0,
          "what": "Projector.forward (dropout 0.1, mask drawn on the device) + backward replayed as one CUDA graph with the gradients accumulated into flat "
                  "bucket views, NCCL all-reduce of dW2,db2,dW1,db1 (23 MB), fused clip + AdamW (train_projector.py:51-73); global batch 1024"}
    for mode in ("allreduce", "no_allreduce"):
        _trace(f"train_projector {mode}")
        torch.manual_seed(0)
        proj = Projector(ProjectorArgs(proj_dropout=0.1), H, D, dev)      # a fresh module per mode: one capture stream per set of parameters
        proj.train()
        params = [proj.net[3].weight, proj.net[3].bias, proj.net[0].weight, proj.net[0].bias]        # backward-availability order
        sync = GradSync(params, bucket_bytes=8 << 20)
        opt = FusedAdamW(params, **hp)
        from dmi_b200.model.mlp2 import plain_mlp2
        l0, l1 = proj.net[0], proj.net[3]
        # cache=None: the fp32 -> bf16 operand pack of the (just updated) weights is part of every captured step, as in real training
        # grad_in_place: dW, db are accumulated straight into the bucket views (no zero-filled temporaries, no AccumulateGrad adds)
        gs = GraphedStep(lambda: plain_mlp2(x, l0.weight, l0.bias, l1.weight, l1.bias, dropout_p=0.1, cache=None, grad_in_place=True).backward(dy), {}, params=[])

        def train_step():
            sync.zero_grad()
            gs()
            if mode == "allreduce":
                for bkt in sync.buckets:
                    sync.reducer.reduce_bucket(bkt, None)
                sync.reducer.wait()
            opt.step(max_grad_norm=1.0)
        ms = timed(train_step, 30)
        tp["grad_bytes"] = sum(q.numel() for q in params) * 4
        tp["ms_per_step_" + mode] = ms
        if mode == "allreduce":
            tp["samples_per_s"] = rows * world / ms * 1e3
        sync.remove()
        del opt, gs, sync
    out["train_projector_B1024_global_dp"] = tp
    _trace("train_projector done")
    # ---- (2) hypernet path under DP with GA semantics ----
    from dmi_b200.graphs import GraphedStep
    Bm, K, GA_local = 4, 128, 5
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)
    mm, m, t, p = rn(Bm, D), rn(K, D), rn(K, D), rn(1, D)
    R = A.get_rotation_matrix(D, dev, random_state=np.random.RandomState(rank))
    dyh = rn(Bm, H) / math.sqrt(H) / (world * GA_local)
    keep = (torch.rand(2, 3 + 2 * K, device=dev, generator=g) >= 0.05)
    entry = {"n_gpus": world, "micro_steps_per_rank": GA_local, "micro_batch": Bm, "support": K,
             "what": "per optimizer step: GA_local micro-steps per rank replayed as ONE CUDA graph (augment + hypernet fwd + lora_forward as written + "
                     "backward, gradients accumulated in place into flat bucket views), hypernet-gradient synchronisation over NCCL, fused clip + AdamW on "
                     "every rank; equals the reference with GA = world x GA_local (train_hypernet.py:119-149).  dense_allreduce: 291 MB of fp32 buckets; "
                     "rank1_factors: the generator gradient travels as (dw, e) pairs (all-gather) and is rebuilt locally, only the 7 MB of pooling "
                     "gradients are all-reduced"}
    for mode in ("dense_allreduce", "rank1_factors", "no_sync"):
        _trace(f"hypernet dp {mode}")
        factors = mode == "rank1_factors"
        with tempfile.NamedTemporaryFile(suffix=".pt") as f:          # a fresh wrapper per mode: one capture stream per set of parameters
            torch.save({"projector_state_dict": proj.state_dict()}, f.name)
            w = HyperNetWrapper(HypnetArgs(hn_arch="attention", hn_hypnet_dim=D, hn_rank=r, hn_alpha=32, hn_n_proj_layers=2, hn_use_pos_encs=True),
                                ProjectorArgs(proj_name_or_path=f.name), H, D, 128, dev)
        w.train()
        hn = w.hypernet
        gen0 = hn.generators[0]
        others = [q for n, q in hn.named_parameters() if not n.startswith("generators")]
        hparams = list(gen0.parameters()) + others                       # H1: generators.1 gets no gradient
        sync = GradSync(others if factors else hparams, bucket_bytes=64 << 20)
        if factors:
            gen0.weight.grad, gen0.bias.grad = torch.zeros_like(gen0.weight), torch.zeros_like(gen0.bias)
        opt = FusedAdamW(hparams, **hp)
        hn.fuse_generator_grad_accumulation = True                        # every gradient is accumulated in place: no AccumulateGrad nodes
        sink = Rank1FactorSync(gen0.weight.shape[0], D, dev, max_terms=GA_local) if factors else None
        hn.factor_sinks = {0: sink} if factors else None

        def ga_loop():
            if sink is not None:
                sink.n = 0
            for _ in range(GA_local):
                x2, z = A.process_embeddings(mm, (m, t, p), R=R, normalize=True)
                a_w, b_w, biases = hn(z, keep_mask=keep, n_layers=1)
                w.projector.lora_forward_first_layer(x2, a_w[0], b_w[0], biases[0]).backward(dyh)
        gs = GraphedStep(ga_loop, dict(mm=mm, R=R), params=[])

        def opt_step():
            sync.zero_grad()
            if factors:
                gen0.weight.grad.zero_()
                gen0.bias.grad.zero_()
            gs()
            if factors:
                sink.n = GA_local
                sink.apply_(gen0.weight.grad, gen0.bias.grad)             # all-gather of (dw, e) + local rank-(world*GA) update
            if mode != "no_sync":
                for bkt in sync.buckets:
                    sync.reducer.reduce_bucket(bkt, None)
                sync.reducer.wait()
            opt.step(max_grad_norm=1.0)
        ms = timed(opt_step, 10, warm=3)
        entry["ms_per_optimizer_step_" + mode] = ms
        entry["ms_per_micro_step_" + mode] = ms / GA_local
        if mode == "dense_allreduce":
            entry["wire_bytes_per_rank_dense"] = sum(q.numel() for q in hparams) * 4
            entry["wire_bytes_per_rank_rank1"] = GA_local * (gen0.weight.shape[0] + D) * 4 + sum(q.numel() for q in others) * 4
        sync.remove()
        del opt, gs, sync, sink, w, hn, gen0, others, hparams
        torch.cuda.empty_cache()
    out["hypernet_path_dp_ga"] = entry
    _trace("hypernet dp done")
    return out


def kernel_breakdown(ops, pk, st, y, xs, dys, gbuf, B, D, H, r, peaks, step_ms):
    """time each kernel of the step alone (CUDA events, 20 launches after 3 warm-ups, rotating inputs)"""
    KX, KH = D + r, H + r

    def timeit(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    dyext, dpre = st.dyext, st.dpre
    kernels = {}
    kernels["gemm_fwd_layer0_gelu  [B,D+r]x[H,D+r]"] = (
        timeit(lambda: ops.gemm_tn(st.xext, pk.w1ext, mode=ops.EPI_GELU, bias=pk.bias0, out0=st.hext[:, :H], out1=st.pre)), 2.0 * B * H * KX)
    kernels["gemm_fwd_layer1_store [B,H+r]x[H,H+r]"] = (
        timeit(lambda: ops.gemm_tn(st.hext, pk.w2ext, bias=pk.bias1, out0=y)), 2.0 * B * H * KH)
    kernels["gemm_bwd_dpre_gelugrad [B,H+r]x[H,H+r]"] = (
        timeit(lambda: ops.gemm_tn(dyext, pk.w2text, mode=ops.EPI_GELU_BWD, out0=dpre, aux=st.pre)), 2.0 * B * H * KH)
    kernels["gemm_skinny_u [B,D]x[r,D]"] = (timeit(lambda: ops.gemm_tn(st.xext[:, :D], pk.a0t, out0=st.xext[:, D:])), 2.0 * B * r * D)
    kernels["gemm_skinny_v [B,H]x[r,H]"] = (timeit(lambda: ops.gemm_tn(st.hext[:, :H], pk.a1t, out0=st.hext[:, H:])), 2.0 * B * r * H)
    kernels["outer_reduce dB1 [r,H]"] = (
        timeit(lambda: ops.outer_reduce(st.hext[:, H:], dyext[:, :H], gbuf["dB1"], colsum=gbuf["dbeta1"])), 2.0 * B * r * H)
    rows = []
    for name, (ms, fl) in kernels.items():
        rows.append({"kernel": name, "ms": ms, "tflops": fl / ms / 1e9, "frac_of_burst_peak": fl / ms / 1e9 / peaks["bf16_burst"]})
    # the fused dpre pass (du, dB0, dbeta0 in one tcgen05 sweep; the step's default from 8192 rows up) is HBM-bound: report GB/s.
    # Informational only -- never allowed to break the bench line.
    try:
        if B >= 8192 and H in (1024, 2048) and r == 32:
            ms = timeit(lambda: ops.panel_fused_tc(dpre, pk.b0, st.xext[:, D:], st.du[:, :r], gbuf["dB0"], colsum=gbuf["dbeta0"]))
            gbs = 2.0 * B * H / ms / 1e6
            rows.append({"kernel": "panel_tc_kernel dpre pass (du, dB0, dbeta0) [B,H] bf16, one sweep", "ms": ms, "achieved_gbs": gbs,
                         "frac_of_hbm_peak": gbs / peaks["hbm"]})
    except Exception as e:      # pragma: no cover
        rows.append({"kernel": "panel_tc_kernel dpre pass", "error": str(e)[:200]})
    top_name, (top_ms, top_fl) = max(list(kernels.items())[:3], key=lambda kv: kv[1][0])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                per_row = json.load(f).get("bytes_per_row", {}).get(top_name.split()[0])
                traffic = None if per_row is None else float(per_row) * B
        except Exception:
            traffic = None
    achieved = top_fl / top_ms / 1e9
    roof = {"bound": "tensor", "kernel": "gemm_tn_kernel<256> " + top_name, "achieved": achieved, "peak": peaks["bf16_burst"],
            "unit": "TFLOP/s", "frac": achieved / peaks["bf16_burst"], "traffic": traffic,
            "traffic_source": "static: per-row DRAM bytes of this kernel from the ncu --set full capture listed in profiles/roofline_traffic.json x rows of this run (not measured live)",
            "peak_source": peaks["source"] + " bf16 burst (kernel timed alone)",
            "share_of_step": top_ms / step_ms}
    return {"roofline": roof, "kernels": rows}


def run_e2e(args, dev, rank, world, w1, b1, w2, b2, adapter, host_dtype=torch.bfloat16):
    """Same metric through the public module API (Projector.lora_forward in 'full' mode + autograd) with HOST inputs:
    every step copies its batch from pinned host memory (prefetched one step ahead on a copy stream), runs
    forward + loss + backward (+ gradient all-reduce) and reads the loss back to the host."""
    from dmi_b200.model.projector import Projector
    from dmi_b200.parallel import allreduce_module_grads
    from dmi_b200.utils.args import ProjectorArgs
    B, D, H, r = args.batch, args.D, args.H, args.r
    proj = Projector(ProjectorArgs(proj_dropout=0.0), H, D, dev)
    with torch.no_grad():
        proj.net[0].weight.copy_(w1); proj.net[0].bias.copy_(b1); proj.net[3].weight.copy_(w2); proj.net[3].bias.copy_(b2)
    proj.eval()
    for p in proj.parameters():
        p.requires_grad_(False)
    proj.lora_forward_mode = "full"
    leaves = [t.clone().requires_grad_(True) for t in adapter]
    A0, B0, be0, A1, B1, be1 = leaves
    G = torch.randn(B, H, device=dev) / math.sqrt(H)
    Gflat = G.reshape(-1)
    NB = 3
    hosts = []
    for i in range(NB):
        x = torch.randn(B, D)
        hosts.append((x / x.norm(dim=1, keepdim=True)).to(host_dtype).pin_memory())      # embedding store on the host (bf16, or fp32 = the reference's mm_dtype)
    copy_stream = torch.cuda.Stream()
    # H2D lands directly in columns [0,D) of the projector's bf16 operand buffer [B, D+r]: no device-side convert / copy
    f32_host = host_dtype == torch.float32
    dev_bufs = [torch.zeros(B, D if f32_host else D + r, device=dev, dtype=host_dtype) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i & 1])
            dev_bufs[i & 1][:, :D].copy_(hosts[i % NB], non_blocking=True)
            ready[i & 1].record(copy_stream)

    def fwd_loss_bwd(buf):
        x = buf[:, :D]
        yy = proj.lora_forward(x, [A0, A1], [B0, B1], [be0, be1])
        loss = torch.dot(yy.reshape(-1), Gflat)          # synthetic scalar loss whose gradient is the fixed upstream dY = G
        return loss, torch.autograd.grad(loss, leaves)

    # The step (module forward + loss + autograd backward) is recorded once per input buffer as a CUDA graph
    # (dmi_b200.graphs: every entry point of the library is capture-safe) and replayed; eager execution is the fallback.
    graphs = [None, None]
    if not args.e2e_eager:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for b in dev_bufs:
                    for _ in range(2):
                        fwd_loss_bwd(b)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            for k in range(2):
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph, stream=side):
                    outs = fwd_loss_bwd(dev_bufs[k])
                graphs[k] = (gph, outs)
            torch.cuda.synchronize()
        except Exception as e:           # capture is an optimisation, never a requirement
            graphs = [None, None]
            sys.stderr.write("bench.py: e2e CUDA-graph capture failed, running eagerly: %r\n" % (e,))

    def e2e_step(i):
        torch.cuda.current_stream().wait_event(ready[i & 1])
        prefetch(i + 1)
        if graphs[i & 1] is not None:
            gph, (loss, grads) = graphs[i & 1]
            gph.replay()
        else:
            loss, grads = fwd_loss_bwd(dev_bufs[i & 1])
        consumed[i & 1].record()
        for t, g_ in zip(leaves, grads):
            t.grad = g_
        allreduce_module_grads(leaves)
        # device -> host read of the step result, pipelined: the copy of step i is enqueued now and its value is read while
        # step i+1 is already running (every step's loss still reaches the host inside the timed region)
        loss_host[i & 1].copy_(loss.detach().reshape(1), non_blocking=True)
        loss_ready[i & 1].record()
        if i > 0:
            loss_ready[(i - 1) & 1].synchronize()
            losses.append(float(loss_host[(i - 1) & 1][0]))

    for e in consumed:
        e.record()
    loss_host = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ready = [torch.cuda.Event() for _ in range(2)]
    losses = []
    steps = max(5, min(args.steps, 50))
    prefetch(0)
    for i in range(3):
        e2e_step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(3, 3 + steps):
        e2e_step(i)
    loss_ready[(2 + steps) & 1].synchronize()          # the last step's loss
    losses.append(float(loss_host[(2 + steps) & 1][0]))
    t1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank != 0:
        return None
    return {"value": world * B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
            "h2d_bytes_per_step": B * D * hosts[0].element_size(), "d2h_bytes_per_step": 4, "h2d_gbs": B * D * hosts[0].element_size() / ms / 1e6,
            "host_dtype": "f32 (the reference's mm_dtype)" if f32_host else "bf16 (half the bytes of the reference's fp32 embeddings; numerically identical to converting on the device, since the GEMM operand is bf16 either way)",
            "api": "dmi_b200.model.Projector.lora_forward(mode='full') + torch.autograd backward" + (" (captured once per input buffer as a CUDA graph and replayed)" if graphs[0] is not None else " (eager)") + "; x = embeddings copied from pinned host memory every step (prefetched one step ahead on a copy stream), loss read back to the host every step"}


def run_e2e_store(args, dev, rank, world, w1, b1, w2, b2, adapter):
    """End-to-end variant with the embedding table resident in HBM (dmi_b200.data.EmbeddingStore): the per-step host
    input is the batch of SAMPLE INDICES (int64, pinned memory -> device every step); the gather + L2 normalisation, the module
    forward, the loss and the autograd backward are replayed as one CUDA graph per input buffer; the loss is read back every step.
    Reported next to `e2e` (which copies the embeddings themselves from the host every step), not instead of it."""
    from dmi_b200.data import EmbeddingStore
    from dmi_b200.model.projector import Projector
    from dmi_b200.parallel import allreduce_module_grads
    from dmi_b200.utils.args import ProjectorArgs
    B, D, H, r = args.batch, args.D, args.H, args.r
    proj = Projector(ProjectorArgs(proj_dropout=0.0), H, D, dev)
    with torch.no_grad():
        proj.net[0].weight.copy_(w1); proj.net[0].bias.copy_(b1); proj.net[3].weight.copy_(w2); proj.net[3].bias.copy_(b2)
    proj.eval()
    for p in proj.parameters():
        p.requires_grad_(False)
    proj.lora_forward_mode = "full"
    leaves = [t.clone().requires_grad_(True) for t in adapter]
    A0, B0, be0, A1, B1, be1 = leaves
    Gflat = (torch.randn(B, H, device=dev) / math.sqrt(H)).reshape(-1)
    n_table = 262144
    store = EmbeddingStore(torch.randn(n_table, D, device=dev).to(torch.bfloat16))        # 400 MB bf16 table in HBM
    NB = 3
    host_idx = [torch.randint(0, n_table, (B,), dtype=torch.int64).pin_memory() for _ in range(NB)]
    copy_stream = torch.cuda.Stream()
    idx_dev = [torch.zeros(B, dtype=torch.int64, device=dev) for _ in range(2)]
    dev_bufs = [torch.zeros(B, D + r, device=dev, dtype=torch.bfloat16) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i & 1])
            idx_dev[i & 1].copy_(host_idx[i % NB], non_blocking=True)
            ready[i & 1].record(copy_stream)

    def gather_fwd_loss_bwd(k):
        store.gather(idx_dev[k], normalize=True, out_bf16=dev_bufs[k][:, :D], want_f32=False)
        yy = proj.lora_forward(dev_bufs[k][:, :D], [A0, A1], [B0, B1], [be0, be1])
        loss = torch.dot(yy.reshape(-1), Gflat)
        return loss, torch.autograd.grad(loss, leaves)

    graphs = [None, None]
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for k in range(2):
                for _ in range(2):
                    gather_fwd_loss_bwd(k)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for k in range(2):
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph, stream=side):
                outs = gather_fwd_loss_bwd(k)
            graphs[k] = (gph, outs)
        torch.cuda.synchronize()
    except Exception as e:
        graphs = [None, None]
        sys.stderr.write("bench.py: e2e_store CUDA-graph capture failed, running eagerly: %r\n" % (e,))

    loss_host = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ready = [torch.cuda.Event() for _ in range(2)]
    losses = []

    def step(i):
        torch.cuda.current_stream().wait_event(ready[i & 1])
        prefetch(i + 1)
        if graphs[i & 1] is not None:
            gph, (loss, grads) = graphs[i & 1]
            gph.replay()
        else:
            loss, grads = gather_fwd_loss_bwd(i & 1)
        consumed[i & 1].record()
        for t, g_ in zip(leaves, grads):
            t.grad = g_
        allreduce_module_grads(leaves)
        loss_host[i & 1].copy_(loss.detach().reshape(1), non_blocking=True)
        loss_ready[i & 1].record()
        if i > 0:
            loss_ready[(i - 1) & 1].synchronize()
            losses.append(float(loss_host[(i - 1) & 1][0]))

    for e in consumed:
        e.record()
    steps = max(5, min(args.steps, 50))
    prefetch(0)
    for i in range(3):
        step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(3, 3 + steps):
        step(i)
    loss_ready[(2 + steps) & 1].synchronize()
    losses.append(float(loss_host[(2 + steps) & 1][0]))
    t1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank != 0:
        return None
    return {"value": world * B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "h2d_bytes_per_step": B * 8, "d2h_bytes_per_step": 4,
            "table": "%d x %d bf16 embeddings resident in HBM" % (n_table, D),
            "api": "EmbeddingStore.gather(sample indices from pinned host memory) + Projector.lora_forward(mode='full') + autograd backward"
                   + (", one CUDA graph per input buffer" if graphs[0] is not None else ", eager") + "; loss read back to the host every step"}


if __name__ == "__main__":
    main()
