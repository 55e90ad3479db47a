"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.
  python profiles/summarize_ncu.py launches gpurun_out/launches.csv  [skip_substr ...]   -> per-kernel time shares
  python profiles/summarize_ncu.py full gpurun_out/prof.ncu-rep                         -> key metrics per profiled launch
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__inst_executed.sum"]


def launches(path, only=("dmi::",)):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
    mine = {k: v for k, v in agg.items() if any(o in k for o in only)}
    tot = sum(sum(v) for v in mine.values())
    print(f"# {path}: {sum(len(v) for v in mine.values())} launches of library kernels, {tot/1e3:.1f} us total (cold-cache, serialised)")
    for n, v in sorted(mine.items(), key=lambda kv: -sum(kv[1])):
        print(f"{100*sum(v)/tot:5.1f}%  {len(v):3d}x  avg {sum(v)/len(v)/1e3:8.2f} us  {n[:110]}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [(k, hdr.index(k)) for k in KEYS if k in hdr]
    extra = [(h, i) for i, h in enumerate(hdr) if "pipe_tensor" in h and h not in KEYS][:6]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("==", r[ki][:120])
        for k, i in cols + extra:
            print(f"   {k:80s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
