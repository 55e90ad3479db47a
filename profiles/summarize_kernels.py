"""ncu CSV (one row per launch and metric) -> one line per kernel class: launches, time, DRAM bytes, achieved TB/s and fraction of the
measured HBM peak, tensor-pipe % (max over launches).   python profiles/summarize_kernels.py gpurun_out/all_kernels.csv > profiles/r2_all_kernels_ncu.txt"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6451.5
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
d = collections.OrderedDict()
for r in rows[1:]:
    e = d.setdefault(r[ii], {"k": r[ki], "grid": r[gi], "block": r[bi]})
    e[r[mi]] = float(r[vi].replace(",", ""))
agg = collections.OrderedDict()
for e in d.values():
    name = re.sub(r"\(.*", "", e["k"].replace("(anonymous namespace)::", "").replace("<unnamed>::", "")).replace("void ", "").replace("dmi::", "")
    if name.startswith("at::") or name.startswith("nvjet") or "cutlass" in name or "elementwise" in name:
        name = "[torch] " + name[:60]
    a = agg.setdefault((name, e["grid"], e["block"]), [])
    a.append(e)
print(f"# one launch list of profiles/all_kernels_probe.py under ncu (--clock-control none; serialised launches); HBM peak used: {peak:.0f} GB/s (MEASURED_PEAKS.json)")
print(f"# {'kernel':66s} {'grid':>8s} {'blk':>4s} {'n':>3s} {'us':>9s} {'rd MB':>8s} {'wr MB':>8s} {'TB/s':>6s} {'of HBM':>6s} {'tensor%':>7s}")
for (name, grid, block), es in sorted(agg.items(), key=lambda kv: -sum(e["gpu__time_duration.sum"] for e in kv[1])):
    t = sum(e["gpu__time_duration.sum"] for e in es) / len(es) / 1e3
    rd = sum(e.get("dram__bytes_read.sum", 0) for e in es) / len(es) / 1e6
    wr = sum(e.get("dram__bytes_write.sum", 0) for e in es) / len(es) / 1e6
    tp = max(e.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0) for e in es)
    tbs = (rd + wr) / t if t > 0 else 0
    g = grid.replace(" ", "")
    print(f"  {name[:66]:66s} {g[:8]:>8s} {block.split(',')[0].strip('( '):>4s} {len(es):3d} {t:9.1f} {rd:8.1f} {wr:8.1f} {tbs:6.2f} {tbs * 1e3 / peak:6.2f} {tp:7.1f}")
