"""Per-kernel SASS opcode summary of the shipped library (what proves the Blackwell-native paths; B200_PROFILING.md table):
   python profiles/sass_opcodes.py > profiles/r2_sass_opcodes.txt
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG / UTMASTG = TMA load / store, UTCBAR = tcgen05.commit,
HMMA = legacy mma.sync, LDGMC = multimem.ld_reduce (NVLS in-switch reduction), LDGSTS = cp.async,
PREEXIT / ACQBULK = griddepcontrol.launch_dependents / .wait (programmatic dependent launch)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "sample-efficient-multimodality_b200", "csrc", "libdmi_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True).stdout
pats = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMACMDFLUSH", "HMMA", "LDGMC", "LDGSTS", "SYNCS", "ELECT", "R2UR",
        "PREEXIT", "ACQBULK"]
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    for p in pats:
        if op == p or op.startswith(p + "."):
            counts[cur][p] += 1
    if op.startswith("UTCHMMA.2CTA"):
        counts[cur]["UTCHMMA.2CTA"] += 0   # already counted by the prefix rule


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], stdout=subprocess.PIPE, text=True).stdout.strip()
    except Exception:
        return n


print("# SASS opcode counts per kernel of libdmi_b200.so (cuobjdump -sass), kernels with tensor-core / TMA / multimem instructions first")
print("# columns:", " ".join(pats))
rows = []
for k, c in counts.items():
    name = demangle(k).replace("dmi::", "").replace("(anonymous namespace)::", "")
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    rows.append((-(c["UTCHMMA"] + c["UTMALDG"] + c["UTMASTG"] + c["LDGMC"] + c["HMMA"]), name, c))
for _, name, c in sorted(rows, key=lambda r: (r[0], r[1])):
    print(f"{name[:70]:70s} " + " ".join(f"{p}={c[p]}" for p in pats if c[p]))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("# total:", " ".join(f"{p}={tot[p]}" for p in pats))
