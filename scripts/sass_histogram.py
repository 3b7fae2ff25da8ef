"""SASS opcode histogram per kernel of libgpar_b200.so (cuobjdump -sass): evidence for which hardware paths the
kernels use (DMMA = fp64 tensor path, UBLKCP = bulk-copy/TMA engine, LDGSTS = cp.async, SYNCS = mbarrier ...).
    python scripts/sass_histogram.py > profiles/r2_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpar_b200", "libgpar_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
KEY = ["DMMA", "DFMA", "DADD", "DMUL", "MUFU", "UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "LDS", "STS", "LDG", "STG",
       "BAR", "ATOMG", "RED", "SHFL", "MEMBAR", "ERRBAR", "NANOSLEEP", "HMMA", "UTCHMMA", "LDTM"]
kernels = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", line)
    if m and cur:
        kernels[cur][m.group(1)] += 1
        kernels[cur]["__total"] += 1
        if m.group(1) == "DMMA" and m.group(2):
            kernels[cur]["DMMA" + m.group(2).split(".")[1] if "." in m.group(2)[1:] else "DMMA" + m.group(2)] += 0


def demangle(name):
    try:
        return subprocess.run(["cu++filt", name], stdout=subprocess.PIPE, text=True).stdout.strip().split("(")[0]
    except Exception:
        return name


print(f"# SASS opcode histogram of {os.path.relpath(lib, ROOT)} (cuobjdump -sass, sm_100a)")
print("# kernel | total instr | " + " ".join(KEY))
for k, c in kernels.items():
    row = " ".join(f"{op}={c[op]}" for op in KEY if c[op])
    print(f"{demangle(k)} | {c['__total']} | {row}")
tot = collections.Counter()
for c in kernels.values():
    tot.update(c)
print("TOTAL | " + str(tot["__total"]) + " | " + " ".join(f"{op}={tot[op]}" for op in KEY if tot[op]))
