"""SASS instruction counts per kernel of the built library (development aid; no GPU needed).
    python tools/sass_counts.py > profiles/sass_counts_rNN.txt
DMMA = FP64 tensor op, UBLKCP = 1-D bulk copy through the TMA unit, SYNCS = mbarrier ops, LDTM/STTM = tensor memory."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "grates_b200", "lib", "libgrates_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
pats = ["DMMA", "DFMA", "DADD", "DMUL", "UBLKCP", "UBLKRED", "UBLKPF", "SYNCS", "LDTM", "STTM", "UTC", "UTMALDG", "LDS", "STS",
        "STG", "LDG", "ACQBULK", "USETMAXREG", "ATOM", "RED", "BAR", "STL", "LDL"]
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        for p in pats:
            if op.startswith(p):
                counts[cur][p] += 1
                break
names = subprocess.run(["cu++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("# SASS instruction counts per kernel of %s (cuobjdump -sass, sm_100a)" % os.path.relpath(lib, ROOT))
print("# STL/LDL = local-memory spills; FP64 atomics would show as ATOM/RED with .F64 (none: see the grep at the end)")
for (k, c), name in zip(counts.items(), names):
    depth, cut = 0, len(name)
    for i, ch in enumerate(name):       # drop the parameter list, keep the template arguments
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            cut = i
            break
    name = name[:cut].replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("(bool)", "").replace("(int)", "")
    print("%-70s %s" % (name[:70], " ".join("%s=%d" % (p, c[p]) for p in pats if c[p])))
f64_atomics = len(re.findall(r"(?:ATOM|RED)[A-Z.]*\.F64", sass))
print("# FP64 atomic instructions in the whole library: %d" % f64_atomics)
