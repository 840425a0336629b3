"""SASS evidence for profiles/: per kernel of the built library, the number of FP64 tensor-core (DMMA), FP64 FMA (DFMA),
TMA bulk-copy (UBLKCP) / mbarrier (SYNCS), async-copy (LDGSTS), cluster (UCGABAR / distributed shared memory) and
warp-shuffle instructions, from `cuobjdump -sass`. Usage: python scripts/sass_summary.py [lib.so] > profiles/rN_sass.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "spand_public_b200", "_build", "libspand_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
MN = ["DMMA", "DFMA", "UBLKCP", "SYNCS", "LDGSTS", "UCGABAR", "SHFL", "BAR", "MUFU.RCP64H", "LDG", "STG", "LDS", "STS"]
cnt = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"spand::\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*", "", name)
        name = re.sub(r"^void ", "", name)
        cur = cnt.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        cur["_all"] += 1
        for k in MN:
            if op == k or op.startswith(k + ".") or op.startswith(k + "_"):
                cur[k] += 1
print("# SASS summary of `%s` (sm_100a)\n" % os.path.relpath(lib, ROOT))
print("`cuobjdump -sass` of the shipped library, instruction counts per kernel (static code, not executed counts). "
      "FP64 has no tcgen05 kind on Blackwell: the FP64 tensor-core path is `DMMA` (mma.sync.m8n8k4.f64); `UBLKCP` + `SYNCS` "
      "are the TMA bulk copies and mbarrier operations of the staged cold refresh (`SPAND_HC2_TMA=1`); `UCGABAR` is the "
      "cluster barrier of the distributed-shared-memory kernels.\n")
print("| kernel | instr | " + " | ".join(MN) + " |")
print("|---|---:|" + "---:|" * len(MN))
tot = collections.Counter()
for name, c in cnt.items():
    if c["_all"] < 40:
        continue
    print("| `%s` | %d | " % (name[:90], c["_all"]) + " | ".join(str(c[k]) if c[k] else "" for k in MN) + " |")
    tot.update(c)
print("| **total** | %d | " % tot["_all"] + " | ".join(str(tot[k]) for k in MN) + " |")
