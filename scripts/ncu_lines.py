"""Per-source-line summary of an ncu capture taken with --import-source on:
   ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > src.csv ; python scripts/ncu_lines.py src.csv [top]
Aggregates the SASS rows of every CUDA source line: stall samples, executed instructions, dominant stall reasons."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
agg = collections.OrderedDict()
src_of = {}
cur_file = ""
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 6 and r[0] == "Line No":
        hdr = r
        sa, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
        st0 = hdr.index("stall_barrier")
        st1 = hdr.index("stall_wait") + 1
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] == "":
        continue  # SASS rows: already summed in the row of their CUDA line
    try:
        line = int(r[0])
        key = (cur_file, line)
        a = agg.setdefault(key, [0, 0, collections.Counter()])
        src_of[key] = r[1]
        a[0] += int(r[sa]); a[1] += int(r[ie])
        for i in range(st0, st1):
            v = int(r[i] or 0)
            if v: a[2][hdr[i]] += v
    except ValueError:
        pass
tot = sum(a[0] for a in agg.values()) or 1
toti = sum(a[1] for a in agg.values()) or 1
print(f"total samples {tot}, warp instructions {toti}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    reasons = ", ".join(f"{k.replace('stall_','')} {100*v/max(1,a[0]):.0f}%" for k, v in a[2].most_common(3))
    print(f"{100*a[0]/tot:5.1f}% samples {100*a[1]/toti:5.1f}% instr  {key[0]}:{key[1]:<5d} [{reasons}]  {src_of.get(key,'').strip()[:90]}")
