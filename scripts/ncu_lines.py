"""Hot source lines of one kernel in an .ncu-rep: python scripts/ncu_lines.py rep kernel_index [top]"""
import csv, io, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", f":::{kid}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = [i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "Line No"][0]
h = rows[hdr]
ii, ss = h.index("Instructions Executed"), h.index("# Samples")
data, ti, ts = [], 0, 0
for r in rows[hdr + 1:]:
    if len(r) <= ss or r[2] != "-":   # only the per-source-line aggregate rows
        continue
    try:
        ins, sm = int(r[ii]), int(r[ss])
    except ValueError:
        continue
    ti += ins; ts += sm
    data.append((sm, ins, r[0], r[1]))
print(rows[1][1][:100] if len(rows) > 1 else "", "| total warp-instr", ti, "samples", ts)
for sm, ins, ln, src in sorted(data, reverse=True)[:top]:
    print(f"{100*sm/max(ts,1):5.1f}% smp {100*ins/max(ti,1):5.1f}% ins  L{ln}: {src.strip()[:115]}")
