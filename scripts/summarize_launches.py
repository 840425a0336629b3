"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into (a) a compact per-launch CSV and
(b) a per-kernel markdown summary, both under profiles/.
usage: python scripts/summarize_launches.py gpurun_out/launches.csv profiles/<name> "<command that was profiled>" """
import collections, csv, re, sys

src, dst, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
lines = [l for l in open(src) if l.startswith('"')]
tot, cnt, rows = collections.defaultdict(float), collections.Counter(), []
for r in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("spand::<unnamed>::", "")
    if len(name) > 60:
        name = name[:57] + "..."
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    us = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u in ("s", "second") else v
    tot[name] += us
    cnt[name] += 1
    rows.append((r["ID"], name, r["Grid Size"], r["Block Size"], us))
with open(dst + ".csv", "w") as f:
    f.write("id,kernel,grid,block,us\n")
    for r in rows:
        f.write('%s,"%s","%s","%s",%.3f\n' % r)
s = sum(tot.values())
with open(dst + ".md", "w") as f:
    f.write(f"# ncu launch list summary\n\ncommand: `{cmd}`\n\n"
            "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n"
            f"total kernel time {s/1e3:.1f} ms over {len(rows)} launches\n\n| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        f.write(f"| `{k}` | {cnt[k]} | {v/1e3:.2f} | {100*v/s:.1f}% |\n")
print(open(dst + ".md").read())
