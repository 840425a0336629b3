"""Per-kernel time and DRAM traffic of one factorization from an ncu launch list taken with
   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv ...
Writes profiles/<name>.md (per kernel: launches, time share, DRAM bytes) and profiles/<name>.json (per family: DRAM
bytes per launch, the `roofline.traffic` figure of bench.py). The second half of the list is used (the first
factorization of scripts/ncu_launchlist.py is the warm-up).
usage: python scripts/ncu_traffic.py gpurun_out/launches.csv profiles/r2_ncu_traffic "<command>" """
import collections, csv, json, re, sys

src, dst, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
lines = [l for l in open(src) if l.startswith('"')]
per = collections.OrderedDict()  # launch id -> {name, time_us, rd, wr}
for r in csv.DictReader(lines):
    i = int(r["ID"])
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("spand::<unnamed>::", "").replace("spand::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    e = per.setdefault(i, {"name": name, "us": 0.0, "rd": 0.0, "wr": 0.0})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    m = r["Metric Name"]
    if m.startswith("gpu__time"):
        e["us"] = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v * 1e6 if u in ("s", "second") else v
    else:
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        e["rd" if "read" in m else "wr"] = v * mult
ids = sorted(per)
half = ids[len(ids) // 2:]  # the measured factorization
fam_of = lambda n: ("rrqr" if n.startswith("rrqr") else "trsm" if ("trsm" in n or "scale" in n or "trtri" in n or "rowperm" in n)
                    else "gemm" if "gemm" in n else "potrf" if ("potrf" in n or "getrf" in n) else
                    "copy" if ("copy" in n or "memset" in n.lower()) else "other")
k = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
f = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for i in half:
    e = per[i]
    short = re.sub(r"<.*", "", e["name"])
    for d, key in ((k, e["name"][:70]), (f, fam_of(short))):
        d[key][0] += 1; d[key][1] += e["us"]; d[key][2] += e["rd"]; d[key][3] += e["wr"]
tot = sum(v[1] for v in k.values())
with open(dst + ".md", "w") as out:
    out.write(f"# ncu launch list: time and DRAM traffic per kernel\n\ncommand: `{cmd}`\n\nPer-launch times under ncu are "
              "cold-cache and serialised: compare SHARES, not absolutes. DRAM bytes = dram__bytes_read.sum + "
              f"dram__bytes_write.sum.\n\n{len(half)} launches, {tot/1e3:.1f} ms of kernel time\n\n"
              "| family | launches | ms | share | DRAM GB read | DRAM GB written |\n|---|---:|---:|---:|---:|---:|\n")
    for key, v in sorted(f.items(), key=lambda x: -x[1][1]):
        out.write(f"| {key} | {v[0]} | {v[1]/1e3:.2f} | {100*v[1]/tot:.1f}% | {v[2]/1e9:.2f} | {v[3]/1e9:.2f} |\n")
    out.write("\n| kernel | launches | ms | share | DRAM GB read | DRAM GB written |\n|---|---:|---:|---:|---:|---:|\n")
    for key, v in sorted(k.items(), key=lambda x: -x[1][1])[:40]:
        out.write(f"| `{key}` | {v[0]} | {v[1]/1e3:.2f} | {100*v[1]/tot:.1f}% | {v[2]/1e9:.2f} | {v[3]/1e9:.2f} |\n")
json.dump({"source": f"profiles/{dst.split('/')[-1]}.md ({cmd})",
           "families": {key: {"launches": v[0], "kernel_ms_under_ncu": v[1] / 1e3, "dram_bytes": v[2] + v[3],
                              "dram_bytes_per_launch": (v[2] + v[3]) / max(1, v[0])} for key, v in f.items()}},
          open(dst + ".json", "w"), indent=1)
print(open(dst + ".md").read())
