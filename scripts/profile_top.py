"""Runs on the GPU box: (1) ncu launch list of one factorization, (2) for the slowest launch of each requested
kernel-name pattern, one `ncu --set full --import-source on` capture exported to CSV (raw + source pages).
usage: python scripts/profile_top.py <tag> <config> <pattern> [<pattern> ...]   (outputs under gpurun_out/)"""
import csv, os, re, subprocess, sys
tag, cfg, pats = sys.argv[1], sys.argv[2], sys.argv[3:]
cmd = [sys.executable, "bench.py", "--config", cfg, "--steps", "1", "--warmup", "0", "--no-cpu-baseline", "--no-cg"]
out = "gpurun_out"
ll = f"{out}/{tag}_launches.csv"
subprocess.run(["ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "-c", "6000", "--csv",
                "--log-file", ll] + cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
rows = list(csv.DictReader([l for l in open(ll) if l.startswith('"')]))
def us(r):
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    return v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u in ("s", "second") else v
for pi, pat in enumerate(pats):
    best, bi = -1, -1
    for i, r in enumerate(rows):
        if re.search(pat, r["Kernel Name"]) and us(r) > best:
            best, bi = us(r), i
    if bi < 0:
        print("no launch matches", pat); continue
    print(f"pattern {pat}: launch {bi} {rows[bi]['Kernel Name'][:90]} grid {rows[bi]['Grid Size']} {best:.1f} us", flush=True)
    rep = f"/tmp/{tag}_{pi}"
    subprocess.run(["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-s", str(bi), "-c", "1",
                    "-o", rep, "-f"] + cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for page in ("raw", "source"):
        with open(f"{out}/{tag}_{pi}_{page}.csv", "w") as f:
            subprocess.run(["ncu", "-i", rep + ".ncu-rep", "--page", page, "--csv"], stdout=f, stderr=subprocess.DEVNULL)
