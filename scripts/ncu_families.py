"""Per-kernel summary of an `ncu --set full` report: for every kernel name the slowest captured launch with its launch
shape, DRAM traffic, pipe utilisation, occupancy and the largest warp-stall reasons.
usage: python scripts/ncu_families.py report.ncu-rep "<command that produced it>" > profiles/<name>.md"""
import collections, csv, io, re, subprocess, sys

rep, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rd = csv.reader(io.StringIO(out))
hdr = next(rd)
units = next(rd)
rows = list(rd)
col = {h: i for i, h in enumerate(hdr)}


def num(r, name, default=0.0):
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"):
        return default
    try:
        return float(r[i].replace(",", ""))
    except ValueError:
        return default


def unit(name):
    i = col.get(name)
    return units[i] if i is not None else ""


def to_ms(v, u):
    return v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v * 1e3 if u in ("s", "second") else v


def to_gb(v, u):
    return v * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(u, 1e-9)


stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
by = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "")
    name = re.sub(r"(spand::)?(<?unnamed>|\(anonymous namespace\))::", "", name)
    by.setdefault(name, []).append(r)
print("# ncu `--set full` captures (round 2, final binary)\n")
print(f"`{cmd}`\n\nPer kernel: the slowest captured launch. Numbers under ncu are serialised and cold-cache.\n")
print("| kernel | launches captured | slowest launch | grid x block | regs | DRAM read + written | DRAM throughput | FP64 pipe | "
      "issue slots active | warps active | L2 hit | top stalls (warps per issue) |")
print("|---|---:|---:|---|---:|---:|---:|---:|---:|---:|---:|---|")
for name, rs in by.items():
    r = max(rs, key=lambda x: num(x, "gpu__time_duration.sum"))
    t = to_ms(num(r, "gpu__time_duration.sum"), unit("gpu__time_duration.sum"))
    gb = to_gb(num(r, "dram__bytes_read.sum"), unit("dram__bytes_read.sum")) + to_gb(num(r, "dram__bytes_write.sum"), unit("dram__bytes_write.sum"))
    fp64 = max(num(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
               num(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"))
    stalls = sorted(((num(r, c), c[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for c in stall_cols), reverse=True)[:3]
    print(f"| `{name[:60]}` | {len(rs)} | {t:.3f} ms | {int(num(r, 'launch__grid_size'))} x {int(num(r, 'launch__block_size'))} | "
          f"{int(num(r, 'launch__registers_per_thread'))} | {gb:.3f} GB | {num(r, 'dram__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} % | "
          f"{fp64:.1f} % | {num(r, 'smsp__issue_active.avg.pct'):.1f} % | {num(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} % | "
          f"{num(r, 'lts__t_sector_hit_rate.pct'):.0f} % | " + ", ".join(f"{n} {v:.1f}" for v, n in stalls) + " |")
