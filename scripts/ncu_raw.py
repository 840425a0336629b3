"""Print selected raw metrics of every kernel in an .ncu-rep (run where ncu is installed; no GPU needed)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
pat = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__", "sm__throughput.avg.pct",
                       "sm__warps_active.avg.pct", "smsp__inst_executed.sum", "sm__inst_executed_pipe", "smsp__average_warps_issue_stalled",
                       "l1tex__data_pipe_lsu_wavefronts_mem_shared", "sm__cycles_active.avg", "smsp__issue_active.avg.pct", "lts__t_sector_hit_rate",
                       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "smsp__pcsamp_warps_issue_stalled"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rd = csv.reader(io.StringIO(out))
hdr = next(rd); units = next(rd); rows = list(rd)
for i, h in enumerate(hdr):
    if h in ("Kernel Name", "Grid Size", "Block Size") or any(p in h for p in pat):
        vals = [r[i][:28] for r in rows]
        if all(v in ("0", "0.000000", "") for v in vals): continue
        print(f"{h} [{units[i]}]: " + " | ".join(vals))
