"""Per-kernel-family device time of one factorization (CUDA events around every launch) + optional RRQR task-shape dump.
usage: python scripts/family_times.py <config> [dumpfile]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import spand_public_b200 as S

cfg = bench.parse_config(sys.argv[1] if len(sys.argv) > 1 else "s64")
if len(sys.argv) > 2:
    os.environ["SPAND_DUMP_QR"] = sys.argv[2]
    if os.path.exists(sys.argv[2]):
        os.remove(sys.argv[2])
n, d, L, tol, desc = cfg
A = S.neglapl(n, d); X = S.linspace_nd(n, d)
t = S.Tree(L); t.set_tol(tol); t.set_use_geo(True); t.set_Xcoo(X); t.partition(S.symmetric_graph(A))
t.assemble(A); t.factorize()          # warm-up (also writes the dump)
os.environ.pop("SPAND_DUMP_QR", None)
t.set_profile(True)
t.assemble(A); t.factorize()
lg = t.log()
out = {"config": desc, "factorize_s": t.factorize_seconds(), "families": t.family_stats(),
       "phases": {k: float(lg[k].sum()) for k in ("t_elim", "t_scale", "t_spars", "t_merge", "t_host")},
       "plan": {k: [round(float(v), 4) for v in lg[k]] for k in ("t_plan_elim", "t_plan_scale", "t_plan_spars", "t_plan_merge")},
       "dev": {k: [round(float(v), 4) for v in lg[k]] for k in ("t_elim", "t_scale", "t_spars", "t_merge")},
       "flops": {k: float(lg[k].sum()) for k in ("fl_pivot", "fl_panel", "fl_schur", "fl_rrqr_rank", "fl_rrqr_full")},
       "bytes": {k: float(lg[k].sum()) for k in ("by_scale", "by_rrqr", "by_merge")}}
print(json.dumps(out))
