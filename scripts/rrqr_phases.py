"""Per-phase clock64 breakdown of the RRQR kernel (needs a -DSPAND_RRQR_TIMING build:
make -C spand_public_b200/csrc OUT=../_build_timing EXTRA=-DSPAND_RRQR_TIMING; SPAND_B200_BUILD=_build_timing).
usage: python scripts/rrqr_phases.py <config>      prints, per kernel shape class, the share of every phase"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import spand_public_b200 as S

cfg = bench.parse_config(sys.argv[1] if len(sys.argv) > 1 else "s64")
n, d, L, tol, desc = cfg
A = S.neglapl(n, d); X = S.linspace_nd(n, d)
t = S.Tree(L); t.set_tol(tol); t.set_use_geo(True); t.set_Xcoo(X); t.partition(S.symmetric_graph(A))
t.assemble(A); t.factorize()
buf = (C.c_ulonglong * 48)()
lib = S.lib()
lib.spand_debug_rrqr_phases.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
lib.spand_debug_rrqr_phases.restype = None
lib.spand_debug_rrqr_phases(buf, 1)
_h0 = (C.c_ulonglong * 16)()
lib.spand_debug_hc2_stats.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
lib.spand_debug_hc2_stats.restype = None
lib.spand_debug_hc2_stats(_h0, 1)
t.assemble(A); t.factorize()
lib.spand_debug_rrqr_phases(buf, 1)
names = ["setup(gather+norms)", "select+pull", "aux=V^T v", "sweep(warp0)", "wait slowest warp", "trailing update",
         "propose", "cluster barrier", "scatter", "-"]
out = {"config": desc, "factorize_ms": t.factorize_seconds() * 1e3}
for ci, cname in enumerate(["smem panel", "streaming 256 thr", "global 512 thr"]):
    v = [buf[ci * 16 + i] for i in range(16)]
    tot = sum(v[:10])
    if tot == 0:
        continue
    out[cname] = {"ctas": v[10], "cta_cycles_total": tot, "share_pct": {names[i]: round(100.0 * v[i] / tot, 1) for i in range(9)},
                  "hot_cold": {"cta_steps": v[11], "blocks_full": v[12], "blocks_early": v[13],
                               "steps_per_block": v[11] / max(1, v[12] + v[13]),
                               "hot_fraction_at_block_start": v[14] / max(1, v[15]),
                               "cycles_per_cta_step": tot / max(1, v[11])}}
h = (C.c_ulonglong * 16)()
lib.spand_debug_hc2_stats.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
lib.spand_debug_hc2_stats.restype = None
lib.spand_debug_hc2_stats(h, 1)
h = list(h)
if h[1]:
    cyc = sum(h[6:10])
    out["hot-set kernel"] = {"task_steps": h[0], "blocks": h[1], "blocks_closed_early": h[2],
                             "steps_per_block": h[0] / h[1], "hot_columns_per_block": h[3] / h[1],
                             "hot_fraction": h[3] / max(1, h[4]), "threshold_retries": h[5],
                             "cycles_per_step_first_cta": cyc / max(1, h[0]),
                             "share_pct": {"block boundary (exchange, classify, copy)": round(100.0 * h[6] / cyc, 1),
                                           "hot loop": round(100.0 * h[7] / cyc, 1),
                                           "block end (T, write back, cold refresh)": round(100.0 * h[8] / cyc, 1),
                                           "gather + norms + scatter": round(100.0 * h[9] / cyc, 1)}}
print(json.dumps(out, indent=1))
