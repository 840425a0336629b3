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
    out[cname] = {"ctas": v[10], "cta_cycles_total": tot, "share_pct": {names[i]: round(100.0 * v[i] / tot, 1) for i in range(9)}}
print(json.dumps(out, indent=1))
