"""Where assemble() spends its time (host side): SPAND_TIMING=1 python scripts/assemble_timing.py <config>"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import spand_public_b200 as S
cfg = bench.parse_config(sys.argv[1] if len(sys.argv) > 1 else "s64")
n, d, L, tol, desc = cfg
A = S.neglapl(n, d); X = S.linspace_nd(n, d)
t = S.Tree(L); t.set_tol(tol); t.set_use_geo(True); t.set_Xcoo(X)
t0 = time.perf_counter(); t.partition(S.symmetric_graph(A)); print("partition %.3f s" % (time.perf_counter() - t0), flush=True)
for rep in range(3):
    t0 = time.perf_counter(); t.assemble(A); t1 = time.perf_counter(); t.factorize(); t2 = time.perf_counter()
    print("rep %d: assemble %.1f ms  factorize (wall) %.1f ms (device %.1f ms)" % (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, t.factorize_seconds() * 1e3), flush=True)
b = S.random(A.shape[0], 2019)
t0 = time.perf_counter(); x = t.solve(b); print("solve %.1f ms" % ((time.perf_counter() - t0) * 1e3))
t0 = time.perf_counter(); it, x = t.cg(A, b, 500, 1e-12); print("cg %d iterations, wall %.1f ms, inner %.1f ms" % (it, (time.perf_counter() - t0) * 1e3, t.t_cg * 1e3))
