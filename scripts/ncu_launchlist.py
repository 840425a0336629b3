"""One factorization of a bench configuration for an ncu launch list: warm-up factorization first (skipped with
`ncu -s`), then the measured one. usage (under ncu): python scripts/ncu_launchlist.py c4"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import spand_public_b200 as S
cfg = bench.parse_config(sys.argv[1] if len(sys.argv) > 1 else "c4")
n, d, L, tol, desc = cfg
gen = bench.is_aniso(cfg)
A = bench.matrix_of(S, cfg)
t = S.Tree(L)
if gen:
    t.set_symm_kind(S.GEN); t.set_scaling_kind(S.PLU)
t.set_tol(tol); t.set_use_geo(True); t.set_Xcoo(S.linspace_nd(n, d)); t.partition(S.symmetric_graph(A))
t.assemble(A); t.factorize()
print("launches of one factorization:", t.kernel_launches(), flush=True)
t.assemble(A); t.factorize()
print("device ms", t.factorize_seconds() * 1e3)
