import os, sys
sys.path.insert(0, "/root/repo")
os.chdir("/root/repo")
import bench
import spand_public_b200 as S
cfg = bench.parse_config("c4")
n, d, L, tol, desc = cfg
A = S.neglapl(n, d); X = S.linspace_nd(n, d)
t = S.Tree(L); t.set_tol(tol); t.set_use_geo(True); t.set_Xcoo(X); t.partition(S.symmetric_graph(A))
t.assemble(A); t.factorize()
os.environ["SPAND_QR_TRACE"] = "1"
t.assemble(A); t.factorize()
