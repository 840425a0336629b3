import sys, json
sys.path.insert(0,'/root/repo')
import numpy as np, bench
import spand_public_b200 as S
cfg = bench.parse_config(sys.argv[1] if len(sys.argv)>1 else "c4")
n,d,L,tol,desc = cfg
A = bench.matrix_of(S,cfg)
t = S.Tree(L); t.set_tol(tol); t.set_use_geo(True); t.set_Xcoo(S.linspace_nd(n,d)); t.partition(S.symmetric_graph(A))
for rep in range(2):
    t.assemble(A); t.factorize()
lg = t.log()
for k in ("t_plan_elim","t_plan_scale","t_plan_spars","t_plan_merge","t_host","t_elim","t_scale","t_spars","t_merge"):
    if k in lg: print(k, [round(float(x)*1e3,1) for x in lg[k]])
print("factorize ms", t.factorize_seconds()*1e3)
