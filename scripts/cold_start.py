"""First call on a new matrix, itemised: partition, first assemble (symbolic analysis + upload), first factorize (device
arena growth), first solve; then the steady state. usage: SPAND_TIMING=1 python scripts/cold_start.py [config]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
import spand_public_b200 as S
cfg = bench.parse_config(sys.argv[1] if len(sys.argv) > 1 else "c4")
n, d, L, tol, desc = cfg
torch.zeros(1, device="cuda").sum().item()  # CUDA context of the process
A = bench.matrix_of(S, cfg)
b = S.random(A.shape[0], 2019)
t = S.Tree(L)
t.set_tol(tol); t.set_use_geo(True); t.set_Xcoo(S.linspace_nd(n, d))
lap = []
def tick(name, f):
    t0 = time.perf_counter(); r = f(); lap.append((name, time.perf_counter() - t0)); return r
tick("partition", lambda: t.partition(S.symmetric_graph(A)))
tick("assemble #1 (symbolic analysis %s)" % "", lambda: t.assemble(A))
tick("factorize #1", lambda: t.factorize())
tick("solve #1", lambda: t.solve(b))
tick("assemble #2", lambda: t.assemble(A))
tick("factorize #2", lambda: t.factorize())
tick("solve #2", lambda: t.solve(b))
print("analysis inside assemble #1: %.3f s" % t.analyze_seconds())
for k, v in lap:
    print("%-40s %.3f s" % (k, v))
