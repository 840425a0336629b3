"""One C4 factorization per process: residual, rank checksum, nnz and device properties, to compare across processes
and boxes (the factorization is deterministic by construction: fixed task order, no atomics in the numeric path)."""
import hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench, spand_public_b200 as S
cfg = bench.parse_config(sys.argv[1] if len(sys.argv) > 1 else "c4")
n, d, L, tol, desc = cfg
A = S.neglapl(n, d); X = S.linspace_nd(n, d)
t = S.Tree(L); t.set_tol(tol); t.set_use_geo(True); t.set_Xcoo(X); t.partition(S.symmetric_graph(A))
out = []
for rep in range(2):
    t.assemble(A); t.factorize()
    ranks = t.stats()[2]
    b = S.random(A.shape[0], 2019); x = t.solve(b)
    lg = t.log()
    out.append({"residual": float(np.linalg.norm(A @ x - b) / np.linalg.norm(b)), "rank_sum": int(ranks.sum()),
                "rank_md5": hashlib.md5(ranks.tobytes()).hexdigest(), "nnz": int(t.nnz()),
                "x_md5": hashlib.md5(x.tobytes()).hexdigest(),
                "dofs_left": [int(v) for v in lg["dofs_left_spars"]], "ms": round(t.factorize_seconds() * 1e3, 1)})
p = torch.cuda.get_device_properties(0)
print(json.dumps({"device": p.name, "sms": p.multi_processor_count, "l2": p.L2_cache_size, "runs": out}))
