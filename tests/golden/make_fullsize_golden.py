"""Runs the CPU oracle ONCE on the full-size configurations BASELINE.json names and writes
tests/golden/{c3,c4,c5s,s64}_oracle.json.gz (committed fixtures; the oracle is far too slow to run inside the GPU
tests at these sizes).

    python tests/golden/make_fullsize_golden.py c3 c4 c5s s64

Each fixture holds what the reference's driver prints / write_stats exports for the run (tests/spaND.cpp:274-368,
include/tree.h:242-251): the per-cluster (order id, original size, final rank) triples, dofs left after eliminate /
sparsify per level, the factor nnz, the one-solve residual, the CG / GMRES iteration count (b = random(N, 2019),
x0 = 0, 500 iterations max, solver tolerance 1e-12) and the oracle's own factorize wall time on this host (tfact),
with the host's core count. Protocol: reference tests/spaND.cpp:310-368; thresholds: tests/tests.cpp:799-856.

Needs oracle/_build/liboracle.so (make -C oracle) and ~20 GB of host memory for c4. Does not touch the GPU.
"""
import gzip
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle_lib as O  # noqa: E402
import spand_public_b200 as S  # noqa: E402
from bench import CONFIGS, is_aniso, matrix_of  # noqa: E402


def run(name):
    cfg = CONFIGS[name]
    n, d, L, tol, desc = cfg
    gen = is_aniso(cfg)
    O.lib().orc_set_threads(1)
    A = matrix_of(S, cfg)
    N = A.shape[0]
    X = S.linspace_nd(n, d)
    G = S.symmetric_graph(A)
    t = O.OracleTree(L, tol=tol, symm_kind=O.GEN, scaling_kind=O.PLU) if gen else O.OracleTree(L, tol=tol)
    t.set_coords(X)
    t0 = time.perf_counter()
    t.partition(G)
    t1 = time.perf_counter()
    t.assemble(A)
    t2 = time.perf_counter()
    t.factorize()
    t3 = time.perf_counter()
    b = S.random(N, 2019)
    x = t.solve(b)
    t4 = time.perf_counter()
    res = float(np.linalg.norm(A @ x - b) / np.linalg.norm(b))
    if gen:
        it, xs = t.gmres(A, b, 500, 100, 1e-12)
    else:
        it, xs = t.cg(A, b, 500, 1e-12)
    t5 = time.perf_counter()
    res_it = float(np.linalg.norm(A @ xs - b) / np.linalg.norm(b))
    lg = t.log()
    ids, size, rank = t.stats()
    out = {
        "config": name, "workload": desc, "N": int(N), "nlevels": L, "tol": tol, "solver": "gmres" if gen else "cg",
        "iterations": int(it), "residual_one_solve": res, "residual_after_solver": res_it, "nnz": int(t.nnz()),
        "stop": int(t.get_stop()),
        "dofs_left_elim": lg["dofs_left_elim"].astype(np.int64).tolist(),
        "dofs_left_spars": lg["dofs_left_spars"].astype(np.int64).tolist(),
        "rank_before": lg["rank_before"].astype(np.int64).tolist(),
        "rank_after": lg["rank_after"].astype(np.int64).tolist(),
        "gflop": float(sum(lg[k].sum() for k in ("fl_pivot", "fl_panel", "fl_schur", "fl_rrqr_rank")) / 1e9),
        "id": ids.tolist(), "size": size.tolist(), "rank": rank.tolist(),
        "host": {"cpus": os.cpu_count(), "blas_threads": 1, "tpart_s": t1 - t0, "tassm_s": t2 - t1, "tfact_s": t3 - t2,
                 "tsolve_s": t4 - t3, "titer_s": t5 - t4},
    }
    path = os.path.join(HERE, f"{name}_oracle.json.gz")
    with gzip.open(path, "wt") as f:
        json.dump(out, f)
    print(f"{name}: N={N} tfact={t3 - t2:.1f}s nnz={out['nnz']} it={it} res={res:.3e} -> {path}", flush=True)


if __name__ == "__main__":
    for nm in sys.argv[1:] or ["c3", "c4", "c5s", "s64"]:
        run(nm)
