"""GPU parity at the sizes BASELINE.json names, against committed oracle fixtures.

tests/golden/{c3,c4,c5s,s64}_oracle.json.gz were written by tests/golden/make_fullsize_golden.py: ONE run of the CPU
oracle (the restatement of the reference path, oracle/) per configuration, on the build container's host (C4: 176 s of
factorize, 20 GB). They hold what the reference's driver prints (tests/spaND.cpp:274-368): the write_stats triples
(order id, size, rank) of every cluster, dofs left per level, factor nnz, one-solve residual and the CG / GMRES
iteration count for b = random(N, 2019), x0 = 0, solver tolerance 1e-12.

Asserted (north_star): ranks equal or +-1 at ties (rule in tests/rank_parity.py, every differing cluster listed),
dofs left per level within 3 %, nnz within 0.5 %, one-solve residual within 25 % of the oracle's and <= 200 tol
(tests/tests.cpp:799-856), iteration count = oracle +- 1."""
import gzip
import json
import os

import numpy as np
import pytest

import spand_public_b200 as S
from rank_parity import rank_parity

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _golden(name):
    with gzip.open(os.path.join(HERE, "golden", f"{name}_oracle.json.gz"), "rt") as f:
        return json.load(f)


def _matrix(name, n, d):
    return S.aniso_convdiff(n) if name == "c5s" else S.neglapl(n, d)


@pytest.mark.parametrize("name,n,d", [("s64", 64, 3), ("c3", 1024, 2), ("c4", 128, 3), ("c5s", 96, 3)])
def test_against_oracle_golden(name, n, d):
    gd = _golden(name)
    L, tol = gd["nlevels"], gd["tol"]
    A = _matrix(name, n, d)
    N = A.shape[0]
    assert N == gd["N"]
    t = S.Tree(L)
    t.set_tol(tol)
    if gd["solver"] == "gmres":
        t.set_symm_kind(S.GEN)
        t.set_scaling_kind(S.PLU)
    t.set_use_geo(True)
    t.set_Xcoo(S.linspace_nd(n, d))
    t.partition(S.symmetric_graph(A))
    t.assemble(A)
    t.factorize()
    ids, size, rank = t.stats()
    rank_parity(ids, size, rank, gd["id"], gd["size"], gd["rank"], label=name)
    lg = t.log()
    for key in ("dofs_left_elim", "dofs_left_spars"):
        # +-1 decisions at ties are replicated by the translation symmetry of the grid (C3: 1.7 % of the clusters
        # differ, all by one): the totals per level stay within 3 %
        assert np.allclose(lg[key][:L], np.asarray(gd[key], dtype=float), rtol=0.03, atol=4), key
    assert abs(t.nnz() - gd["nnz"]) <= 0.005 * gd["nnz"]
    assert t.get_stop() == gd["stop"] or abs(t.get_stop() - gd["stop"]) <= 0.01 * gd["stop"] + 4
    b = S.random(N, 2019)
    x = t.solve(b)
    res = np.linalg.norm(A @ x - b) / np.linalg.norm(b)
    ro = gd["residual_one_solve"]
    print(f"[{name}] one-solve residual {res:.6e} (oracle {ro:.6e}), nnz {t.nnz()} (oracle {gd['nnz']})")
    assert res <= 200 * tol and abs(res - ro) <= 0.25 * ro
    if gd["solver"] == "gmres":
        it, xs = t.gmres(A, b, 500, 100, 1e-12)
    else:
        it, xs = t.cg(A, b, 500, 1e-12)
    print(f"[{name}] {gd['solver']} iterations {it} (oracle {gd['iterations']})")
    assert abs(it - gd["iterations"]) <= 1
    assert np.linalg.norm(A @ xs - b) / np.linalg.norm(b) <= 1e-10
