"""GPU tests at BASELINE.json's full sizes, where the oracle is too slow to run: size-independent properties of the
factorization instead (reference tests/tests.cpp:799-856 `ApproxTest.Approx`: one-solve residual <= 200 tol;
:917-994 `ApproxTest.Repro`: bitwise repeatability; PCG converges to 1e-12 in a handful of iterations; the
preconditioner is linear; fewer dofs survive every level than entered it)."""
import numpy as np
import pytest

import spand_public_b200 as S

pytestmark = pytest.mark.gpu


def _tree(n, d, L, tol):
    A = S.neglapl(n, d)
    t = S.Tree(L)
    t.set_tol(tol)
    t.set_use_geo(True)
    t.set_Xcoo(S.linspace_nd(n, d))
    t.partition(S.symmetric_graph(A))
    t.assemble(A)
    t.factorize()
    return A, t


@pytest.mark.parametrize("n,d,L,tol,max_cg", [(1024, 2, 14, 1e-3, 8), (128, 3, 16, 1e-2, 12)])
def test_full_size_properties(n, d, L, tol, max_cg):
    """configs C3 and C4 of BASELINE.json"""
    A, t = _tree(n, d, L, tol)
    N = A.shape[0]
    b = S.random(N, 2019)
    x = t.solve(b)
    assert np.linalg.norm(A @ x - b) / np.linalg.norm(b) <= 200 * tol          # ApproxTest.Approx
    ranks = t.stats()[2].copy()
    nnz = t.nnz()
    t.assemble(A)
    t.factorize()
    assert np.array_equal(t.stats()[2], ranks) and t.nnz() == nnz              # ApproxTest.Repro
    assert np.array_equal(t.solve(b), x)
    c = S.random(N, 7)
    y = t.solve(c)
    z = t.solve(2.0 * b - 3.0 * c)                                            # the preconditioner is linear
    assert np.linalg.norm(z - (2.0 * x - 3.0 * y)) <= 1e-10 * np.linalg.norm(z)
    lg = t.log()
    left = lg["dofs_left_spars"][: L - 1]
    assert np.all(np.diff(left) <= 0) and left[-1] < 0.01 * N                  # every level shrinks the problem
    it, xc = t.cg(A, b, 500, 1e-12)
    assert it <= max_cg
    assert np.linalg.norm(A @ xc - b) / np.linalg.norm(b) <= 1e-11
