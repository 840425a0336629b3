"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Tolerances (north_star): identical ordering and hierarchy; ranks equal or within +-1 at RRQR pivot ties (reported);
factor-applied residual within 1e-10 relative of the oracle's for exact configurations; PCG iterations +-1.
Entry-wise comparison of the trailing matrix is possible wherever no rank-revealing QR has run yet (its pivot order
is only defined up to ties), so those stop points are compared to 1e-13; after a sparsification the comparison uses
orthogonal invariants (ranks, Frobenius norm) and the solve.
"""
import numpy as np
import pytest

import oracle_lib as O
import spand_public_b200 as S
from rank_parity import rank_parity

pytestmark = pytest.mark.gpu


def _pair(n, d, L, tol, skip=0, geo=True, A=None):
    A = S.neglapl(n, d) if A is None else A
    X = S.linspace_nd(n, d)
    g = S.Tree(L)
    g.set_tol(tol)
    g.set_skip(skip)
    o = O.OracleTree(L, tol=tol, skip=skip)
    if geo:
        g.set_use_geo(True)
        g.set_Xcoo(X)
        o.set_coords(X)
    g.partition(A)
    o.partition(A)
    return A, g, o


def _fro(T):
    return float(np.sqrt((T.data**2).sum()))


@pytest.mark.parametrize("n,d,L", [(5, 2, 3), (32, 2, 5), (10, 3, 4), (15, 3, 5)])
def test_ordering_and_hierarchy_identical(n, d, L):
    A, g, o = _pair(n, d, L, 1e-2)
    assert np.array_equal(g.get_assembly_perm(), o.perm())
    for a, b in zip(g.partition_ids(), o.partition_ids()):
        assert np.array_equal(a, b)
    g.assemble(A)
    o.assemble(A)
    assert np.array_equal(g.stats()[0], o.stats()[0]) and np.array_equal(g.stats()[1], o.stats()[1])
    assert abs(g.get_trailing_mat() - o.trailing_mat()).max() == 0.0  # Assembly.Consistency on the device blocks


@pytest.mark.parametrize("n,d,L,tol", [(32, 2, 5, 1e-2), (10, 3, 4, 1e-2), (20, 2, 4, 0.0), (15, 3, 5, 1e-14)])
def test_stop_points_before_first_rrqr(n, d, L, tol):
    """eliminate (POTRF+TRSM+Schur GEMM with fill-in) and scale at level 0, entry by entry."""
    A, g, o = _pair(n, d, L, tol)
    for phase in (0, 1):
        g.set_stop(0, phase)
        o.set_stop(0, phase)
        g.assemble(A)
        o.partition(A)
        o.assemble(A)
        g.factorize()
        o.factorize()
        Tg, To = g.get_trailing_mat(), o.trailing_mat()
        assert Tg.nnz == To.nnz
        assert abs(Tg - To).max() <= 1e-13 * abs(To).max()
        assert g.nnz() == o.nnz()


@pytest.mark.parametrize("n,d,L,tol", [(20, 2, 4, 0.0), (15, 3, 5, 1e-14), (10, 2, 3, 0.0), (16, 2, 6, 1e-14)])
def test_exact_configs_match_everywhere(n, d, L, tol):
    """tol in {0, 1e-14}: no truncation, so every stop point of every level matches entry-wise."""
    A, g, o = _pair(n, d, L, tol)
    for lvl in range(L):
        for phase in range(4):
            g.set_stop(lvl, phase)
            o.set_stop(lvl, phase)
            g.assemble(A)
            o.partition(A)
            o.assemble(A)
            g.factorize()
            o.factorize()
            Tg, To = g.get_trailing_mat(), o.trailing_mat()
            assert Tg.nnz == To.nnz
            if To.nnz:
                assert abs(Tg - To).max() <= 1e-12 * abs(To).max(), (lvl, phase)
            assert np.array_equal(g.stats()[2], o.stats()[2])
            assert g.nnz() == o.nnz()


@pytest.mark.parametrize("n,d,L,tol,skip", [(5, 2, 3, 1e-14, 0), (10, 2, 4, 1e-14, 4), (20, 2, 5, 0.0, 1000),
                                            (5, 3, 2, 1e-14, 0), (15, 3, 5, 0.0, 1000), (15, 3, 4, 1e-14, 0),
                                            (20, 2, 1, 1e-14, 0)])
def test_exact_residual(n, d, L, tol, skip):
    # reference tests/tests.cpp:562-609 (ApproxTest.Exact, SPD/LLT/MND rows): one solve gives <= 1e-10
    A, g, o = _pair(n, d, L, tol, skip, geo=False)
    g.assemble(A)
    g.factorize()
    b = S.random(A.shape[0], L + 2019)
    x = g.solve(b)
    assert np.linalg.norm(A @ x - b) / np.linalg.norm(b) <= 1e-10


@pytest.mark.parametrize("n,d", [(10, 2), (20, 2), (5, 3), (15, 3)])
def test_approx_residual_thresholds(n, d):
    # reference tests/tests.cpp:799-856 (ApproxTest.Approx): err <= 5e-12 (tol = 0) else <= 200 tol
    A = S.neglapl(n, d)
    b = S.random(A.shape[0], 2019)
    for L in (1, 3, 5):
        for skip in (0, 1, 2):
            g = S.Tree(L)
            g.set_skip(skip)
            g.partition(A)
            for tol in (0.0, 1e-10, 1e-6, 1e-2, 10.0):
                g.set_tol(tol)
                g.assemble(A)
                g.factorize()
                x = g.solve(b)
                err = np.linalg.norm(A @ x - b) / np.linalg.norm(b)
                assert err <= (5e-12 if tol == 0.0 else tol * 2e2), (L, skip, tol, err)


def _rank_report(g, o, label=""):
    """north_star rank rule (tests/rank_parity.py): equal, or +-1 at ties, every differing cluster listed."""
    ig, sg, rg = g.stats()
    io, so, ro = o.stats()
    ndiff = rank_parity(ig, sg, rg, io, so, ro, label=label)
    return rg.astype(int) - ro.astype(int), ndiff


@pytest.mark.parametrize("n,d,L,tol,coords", [(32, 2, 5, 1e-2, "c1"), (30, 3, 8, 1e-2, "c2"), (64, 2, 8, 1e-3, "lin"),
                                              (20, 3, 6, 1e-2, "lin")])
def test_full_factorization_vs_oracle(n, d, L, tol, coords):
    """Configs C1 / C2 of BASELINE.json and two more: ranks, dofs-left per level, nnz, residual, PCG count."""
    A = S.neglapl(n, d)
    X = S.linspace_nd(n, d)
    if coords == "c1":
        X = X[::-1] + 1.0  # reference mats/32x32.mm
    g = S.Tree(L)
    g.set_tol(tol)
    g.set_use_geo(True)
    g.set_Xcoo(X)
    o = O.OracleTree(L, tol=tol)
    o.set_coords(X)
    G = S.symmetric_graph(A)
    g.partition(G)
    o.partition(G)
    g.assemble(A)
    o.assemble(A)
    g.factorize()
    o.factorize()
    diff, ndiff = _rank_report(g, o)
    lg, lo = g.log(), o.log()
    assert np.allclose(lg["dofs_left_elim"], lo["dofs_left_elim"], rtol=0.02, atol=4)
    assert np.allclose(lg["dofs_left_spars"], lo["dofs_left_spars"], rtol=0.02, atol=4)
    assert abs(g.nnz() - o.nnz()) <= 0.005 * o.nnz()
    b = S.random(A.shape[0], 2019)
    xg, xo = g.solve(b), o.solve(b)
    rg = np.linalg.norm(A @ xg - b) / np.linalg.norm(b)
    ro = np.linalg.norm(A @ xo - b) / np.linalg.norm(b)
    assert rg <= 200 * tol and abs(rg - ro) <= 0.25 * ro
    itg, xg = g.cg(A, b, 500, 1e-12)
    ito, _ = o.cg(A, b, 500, 1e-12)
    assert abs(itg - ito) <= 1
    assert np.linalg.norm(A @ xg - b) / np.linalg.norm(b) < 1e-11


@pytest.mark.parametrize("env", [{"SPAND_RRQR_FORCE_GLOBAL": "1"}, {"SPAND_RRQR_FORCE_GLOBAL": "1", "SPAND_RRQR_GTOP": "8"},
                                 {"SPAND_RRQR_FORCE_GLOBAL": "1", "SPAND_RRQR_GTOP": "2", "SPAND_RRQR_THETA": "0.9"},
                                 {"SPAND_RRQR_COL": "0"}, {"SPAND_RRQR_COL": "0", "SPAND_RRQR_HC2": "2", "SPAND_HC2_MINKB": "0"},
                                 {"SPAND_RRQR_COL": "0", "SPAND_RRQR_HC2": "2", "SPAND_HC2_MINKB": "0", "SPAND_HC2_TMA": "1"},
                                 {"SPAND_RRQR_COLGP": "1", "SPAND_RRQR_COLROWS": "128"},
                                 {"SPAND_RRQR_FORCE_G": "1"}, {"SPAND_RRQR_FORCE_G": "2"},
                                 {"SPAND_RRQR_FORCE_G": "4"}, {"SPAND_RRQR_FORCE_G": "8"}, {"SPAND_RRQR_FORCE_G": "16"},
                                 {"SPAND_RRQR_MODE": "smem"},
                                 {"SPAND_RRQR_TMIN": "1", "SPAND_RRQR_SMEM1KB": "0"},
                                 {"SPAND_RRQR_TMIN": "1", "SPAND_RRQR_SMEM1KB": "0", "SPAND_RRQR_CTAS": "100000"}])
def test_every_rrqr_kernel_shape_matches_oracle(env, monkeypatch):
    """The RRQR launch shape (warp team, CTA, 2..16-CTA cluster with the panel in distributed shared memory, 16-CTA
    cluster streaming the panel from global scratch, 256-thread streaming CTAs with 1..16-CTA clusters) is picked from
    the task size and the width of the wavefront; force each shape on one problem and compare ranks / trailing matrix
    / residual with the oracle."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    n, d, L, tol = 20, 3, 6, 1e-2
    A, g, o = _pair(n, d, L, tol)
    g.assemble(A)
    o.assemble(A)
    g.factorize()
    o.factorize()
    diff, ndiff = _rank_report(g, o)
    assert abs(g.nnz() - o.nnz()) <= 0.005 * o.nnz()
    b = S.random(A.shape[0], 2019)
    xg, xo = g.solve(b), o.solve(b)
    rg = np.linalg.norm(A @ xg - b) / np.linalg.norm(b)
    ro = np.linalg.norm(A @ xo - b) / np.linalg.norm(b)
    assert rg <= 200 * tol and abs(rg - ro) <= 0.25 * ro
    # exact configuration: the factorization must stay exact whatever the shape
    A, g, o = _pair(12, 3, 4, 1e-14)
    g.assemble(A)
    g.factorize()
    x = g.solve(b[:A.shape[0]])
    assert np.linalg.norm(A @ x - b[:A.shape[0]]) / np.linalg.norm(b[:A.shape[0]]) <= 1e-10


def test_symbolic_plan_is_reused_and_refreshed():
    """The block structure is analysed by the first assemble() of a (partition, pattern) pair and reused afterwards:
    same values -> bit-identical solve; new values on the same pattern -> correct factorization without re-analysis;
    another pattern -> re-analysis."""
    n, d, L = 12, 3, 5
    A = S.neglapl(n, d)
    X = S.linspace_nd(n, d)
    g = S.Tree(L)
    g.set_tol(0.0)
    g.set_use_geo(True)
    g.set_Xcoo(X)
    g.partition(A)
    g.assemble(A)
    t_first = g.analyze_seconds()
    g.factorize()
    b = S.random(A.shape[0], 11)
    x1 = g.solve(b)
    g.assemble(A)
    assert g.analyze_seconds() == t_first          # plan reused: no new analysis
    g.factorize()
    assert np.array_equal(x1, g.solve(b))
    B = A.copy()
    B.data = B.data * 1.5                          # same pattern, other values
    g.assemble(B)
    assert g.analyze_seconds() == t_first
    g.factorize()
    x2 = g.solve(b)
    assert np.linalg.norm(B @ x2 - b) / np.linalg.norm(b) < 1e-12
    import scipy.sparse as sp
    C = (A + sp.eye(A.shape[0], k=2, format="csc") * 1e-3 + sp.eye(A.shape[0], k=-2, format="csc") * 1e-3).tocsc()
    C.sort_indices()
    g.partition(S.symmetric_graph(C))
    g.assemble(C)                                  # other pattern (and partition): analysed again
    g.factorize()
    x3 = g.solve(b)
    assert np.linalg.norm(C @ x3 - b) / np.linalg.norm(b) < 1e-11


def test_solve_matches_oracle_when_factors_match():
    """Same factor (exact config) => GPU replay of the recorded operations equals the oracle's to rounding."""
    A, g, o = _pair(15, 3, 5, 0.0)
    g.assemble(A)
    o.assemble(A)
    g.factorize()
    o.factorize()
    for seed in (1, 2, 3):
        b = S.random(A.shape[0], seed)
        xg, xo = g.solve(b), o.solve(b)
        assert np.linalg.norm(xg - xo) <= 1e-12 * np.linalg.norm(xo)


def test_repro_bit_identical():
    # reference tests/tests.cpp:917-994 (ApproxTest.Repro): repeated factorizations give bit-identical x
    A = S.neglapl(16, 2)
    b = S.random(256, 2019)
    g = S.Tree(4)
    g.set_tol(1e-2)
    g.partition(A)
    ref = None
    for _ in range(5):
        g.assemble(A)
        g.factorize()
        x = g.solve(b)
        if ref is None:
            ref = x
        assert np.array_equal(ref, x)
    A3 = S.neglapl(12, 3)
    b3 = S.random(12**3, 7)
    outs = []
    for _ in range(3):
        t = S.Tree(5)
        t.set_tol(1e-2)
        t.partition(A3)
        t.assemble(A3)
        t.factorize()
        outs.append(t.solve(b3))
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


def test_non_spd_pivot_raises():
    # src/tree.cpp:587-590 -> "Error: Non-SPD Pivot"
    A = -S.neglapl(8, 2)
    g = S.Tree(3)
    g.partition(S.symmetric_graph(A))
    g.assemble(A.tocsc())
    with pytest.raises(RuntimeError, match="Non-SPD"):
        g.factorize()


def test_edge_cases():
    # one level (dense Cholesky of everything), N smaller than a block, tol >= 1 (everything dropped)
    A = S.neglapl(3, 2)
    b = S.random(9, 1)
    for L in (1, 2):
        g = S.Tree(L)
        g.set_tol(0.0)
        g.partition(A)
        g.assemble(A)
        g.factorize()
        x = g.solve(b)
        assert np.linalg.norm(A @ x - b) / np.linalg.norm(b) < 1e-13
    A = S.neglapl(10, 2)
    g = S.Tree(4)
    g.set_tol(10.0)
    o = O.OracleTree(4, tol=10.0)
    g.partition(A)
    o.partition(A)
    g.assemble(A)
    o.assemble(A)
    g.factorize()
    o.factorize()
    assert np.array_equal(g.stats()[2], o.stats()[2]) and g.nnz() == o.nnz()
    b = S.random(100, 3)
    assert np.allclose(g.solve(b), o.solve(b), rtol=1e-12, atol=1e-14)


def test_large_blocks_exercise_tiled_kernels():
    """Few levels on a 3-D grid => pivots of several hundred rows: blocked POTRF/TRSM and the DMMA GEMM tiles."""
    A, g, o = _pair(14, 3, 2, 0.0)
    g.assemble(A)
    o.assemble(A)
    g.factorize()
    o.factorize()
    assert max(g.stats()[1]) > 150
    b = S.random(A.shape[0], 5)
    xg, xo = g.solve(b), o.solve(b)
    assert np.linalg.norm(A @ xg - b) / np.linalg.norm(b) < 1e-12
    assert np.linalg.norm(xg - xo) <= 1e-11 * np.linalg.norm(xo)


# ---------------------------------------------------------------------------------------------------------------------
# GEN / PLU path (reference src/tree.cpp:614-689, :929-956; tests/tests.cpp:98-109 builds the unsymmetric matrices as
# Laplacian + U(-0.1, 0.1) noise on every stored non-zero)
# ---------------------------------------------------------------------------------------------------------------------
def _unsym(A, seed):
    rng = np.random.RandomState(seed)
    B = A.copy().tocsc()
    B.data = B.data + rng.uniform(-0.1, 0.1, B.nnz)
    return B


def _pair_gen(n, d, L, tol, skip=0, seed=7):
    A = _unsym(S.neglapl(n, d), seed)
    X = S.linspace_nd(n, d)
    g = S.Tree(L)
    g.set_tol(tol)
    g.set_skip(skip)
    g.set_symm_kind(S.GEN)
    g.set_scaling_kind(S.PLU)
    g.set_use_geo(True)
    g.set_Xcoo(X)
    o = O.OracleTree(L, tol=tol, skip=skip, symm_kind=O.GEN, scaling_kind=O.PLU)
    o.set_coords(X)
    G = S.symmetric_graph(A)
    g.partition(G)
    o.partition(G)
    return A, g, o


@pytest.mark.parametrize("n,d,L,tol", [(16, 2, 4, 1e-2), (8, 3, 3, 1e-2), (12, 3, 4, 0.0)])
def test_plu_stop_points(n, d, L, tol):
    """Trailing matrix after eliminate / scale / sparsify / merge of the first levels, GPU vs oracle (GEN, PLU)."""
    A, g, o = _pair_gen(n, d, L, tol)
    assert (g.get_assembly_perm() == o.perm()).all()
    for lvl in range(min(L, 2)):
        for ph in range(4):
            g.set_stop(lvl, ph)
            o.set_stop(lvl, ph)
            g.assemble(A)
            o.partition(S.symmetric_graph(A))
            o.assemble(A)
            g.factorize()
            o.factorize()
            Tg, To = g.get_trailing_mat(), o.trailing_mat()
            assert Tg.shape == To.shape
            ranks_equal = (g.stats()[2] == o.stats()[2]).all()
            if ranks_equal:
                err = abs(Tg - To).max() if Tg.nnz + To.nnz else 0.0
                ref = abs(To).max() if To.nnz else 1.0
                # after an RRQR the signs/rounding of R may differ slightly; before it the match is to rounding
                lim = 1e-10 if (lvl == 0 and ph < 2) or tol == 0.0 else 1e-6
                assert err <= lim * max(ref, 1.0), (lvl, ph, err, ref)
            assert g.nnz() == o.nnz() or not ranks_equal


@pytest.mark.parametrize("n,d,L,tol,skip", [(5, 2, 3, 1e-14, 0), (10, 2, 4, 1e-14, 4), (20, 2, 5, 0.0, 1000),
                                            (5, 3, 3, 0.0, 0), (12, 3, 5, 1e-14, 0)])
def test_plu_exact_residual(n, d, L, tol, skip):
    """ApproxTest.Exact (tests/tests.cpp:562-609) for (GEN, PLU): one solve gives |Ax-b|/|b| <= 1e-10."""
    A, g, o = _pair_gen(n, d, L, tol, skip)
    g.assemble(A)
    g.factorize()
    b = S.random(A.shape[0], 2019)
    x = g.solve(b)
    assert np.linalg.norm(A @ x - b) / np.linalg.norm(b) <= 1e-10


@pytest.mark.parametrize("n,d,L,tol", [(32, 2, 5, 1e-2), (16, 3, 5, 1e-2), (20, 3, 6, 1e-3)])
def test_plu_full_factorization_vs_oracle(n, d, L, tol):
    A, g, o = _pair_gen(n, d, L, tol)
    g.assemble(A)
    o.assemble(A)
    g.factorize()
    o.factorize()
    diff, ndiff = _rank_report(g, o)
    assert abs(g.nnz() - o.nnz()) <= 0.005 * o.nnz()
    lg, lo = g.log(), o.log()
    assert np.allclose(lg["dofs_left_spars"], lo["dofs_left_spars"], rtol=0.02, atol=4)
    b = S.random(A.shape[0], 2019)
    xg, xo = g.solve(b), o.solve(b)
    rg = np.linalg.norm(A @ xg - b) / np.linalg.norm(b)
    ro = np.linalg.norm(A @ xo - b) / np.linalg.norm(b)
    assert rg <= 200 * tol and abs(rg - ro) <= 0.25 * ro


def test_plu_large_pivots_blocked_getrf():
    """Uncompressed top separators (576 and 24-wide strips) go through the blocked GETRF / laswp / TRSM_LLU path."""
    A, g, o = _pair_gen(24, 3, 3, 0.0)
    g.assemble(A)
    o.assemble(A)
    g.factorize()
    o.factorize()
    b = S.random(A.shape[0], 2019)
    xg, xo = g.solve(b), o.solve(b)
    assert np.linalg.norm(A @ xg - b) / np.linalg.norm(b) <= 1e-10
    assert np.linalg.norm(xg - xo) / np.linalg.norm(xo) <= 1e-9
    assert g.nnz() == o.nnz()


def test_plu_singular_pivot_raises():
    """getf_cluster throws 'Error: Singular Pivot' (src/tree.cpp:625-628)."""
    import scipy.sparse as sp
    n = 8
    A = S.neglapl(n, 2).tolil()
    A[0, :] = 0.0
    A[:, 0] = 0.0
    A = sp.csc_matrix(A)
    X = S.linspace_nd(n, 2)
    g = S.Tree(2)
    g.set_tol(0.0)
    g.set_symm_kind(S.GEN)
    g.set_scaling_kind(S.PLU)
    g.set_use_geo(True)
    g.set_Xcoo(X)
    g.partition(S.symmetric_graph(S.neglapl(n, 2)))
    g.assemble(A)
    with pytest.raises(RuntimeError, match="Singular Pivot"):
        g.factorize()


# ---- GMRES (src/is.cpp:123-300) on the GPU vs the oracle's restatement ----
@pytest.mark.parametrize("n,d,L,tol,restart", [(32, 2, 5, 1e-2, 100), (16, 3, 5, 1e-2, 3), (20, 2, 4, 0.0, 100)])
def test_gmres_spd_matches_oracle(n, d, L, tol, restart):
    """Iteration count within +-1 of the oracle and the same solution (SPD matrix, LLT preconditioner)."""
    A, g, o = _pair(n, d, L, tol)
    g.assemble(A)
    o.assemble(A)
    g.factorize()
    o.factorize()
    b = S.random(A.shape[0], 2019)
    itg, xg = g.gmres(A, b, 100, restart, 1e-12)
    ito, xo = o.gmres(A, b, 100, restart, 1e-12)
    assert abs(itg - ito) <= 1, (itg, ito)
    assert np.linalg.norm(A @ xg - b) / np.linalg.norm(b) <= 1e-10
    assert np.linalg.norm(xg - xo) / np.linalg.norm(xo) <= 1e-8


@pytest.mark.parametrize("n,L,tol,restart", [(12, 5, 1e-2, 100), (16, 6, 1e-2, 4)])
def test_gmres_c5_family_matches_oracle(n, L, tol, restart):
    """Config C5's matrix family (anisotropic convection-diffusion, non-symmetric), GEN + PLU, GMRES(restart)."""
    A = S.aniso_convdiff(n)
    X = S.linspace_nd(n, 3)
    g = S.Tree(L)
    g.set_tol(tol)
    g.set_symm_kind(S.GEN)
    g.set_scaling_kind(S.PLU)
    g.set_use_geo(True)
    g.set_Xcoo(X)
    o = O.OracleTree(L, tol=tol, symm_kind=O.GEN, scaling_kind=O.PLU)
    o.set_coords(X)
    G = S.symmetric_graph(A)
    g.partition(G)
    o.partition(G)
    g.assemble(A)
    o.assemble(A)
    g.factorize()
    o.factorize()
    b = S.random(A.shape[0], 2019)
    itg, xg = g.gmres(A, b, 200, restart, 1e-12)
    ito, xo = o.gmres(A, b, 200, restart, 1e-12)
    assert abs(itg - ito) <= 1, (itg, ito)
    rg = np.linalg.norm(A @ xg - b) / np.linalg.norm(b)
    ro = np.linalg.norm(A @ xo - b) / np.linalg.norm(b)
    assert rg <= max(10 * ro, 1e-9), (rg, ro)


def test_gmres_edge_cases():
    """Zero right-hand side returns x = 0 (is.cpp:137-141); iteration cap is honoured (is.cpp:231)."""
    A, g, o = _pair(16, 2, 4, 1e-1)
    g.assemble(A)
    o.assemble(A)
    g.factorize()
    o.factorize()
    N = A.shape[0]
    it, x = g.gmres(A, np.zeros(N), 10, 5, 1e-12, x0=np.ones(N))
    assert it == 1 and not x.any()
    b = S.random(N, 2019)
    itg, xg = g.gmres(A, b, 3, 100, 1e-14)
    ito, xo = o.gmres(A, b, 3, 100, 1e-14)
    assert itg == ito == 3
    assert np.linalg.norm(xg - xo) / np.linalg.norm(xo) <= 1e-8


def test_plu_row_pivoting_matters():
    """A pivot block whose diagonal does not dominate: dgetrf really permutes rows, so ScalingPLUQ's P (row
    permutation of the in-edges, src/tree.cpp:838-853, and of the right-hand side, operations.cpp) is exercised.
    The diagonally dominant matrices of the other PLU tests never pivot."""
    n, L = 16, 4
    A = S.neglapl(n, 2).tocsc().astype(np.float64)
    rng = np.random.RandomState(11)
    A.data = A.data + rng.uniform(-0.3, 0.3, A.nnz)
    A.setdiag(rng.uniform(0.3, 0.8, A.shape[0]) * rng.choice([-1.0, 1.0], A.shape[0]))  # weak, sign-mixed diagonal
    A = A.tocsc()
    X = S.linspace_nd(n, 2)
    G = S.symmetric_graph(A)
    g = S.Tree(L)
    g.set_tol(0.0)
    g.set_symm_kind(S.GEN)
    g.set_scaling_kind(S.PLU)
    g.set_use_geo(True)
    g.set_Xcoo(X)
    o = O.OracleTree(L, tol=0.0, symm_kind=O.GEN, scaling_kind=O.PLU)
    o.set_coords(X)
    g.partition(G)
    o.partition(G)
    g.assemble(A)
    o.assemble(A)
    g.factorize()
    o.factorize()
    b = S.random(A.shape[0], 2019)
    xg, xo = g.solve(b), o.solve(b)
    ro = np.linalg.norm(A @ xo - b) / np.linalg.norm(b)
    rg = np.linalg.norm(A @ xg - b) / np.linalg.norm(b)
    assert ro <= 1e-8, ro  # the oracle solves it: the test matrix is usable
    assert rg <= max(100 * ro, 1e-9), (rg, ro)
    assert np.linalg.norm(xg - xo) / np.linalg.norm(xo) <= 1e-6


@pytest.mark.parametrize("n,d,L,tol", [(20, 2, 4, 0.0), (12, 3, 4, 0.0), (15, 3, 5, 1e-14)])
def test_flop_tuple_log_matches_oracle(n, d, L, tol):
    """set_monitor_flops / write_log_flops (reference src/tree.cpp:60-77, pushes at :592,648,662,792,1312): the
    (level, kind, rows, cols, inner) tuples listed from the plan are the ones the oracle pushes call by call."""
    A, g, o = _pair(n, d, L, tol)
    g.set_monitor_flops(True)
    o.set_monitor_flops(True)
    g.assemble(A)
    o.assemble(A)
    g.factorize()
    o.factorize()
    assert np.array_equal(g.stats()[2], o.stats()[2])  # same ranks on these configurations
    fg, fo = g.flops_log(), o.flops_log()
    import collections
    cg_, co_ = collections.Counter(map(tuple, fg.tolist())), collections.Counter(map(tuple, fo.tolist()))
    only_g, only_o = cg_ - co_, co_ - cg_
    assert not only_g and not only_o, f"only in the product's log: {dict(only_g)}; only in the oracle's: {dict(only_o)}"
    assert len(fg) > 0
