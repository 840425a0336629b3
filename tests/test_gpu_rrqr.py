"""Kernel-level parity of the batched truncated QRCP (spand_geqp3_truncated) against LAPACK dgeqp3 + the reference's
choose_rank (src/util.cpp:383-452) + triu(R[:rank]) P^T (src/tree.cpp:1334-1335), on matrices without pivot ties, for
every launch shape: panel in shared / distributed shared / global memory, hot/cold and full-sweep variants, cluster
widths 1..16, gathered from plain and transposed source blocks.

Tolerance: pivots are unique here (random matrices with geometric singular value decay), so ranks must be EQUAL and
R[:rank] P^T must agree to 1e-11 relative (FP64, different summation orders only)."""
import numpy as np
import pytest
import scipy.linalg as sla

import oracle_lib as O
import spand_public_b200 as S

pytestmark = pytest.mark.gpu


def _matrix(rows, cols, decay, seed):
    rng = np.random.default_rng(seed)
    U, _ = np.linalg.qr(rng.standard_normal((rows, rows)))
    Vt, _ = np.linalg.qr(rng.standard_normal((cols, rows)))
    s = decay ** np.arange(rows)
    A = (U * s) @ Vt.T
    # uneven column scaling: a realistic spread of column norms (near and far neighbours)
    return A * (0.05 + rng.random(cols))


def _reference(A, tol):
    R, piv = sla.qr(A, mode="r", pivoting=True)
    mn = min(A.shape)
    d = np.ascontiguousarray(np.diag(R)[:mn])
    rank = O.lib().orc_choose_rank(d, mn, tol)
    out = np.zeros((rank, A.shape[1]))
    out[:, piv] = np.triu(R[:rank, :])
    return rank, out, piv


SHAPES = [
    # rows, cols, decay, tol, kwargs
    (24, 200, 0.7, 1e-2, dict(G=1, nthreads=128, in_smem=True, nb=16)),
    (60, 480, 0.85, 1e-2, dict(G=2, nthreads=256, in_smem=True, nb=16)),
    (60, 480, 0.85, 1e-2, dict(G=1, nthreads=256, in_smem=False, nb=8, theta=0.0)),
    (60, 480, 0.85, 1e-2, dict(G=1, nthreads=256, in_smem=False, nb=8, theta=0.5)),
    (61, 483, 0.85, 1e-2, dict(G=1, nthreads=256, in_smem=False, nb=8, theta=0.5, nsrc=3)),
    (61, 483, 0.85, 1e-2, dict(G=2, nthreads=256, in_smem=False, nb=8, theta=0.5, nsrc=3, transposed=True)),
    (150, 1200, 0.93, 1e-2, dict(G=1, nthreads=256, in_smem=False, nb=8, theta=0.5)),
    (150, 1200, 0.93, 1e-2, dict(G=4, nthreads=256, in_smem=False, nb=8, theta=0.5, nsrc=4)),
    (150, 1200, 0.93, 1e-2, dict(G=4, nthreads=256, in_smem=False, nb=8, theta=0.25)),
    (150, 1200, 0.93, 1e-2, dict(G=4, nthreads=256, in_smem=False, nb=8, theta=0.9)),
    (150, 1200, 0.93, 1e-2, dict(G=8, nthreads=256, in_smem=False, nb=4, theta=0.5)),
    (245, 1760, 0.96, 1e-2, dict(G=4, nthreads=256, in_smem=False, nb=8, theta=0.5, nsrc=5, transposed=True)),
    (245, 1760, 0.96, 1e-2, dict(G=16, nthreads=256, in_smem=False, nb=8, theta=0.5)),
    (400, 2400, 0.975, 1e-2, dict(G=8, nthreads=512, in_smem=False, nb=16, theta=0.5)),
    (400, 2400, 0.975, 1e-2, dict(G=16, nthreads=512, in_smem=False, nb=16, theta=0.5, nsrc=2)),
    (400, 2400, 0.975, 1e-2, dict(G=16, nthreads=512, in_smem=False, nb=16, theta=0.0)),
    (629, 2600, 0.985, 1e-2, dict(G=16, nthreads=512, in_smem=False, nb=16, theta=0.5)),
    (120, 90, 0.9, 1e-3, dict(G=1, nthreads=256, in_smem=False, nb=8, theta=0.5)),     # cols < rows
    (64, 700, 0.999, 1e-2, dict(G=2, nthreads=256, in_smem=False, nb=8, theta=0.5)),   # full rank: nothing happens
    (100, 800, 0.9, 0.0, dict(G=2, nthreads=256, in_smem=False, nb=8, theta=0.5)),     # tol = 0: full factorization
    (100, 800, 0.5, 1e-6, dict(G=2, nthreads=256, in_smem=False, nb=8, theta=0.5)),    # fast decay, small rank
]


@pytest.mark.parametrize("rows,cols,decay,tol,kw", SHAPES)
def test_truncated_qrcp_matches_lapack(rows, cols, decay, tol, kw):
    A = _matrix(rows, cols, decay, 7 * rows + cols)
    rank_ref, R_ref, piv = _reference(A, tol)
    rank, R, V, tau = S.geqp3_truncated(A, tol, **kw)
    assert rank == rank_ref, (rank, rank_ref)
    if rank_ref >= rows:
        assert R is None
        return
    scale = np.abs(R_ref).max()
    assert np.abs(R - R_ref).max() <= 1e-11 * scale
    # the Orthogonal op (src/tree.cpp:1322-1331): Q from (V, tau) reproduces A = Q [R; *]
    Q = np.eye(rows)
    for k in range(rank - 1, -1, -1):
        v = np.zeros(rows)
        v[k] = 1.0
        v[k + 1:] = V[k + 1:, k]
        Q = Q - tau[k] * np.outer(v, v @ Q)
    top = (Q.T @ A)[:rank, :]
    assert np.abs(top - R_ref).max() <= 1e-11 * scale


def test_hot_cold_equals_full_sweep_on_tied_columns():
    """Exact ties (duplicated columns): LAPACK's first-index rule must survive the hot / cold split."""
    rows, cols = 80, 600
    A = _matrix(rows, cols // 2, 0.9, 3)
    A = np.concatenate([A, A], axis=1)
    r0, R0, _, _ = S.geqp3_truncated(A, 1e-2, G=2, nthreads=256, in_smem=False, nb=8, theta=0.0)
    r1, R1, _, _ = S.geqp3_truncated(A, 1e-2, G=2, nthreads=256, in_smem=False, nb=8, theta=0.5)
    rank_ref, R_ref, _ = _reference(A, 1e-2)
    assert r0 == r1 == rank_ref
    assert np.abs(R1 - R0).max() <= 1e-11 * np.abs(R0).max()
