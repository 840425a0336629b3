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
    r = min(rows, cols)
    U, _ = np.linalg.qr(rng.standard_normal((rows, r)))
    Vt, _ = np.linalg.qr(rng.standard_normal((cols, r)))
    s = decay ** np.arange(r)
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
    # column kernel (rows <= 64, one thread per column, panel in the shared memory of one CTA)
    (9, 116, 0.6, 1e-2, dict(col=True, nthreads=128)),
    (9, 116, 0.95, 1e-2, dict(col=True, nthreads=128, nsrc=4)),           # full rank: nothing happens
    (17, 200, 0.8, 1e-2, dict(col=True, nthreads=128, nsrc=5, transposed=True)),
    (31, 516, 0.85, 1e-2, dict(col=True, nthreads=256, nsrc=3)),
    (47, 462, 0.88, 1e-2, dict(col=True, nthreads=256, nsrc=6, transposed=True)),
    (64, 380, 0.9, 1e-2, dict(col=True, nthreads=256)),
    (49, 692, 0.88, 1e-2, dict(col="gp", nthreads=512, nsrc=4)),           # panel in L2 (does not fit shared memory)
    (89, 948, 0.93, 1e-2, dict(col="gp", nthreads=512, nsrc=6, transposed=True)),
    (128, 700, 0.95, 1e-2, dict(col="gp", nthreads=256)),
    (100, 200, 0.9, 1e-2, dict(col=True, nthreads=256)),
    (40, 30, 0.8, 1e-3, dict(col=True, nthreads=128)),                     # cols < rows
    (33, 300, 0.8, 0.0, dict(col=True, nthreads=256)),                     # tol = 0
    # hot-set kernel (rrqr_hc2.cu): capacities from "everything hot" down to 1, all cluster widths, all row classes
    (60, 480, 0.85, 1e-2, dict(G=1, hot=64)),
    (60, 480, 0.85, 1e-2, dict(G=1, hot=300)),
    (61, 483, 0.85, 1e-2, dict(G=2, hot=24, nsrc=3, transposed=True)),
    (61, 483, 0.85, 1e-2, dict(G=1, hot=1)),
    (61, 483, 0.85, 1e-2, dict(G=4, hot=2)),
    (150, 1200, 0.93, 1e-2, dict(G=1, hot=48)),
    (150, 1200, 0.93, 1e-2, dict(G=4, hot=48, nsrc=4)),
    (150, 1200, 0.93, 1e-2, dict(G=8, hot=16, theta=0.25)),
    (150, 1200, 0.93, 1e-2, dict(G=2, hot=96, theta=0.9)),
    (245, 1760, 0.96, 1e-2, dict(G=4, hot=40, nsrc=5, transposed=True)),
    (245, 1760, 0.96, 1e-2, dict(G=16, hot=64)),
    (256, 1000, 0.96, 1e-2, dict(G=2, hot=64)),
    (257, 1000, 0.96, 1e-2, dict(G=2, hot=64)),
    (383, 2400, 0.975, 1e-2, dict(G=8, hot=48)),
    (400, 2400, 0.975, 1e-2, dict(G=16, hot=40, nsrc=2)),
    (629, 2600, 0.985, 1e-2, dict(G=16, hot=24)),
    (640, 1300, 0.985, 1e-2, dict(G=8, hot=20)),
    (120, 90, 0.9, 1e-3, dict(G=1, hot=32)),      # cols < rows
    (120, 90, 0.9, 0.0, dict(G=2, hot=32)),       # cols < rows, tol = 0: every column becomes a pivot
    (64, 700, 0.999, 1e-2, dict(G=2, hot=32)),    # full rank: nothing happens
    (100, 800, 0.9, 0.0, dict(G=2, hot=32)),      # tol = 0: full factorization
    (100, 800, 0.5, 1e-6, dict(G=2, hot=32)),     # fast decay, small rank
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
    B = np.concatenate([A[:40, :150], A[:40, :150]], axis=1)  # 40 x 300, every column twice
    rb0, RB0, _, _ = S.geqp3_truncated(B, 1e-2, G=1, nthreads=256, in_smem=True, nb=16)
    rb1, RB1, _, _ = S.geqp3_truncated(B, 1e-2, col=True, nthreads=256)
    assert rb0 == rb1 and np.abs(RB1 - RB0).max() <= 1e-11 * np.abs(RB0).max()
    r2, R2, _, _ = S.geqp3_truncated(A, 1e-2, G=2, hot=16)
    r3, R3, _, _ = S.geqp3_truncated(A, 1e-2, G=1, hot=200)
    rank_ref, R_ref, _ = _reference(A, 1e-2)
    assert r0 == r1 == r2 == r3 == rank_ref
    assert np.abs(R1 - R0).max() <= 1e-11 * np.abs(R0).max()
    assert np.abs(R2 - R0).max() <= 1e-11 * np.abs(R0).max()
    assert np.abs(R3 - R0).max() <= 1e-11 * np.abs(R0).max()


def test_hot_set_kernel_cluster_widths_agree():
    """Sub-tree sharding changes the cluster width of a task and with it which columns are hot (applied reflector by
    reflector) or cold (refreshed in compact-WY form): same pivots, results equal to rounding."""
    A = _matrix(150, 1200, 0.93, 11)
    ref = S.geqp3_truncated(A, 1e-2, G=1, hot=40)
    for G in (2, 4, 8, 16):
        out = S.geqp3_truncated(A, 1e-2, G=G, hot=40)
        assert out[0] == ref[0]
        assert np.abs(out[1] - ref[1]).max() <= 1e-12 * np.abs(ref[1]).max()
        assert np.abs(out[2] - ref[2]).max() <= 1e-11
