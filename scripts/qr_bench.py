"""Micro-benchmark of the sparsification kernels on synthetic panels of the shapes C4 produces
(SPAND_QR_COPIES hook of spand_geqp3_truncated: `copies` replicas of one task in a single launch).
usage: python scripts/qr_bench.py  [prints one QRBENCH line per variant on stderr]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spand_public_b200 as S


def matrix(rows, cols, decay, seed):
    rng = np.random.default_rng(seed)
    r = min(rows, cols)
    U, _ = np.linalg.qr(rng.standard_normal((rows, r)))
    Vt, _ = np.linalg.qr(rng.standard_normal((cols, r)))
    return ((U * decay ** np.arange(r)) @ Vt.T) * (0.05 + rng.random(cols))


CASES = [
    # rows, cols, decay (rank ~ 0.6 rows at tol 1e-2), nsrc, [(copies, kwargs)...]; hot = capacity per CTA
    (108, 917, 0.93, 20, [(1, dict(G=1, hot=128)), (148, dict(G=1, hot=128)), (296, dict(G=1, hot=128)),
                          (148, dict(G=1, hot=64)), (148, dict(G=1, hot=128, theta=0.5)),
                          (148, dict(G=4, nthreads=256, in_smem=False, nb=8, theta=0.5))]),
    (150, 1200, 0.95, 24, [(1, dict(G=1, hot=96)), (148, dict(G=1, hot=96)), (74, dict(G=2, hot=96)),
                           (37, dict(G=4, hot=96)), (148, dict(G=1, hot=48)), (148, dict(G=1, hot=96, theta=0.5)),
                           (148, dict(G=4, nthreads=256, in_smem=False, nb=8, theta=0.5))]),
    (245, 1760, 0.97, 40, [(1, dict(G=1, hot=64)), (1, dict(G=8, hot=64)), (44, dict(G=2, hot=64)),
                           (32, dict(G=4, hot=64)), (16, dict(G=8, hot=64)), (32, dict(G=4, hot=32)),
                           (32, dict(G=4, hot=64, theta=0.5)),
                           (44, dict(G=16, nthreads=512, in_smem=False, nb=16, theta=0.5))]),
    (400, 2400, 0.982, 100, [(1, dict(G=1, hot=32)), (1, dict(G=4, hot=32)), (1, dict(G=16, hot=32)),
                             (8, dict(G=16, hot=32)), (16, dict(G=8, hot=32)), (8, dict(G=16, hot=16)),
                             (8, dict(G=16, hot=32, theta=0.5)),
                             (8, dict(G=16, nthreads=512, in_smem=False, nb=16, theta=0.5))]),
    (630, 2600, 0.988, 150, [(1, dict(G=16, hot=20)), (4, dict(G=16, hot=20)), (8, dict(G=16, hot=20)),
                             (4, dict(G=16, nthreads=512, in_smem=False, nb=16, theta=0.5))]),
]
only = sys.argv[1:]
for rows, cols, decay, nsrc, variants in CASES:
    if only and str(rows) not in only:
        continue
    A = matrix(rows, cols, decay, rows)
    for vi, (copies, kw) in enumerate(variants):
        if os.environ.get("QRB_VARIANT") and int(os.environ["QRB_VARIANT"]) != vi:
            continue
        os.environ["SPAND_QR_COPIES"] = str(copies)
        kw = dict(kw)
        kw.setdefault("theta", 0.25 if "hot" in kw else 0.5)
        import ctypes as C
        lib = S.lib()
        h = (C.c_ulonglong * 16)()
        lib.spand_debug_hc2_stats.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
        lib.spand_debug_hc2_stats.restype = None
        lib.spand_debug_hc2_stats(h, 1)
        try:
            S.geqp3_truncated(A, 1e-2, nsrc=nsrc, **kw)
        except Exception as ex:
            print("QRBENCH failed", rows, cols, kw, ex, file=sys.stderr)
            continue
        lib.spand_debug_hc2_stats(h, 1)
        h = list(h)
        if h[1]:
            ntask = copies * 2 + 1
            print(f"   hc2 stats per task: steps {h[0]/ntask:.0f} blocks {h[1]/ntask:.1f} early {h[2]/ntask:.1f} "
                  f"hot/block {h[3]/h[1]:.1f} | kcycles per task: boundary {h[6]/ntask/1e3:.0f} hot loop {h[7]/ntask/1e3:.0f} "
                  f"block end {h[8]/ntask/1e3:.0f} gather+scatter {h[9]/ntask/1e3:.0f}", file=sys.stderr)
