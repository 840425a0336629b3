"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so) — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = os.path.join(_ROOT, "oracle", "_build", "liboracle.so")

_i = C.c_int
_d = C.c_double
_p = C.c_void_p
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "oracle")])


def _load():
    if not os.path.exists(_LIB):
        build()
    lib = C.CDLL(_LIB)
    lib.orc_create.restype = _p
    lib.orc_create.argtypes = [_i]
    lib.orc_destroy.argtypes = [_p]
    lib.orc_last_error.restype = C.c_char_p
    lib.orc_last_error.argtypes = [_p]
    lib.orc_set_params.argtypes = [_p, _d, _i, _i, _i, _i, _i, _i]
    lib.orc_set_stop.argtypes = [_p, _i, _i]
    lib.orc_set_coords.argtypes = [_p, _i, _i, _dp]
    lib.orc_partition.argtypes = [_p, _i, _ip, _ip]
    lib.orc_assemble.argtypes = [_p, _i, _ip, _ip, _dp]
    lib.orc_factorize.argtypes = [_p]
    lib.orc_solve.argtypes = [_p, _dp]
    lib.orc_cg.argtypes = [_p, _i, _ip, _ip, _dp, _dp, _dp, _i, _d, _i]
    lib.orc_gmres.argtypes = [_p, _i, _ip, _ip, _dp, _dp, _dp, _i, _i, _d, _i]
    lib.orc_nnz.restype = C.c_longlong
    lib.orc_nnz.argtypes = [_p]
    lib.orc_get_stop.argtypes = [_p]
    lib.orc_get_N.argtypes = [_p]
    lib.orc_get_perm.argtypes = [_p, _ip]
    lib.orc_get_partition.argtypes = [_p, _ip, _ip, _ip, _ip, _ip, _ip]
    lib.orc_num_clusters.argtypes = [_p]
    lib.orc_get_stats.argtypes = [_p, _ip, _ip, _ip]
    lib.orc_get_log.argtypes = [_p, _dp]
    lib.orc_trailing.argtypes = [_p, _p, _p, _p]
    lib.orc_choose_rank.argtypes = [_dp, _i, _d]
    lib.orc_swap2perm.argtypes = [_ip, _i, _ip]
    lib.orc_block2dense.argtypes = [_i, _ip, _ip, _dp, _i, _i, _i, _i, _dp, _i]
    lib.orc_set_threads.argtypes = [_i]
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


LOG_FIELDS = ["dofs_nd", "dofs_left_nd", "dofs_left_elim", "dofs_left_spars", "fact_nnz", "rank_before", "rank_after",
              "nspars", "ignored", "nbrs", "t_elim", "t_scale", "t_spars", "t_merge", "fl_pivot", "fl_panel", "fl_schur",
              "fl_rrqr_rank", "fl_rrqr_full", "by_scale", "by_rrqr", "by_merge"]

SPD, SYM, GEN = 0, 1, 2
LLT, PLU = 0, 3


def _csc(A):
    A = A.tocsc()
    A.sort_indices()
    return (A.shape[0], np.ascontiguousarray(A.indptr, dtype=np.int32), np.ascontiguousarray(A.indices, dtype=np.int32),
            np.ascontiguousarray(A.data, dtype=np.float64))


class OracleTree:
    """Mirror of spaND::Tree (include/tree.h:130-198) over the CPU oracle."""

    def __init__(self, nlevels, tol=10.0, skip=0, symm_kind=SPD, scaling_kind=LLT, use_geo=False, verb=False,
                 use_sparsify=True):
        self._l = lib()
        self.nlevels = nlevels
        self._h = self._l.orc_create(nlevels)
        if not self._h:
            raise RuntimeError("orc_create failed")
        self.p = dict(tol=tol, skip=skip, symm_kind=symm_kind, scaling_kind=scaling_kind, use_geo=use_geo, verb=verb,
                      use_sparsify=use_sparsify)
        self._push()

    def _push(self):
        p = self.p
        self._l.orc_set_params(self._h, p["tol"], p["skip"], p["symm_kind"], p["scaling_kind"], int(p["use_geo"]),
                               int(p["verb"]), int(p["use_sparsify"]))

    def set(self, **kw):
        self.p.update(kw)
        self._push()

    def set_stop(self, level, phase):
        self._l.orc_set_stop(self._h, level, phase)

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.orc_destroy(self._h)
            self._h = None

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self._l.orc_last_error(self._h).decode())

    def set_coords(self, X):
        X = np.asfortranarray(X, dtype=np.float64)
        dim, N = X.shape
        self._check(self._l.orc_set_coords(self._h, dim, N, np.ascontiguousarray(X.T).ravel()))
        self.set(use_geo=True)

    def partition(self, A):
        N, cp, ri, _ = _csc(A)
        self._check(self._l.orc_partition(self._h, N, cp, ri))
        self.N = N

    def assemble(self, A):
        N, cp, ri, v = _csc(A)
        self._check(self._l.orc_assemble(self._h, N, cp, ri, v))

    def factorize(self):
        self._check(self._l.orc_factorize(self._h))

    def solve(self, b):
        x = np.ascontiguousarray(b, dtype=np.float64).copy()
        self._check(self._l.orc_solve(self._h, x))
        return x

    def cg(self, A, b, iters=100, tol=1e-12, verb=False):
        N, cp, ri, v = _csc(A)
        x = np.zeros(N)
        it = self._l.orc_cg(self._h, N, cp, ri, v, np.ascontiguousarray(b, dtype=np.float64), x, iters, tol, int(verb))
        if it < 0:
            raise RuntimeError(self._l.orc_last_error(self._h).decode())
        return it, x

    def gmres(self, A, b, iters=100, restart=100, tol=1e-12, verb=False):
        """src/is.cpp:123-300; x0 = 0, returns (iterations, x)."""
        N, cp, ri, v = _csc(A)
        x = np.zeros(N)
        it = self._l.orc_gmres(self._h, N, cp, ri, v, np.ascontiguousarray(b, dtype=np.float64), x, iters, restart,
                               tol, int(verb))
        if it < 0:
            raise RuntimeError(self._l.orc_last_error(self._h).decode())
        return it, x

    def nnz(self):
        return self._l.orc_nnz(self._h)

    def get_stop(self):
        return self._l.orc_get_stop(self._h)

    def perm(self):
        p = np.zeros(self.N, dtype=np.int32)
        self._l.orc_get_perm(self._h, p)
        return p

    def partition_ids(self):
        a = [np.zeros(self.N, dtype=np.int32) for _ in range(6)]
        self._l.orc_get_partition(self._h, *a)
        return a

    def stats(self):
        n = self._l.orc_num_clusters(self._h)
        a = [np.zeros(n, dtype=np.int32) for _ in range(3)]
        self._l.orc_get_stats(self._h, *a)
        return a

    def log(self):
        out = np.zeros(self.nlevels * len(LOG_FIELDS))
        self._l.orc_get_log(self._h, out)
        out = out.reshape(self.nlevels, len(LOG_FIELDS))
        return {k: out[:, i].copy() for i, k in enumerate(LOG_FIELDS)}

    def set_monitor_flops(self, on):
        self._l.orc_set_monitor_flops.argtypes = [_p, _i]
        self._l.orc_set_monitor_flops(self._h, int(on))

    def flops_log(self):
        fn = self._l.orc_get_flops_log
        fn.restype = C.c_longlong
        fn.argtypes = [_p, C.c_void_p]
        n = fn(self._h, None)
        out = np.zeros((n, 5), dtype=np.int64)
        if n:
            fn(self._h, out.ctypes.data)
        return out

    def trailing_mat(self):
        import scipy.sparse as sp
        nnz = self._l.orc_trailing(self._h, None, None, None)
        cp = np.zeros(self.N + 1, dtype=np.int32)
        ri = np.zeros(nnz, dtype=np.int32)
        v = np.zeros(nnz)
        self._l.orc_trailing(self._h, cp.ctypes.data, ri.ctypes.data, v.ctypes.data)
        return sp.csc_matrix((v, ri, cp), shape=(self.N, self.N))
