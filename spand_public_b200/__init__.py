"""spand_public_b200 — B200-native spaND factorization path.

Python is only the harness here: it loads the C-ABI shared library (include/spand_b200.h, built from
spand_public_b200/csrc by ``make``) and mirrors the reference's ``spaND::Tree`` call contract
(reference include/tree.h:130-198): setters -> partition -> assemble -> factorize -> solve / cg.
There is no CPU fallback: without the CUDA library or without a GPU every compute call raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SPAND_B200_BUILD selects an alternative build directory (kernel tuning experiments: scripts/rrqr_tune.py)
LIB_PATH = os.path.join(_HERE, os.environ.get("SPAND_B200_BUILD", "_build"), "libspand_b200.so")

SPD, SYM, GEN = 0, 1, 2
LLT, PLU = 0, 3

_i, _d, _p = C.c_int, C.c_double, C.c_void_p
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")

# every symbol include/spand_b200.h declares
EXPORTS = [
    "spand_create", "spand_destroy", "spand_last_error", "spand_set_tol", "spand_set_skip", "spand_set_symm_kind",
    "spand_set_scaling_kind", "spand_set_use_geo", "spand_set_verb", "spand_set_use_sparsify", "spand_set_device",
    "spand_set_coords", "spand_set_stop", "spand_partition", "spand_get_partition", "spand_get_perm", "spand_get_N",
    "spand_assemble", "spand_factorize", "spand_solve", "spand_solve_device", "spand_cg", "spand_gmres", "spand_debug_rrqr_phases", "spand_debug_hc2_stats", "spand_geqp3_truncated", "spand_nnz", "spand_get_stop",
    "spand_get_nlevels", "spand_num_clusters", "spand_get_stats", "spand_log_fields", "spand_log_field_name",
    "spand_get_log", "spand_factorize_seconds", "spand_analyze_seconds", "spand_plan_analyze", "spand_plan_live_edges",
    "spand_plan_counts", "spand_get_cluster_layout", "spand_mg_setup", "spand_mg_get_handle", "spand_mg_set_peers",
    "spand_mg_owner_map", "spand_kernel_launches", "spand_arena_bytes", "spand_trailing",
    "spand_util_random", "spand_util_linspace_nd", "spand_util_neglapl", "spand_util_aniso", "spand_util_mm_read",
    "spand_util_mm_read_dense", "spand_set_profile", "spand_set_monitor_flops", "spand_get_flops_log", "spand_num_families", "spand_family_name", "spand_get_family_stats",
]


def build():
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(_HERE, "csrc")])


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `make -C spand_public_b200/csrc` (or __graft_entry__.build()). "
            "There is no CPU fallback for the spaND factorization path.")
    L = C.CDLL(LIB_PATH)
    L.spand_create.restype = _p
    L.spand_create.argtypes = [_i]
    L.spand_destroy.argtypes = [_p]
    L.spand_last_error.restype = C.c_char_p
    L.spand_last_error.argtypes = [_p]
    L.spand_set_tol.argtypes = [_p, _d]
    for f in ("skip", "symm_kind", "scaling_kind", "use_geo", "verb", "use_sparsify", "device"):
        getattr(L, "spand_set_" + f).argtypes = [_p, _i]
    L.spand_set_coords.argtypes = [_p, _i, _i, _dp]
    L.spand_set_stop.argtypes = [_p, _i, _i]
    L.spand_partition.argtypes = [_p, _i, _ip, _ip]
    L.spand_get_partition.argtypes = [_p] + [_ip] * 6
    L.spand_get_perm.argtypes = [_p, _ip]
    L.spand_get_N.argtypes = [_p]
    L.spand_assemble.argtypes = [_p, _i, _ip, _ip, _dp]
    L.spand_factorize.argtypes = [_p]
    L.spand_solve.argtypes = [_p, _dp]
    L.spand_solve_device.argtypes = [_p, _p]
    L.spand_cg.argtypes = [_p, _i, _ip, _ip, _dp, _dp, _dp, _i, _d, _i, C.POINTER(_d)]
    L.spand_gmres.argtypes = [_p, _i, _ip, _ip, _dp, _dp, _dp, _i, _i, _d, _i, C.POINTER(_d)]
    L.spand_nnz.restype = C.c_longlong
    L.spand_nnz.argtypes = [_p]
    L.spand_get_stop.argtypes = [_p]
    L.spand_get_nlevels.argtypes = [_p]
    L.spand_num_clusters.argtypes = [_p]
    L.spand_get_stats.argtypes = [_p, _ip, _ip, _ip]
    L.spand_log_fields.argtypes = []
    L.spand_log_field_name.restype = C.c_char_p
    L.spand_log_field_name.argtypes = [_i]
    L.spand_get_log.argtypes = [_p, _dp]
    L.spand_factorize_seconds.restype = _d
    L.spand_factorize_seconds.argtypes = [_p]
    L.spand_analyze_seconds.restype = _d
    L.spand_analyze_seconds.argtypes = [_p]
    L.spand_plan_analyze.argtypes = [_p, _i, _ip, _ip]
    L.spand_plan_live_edges.argtypes = [_p, _i, _i, _p, _p]
    L.spand_plan_counts.argtypes = [_p, _i, np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")]
    L.spand_get_cluster_layout.argtypes = [_p, _ip, _ip]
    L.spand_mg_setup.argtypes = [_p, _i, _i, C.c_longlong]
    L.spand_mg_get_handle.argtypes = [_p, C.c_char_p]
    L.spand_mg_set_peers.argtypes = [_p, C.c_char_p]
    L.spand_mg_owner_map.argtypes = [_p, _i, _ip]
    L.spand_kernel_launches.restype = C.c_longlong
    L.spand_kernel_launches.argtypes = [_p]
    L.spand_arena_bytes.restype = C.c_longlong
    L.spand_arena_bytes.argtypes = [_p]
    L.spand_trailing.argtypes = [_p, _p, _p, _p]
    L.spand_set_profile.argtypes = [_p, _i]
    L.spand_num_families.argtypes = []
    L.spand_family_name.restype = C.c_char_p
    L.spand_family_name.argtypes = [_i]
    L.spand_get_family_stats.argtypes = [_p, _dp, np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")]
    L.spand_util_random.argtypes = [_i, _i, _dp]
    L.spand_util_linspace_nd.argtypes = [_i, _i, _dp]
    L.spand_util_neglapl.argtypes = [_i, _i, _p, _p, _p]
    L.spand_util_aniso.argtypes = [_i, _p, _p, _p]
    L.spand_util_mm_read.argtypes = [C.c_char_p, C.POINTER(_i), C.POINTER(_i), _p, _p, _p]
    L.spand_util_mm_read_dense.argtypes = [C.c_char_p, C.POINTER(_i), C.POINTER(_i), _p]
    _lib = L
    return L


def _csc(A):
    A = A.tocsc()
    A.sort_indices()
    return (A.shape[0], np.ascontiguousarray(A.indptr, dtype=np.int32), np.ascontiguousarray(A.indices, dtype=np.int32),
            np.ascontiguousarray(A.data, dtype=np.float64))


def geqp3_truncated(A, tol, nsrc=1, transposed=False, G=1, nthreads=256, in_smem=False, nb=8, theta=0.5, hot=0,
                    col=False):
    """geqp3 + choose_rank + triu(R[:rank]) P^T of one dense matrix through the batched sparsification kernel
    (reference src/util.cpp:383-452, src/tree.cpp:1334-1335). Returns (rank, AsnP or None when rank >= rows, V, tau)."""
    A = np.asarray(A, dtype=np.float64)
    rows, cols = A.shape
    w = cols // nsrc
    blocks = [A[:, i * w:(i + 1) * w] for i in range(nsrc)]
    # device layout: block i column-major rows x w, or its transpose w x rows (column-major) for out-edges
    flat = np.concatenate([(b.T if transposed else b).ravel(order="F") for b in blocks])
    R = np.zeros_like(flat)
    mn = min(rows, cols)
    V = np.zeros(rows * mn)
    tau = np.zeros(mn)
    rank = C.c_int(0)
    fn = lib().spand_geqp3_truncated
    fn.argtypes = [_i, _i, _dp, _i, _i, _d, _i, _i, _i, _i, _d, C.POINTER(C.c_int), _dp, _dp, _dp]
    if hot > 0:  # hot-set kernel (rrqr_hc2.cu): in_smem = 2, nb carries the capacity of the hot set
        in_smem, nb = 2, hot
    if col:  # column kernel (one thread per column): in_smem = 3 (panel in shared memory) / 4 ("gp": panel in L2)
        in_smem = 4 if col == "gp" else 3
    rc = fn(rows, cols, np.ascontiguousarray(flat), nsrc, int(transposed), float(tol), G, nthreads, int(in_smem), nb,
            float(theta), C.byref(rank), R, V, tau)
    if rc != 0:
        raise RuntimeError("spand_geqp3_truncated failed (see stderr)")
    r = rank.value
    out = None
    if r < rows:
        cols_out = []
        for i in range(nsrc):
            blk = R[i * w * rows:(i + 1) * w * rows]
            blk = blk.reshape((w, rows), order="F").T if transposed else blk.reshape((rows, w), order="F")
            cols_out.append(blk[:r, :])
        out = np.concatenate(cols_out, axis=1)
    return r, out, V.reshape((rows, mn), order="F"), tau


# ---- host utilities (src/util.cpp, include/mmio.hpp counterparts) ----
def random(size, seed):
    """mt19937 + uniform_real_distribution(-1,1), reference src/util.cpp:549-558."""
    out = np.zeros(size)
    lib().spand_util_random(size, seed, out)
    return out


def linspace_nd(n, dim):
    """dim x n^dim coordinates, first coordinate slowest (reference src/util.cpp:488-517)."""
    out = np.zeros(dim * n**dim)
    lib().spand_util_linspace_nd(n, dim, out)
    return out.reshape(n**dim, dim).T.copy()


def _two_call_csc(fn, N):
    import scipy.sparse as sp
    nnz = fn(None, None, None)
    cp = np.zeros(N + 1, dtype=np.int32)
    ri = np.zeros(nnz, dtype=np.int32)
    v = np.zeros(nnz)
    fn(cp.ctypes.data, ri.ctypes.data, v.ctypes.data)
    return sp.csc_matrix((v, ri, cp), shape=(N, N))


def neglapl(n, d):
    """Same matrix as the reference's mats/neglapl_<d>_<n>.mm once read (full symmetric storage)."""
    L = lib()
    return _two_call_csc(lambda a, b, c: L.spand_util_neglapl(n, d, a, b, c), n**d)


def aniso_convdiff(n):
    """Config C5 of BASELINE.json (definition: SURVEY.md section 8d)."""
    L = lib()
    return _two_call_csc(lambda a, b, c: L.spand_util_aniso(n, a, b, c), n**3)


def mm_read(fn):
    import scipy.sparse as sp
    L = lib()
    r, c = _i(), _i()
    nnz = L.spand_util_mm_read(fn.encode(), C.byref(r), C.byref(c), None, None, None)
    if nnz < 0:
        raise IOError("cannot read " + fn)
    cp = np.zeros(c.value + 1, dtype=np.int32)
    ri = np.zeros(nnz, dtype=np.int32)
    v = np.zeros(nnz)
    L.spand_util_mm_read(fn.encode(), C.byref(r), C.byref(c), cp.ctypes.data, ri.ctypes.data, v.ctypes.data)
    return sp.csc_matrix((v, ri, cp), shape=(r.value, c.value))


def mm_read_dense(fn):
    L = lib()
    r, c = _i(), _i()
    if L.spand_util_mm_read_dense(fn.encode(), C.byref(r), C.byref(c), None) != 0:
        raise IOError("cannot read " + fn)
    out = np.zeros(r.value * c.value)
    L.spand_util_mm_read_dense(fn.encode(), C.byref(r), C.byref(c), out.ctypes.data)
    return out.reshape(c.value, r.value).T.copy()


def symmetric_graph(A):
    """|A| + |A^T| + I (reference src/util.cpp:47-63)."""
    import scipy.sparse as sp
    A = A.tocsc()
    return (abs(A) + abs(A.T) + sp.identity(A.shape[0], format="csc")).tocsc()


class Tree:
    """Mirror of spaND::Tree (reference include/tree.h:130-198) over the CUDA library."""

    def __init__(self, nlevels):
        self._l = lib()
        self.nlevels = nlevels
        self._h = self._l.spand_create(nlevels)
        if not self._h:
            raise RuntimeError("spand_create failed (nlevels must be > 0)")
        self.N = 0

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.spand_destroy(self._h)
            self._h = None

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self._l.spand_last_error(self._h).decode())

    # setters, include/tree.h:133-147
    def set_verb(self, v): self._l.spand_set_verb(self._h, int(v))
    def set_tol(self, tol): self._l.spand_set_tol(self._h, float(tol))
    def set_skip(self, skip): self._l.spand_set_skip(self._h, int(skip))
    def set_symm_kind(self, k): self._l.spand_set_symm_kind(self._h, int(k))
    def set_scaling_kind(self, k): self._l.spand_set_scaling_kind(self._h, int(k))
    def set_use_geo(self, g): self._l.spand_set_use_geo(self._h, int(g))
    def set_use_sparsify(self, u): self._l.spand_set_use_sparsify(self._h, int(u))
    def set_device(self, d): self._l.spand_set_device(self._h, int(d))
    def set_stop(self, level, phase): self._l.spand_set_stop(self._h, level, phase)

    def set_Xcoo(self, X):
        X = np.asarray(X, dtype=np.float64)
        dim, N = X.shape
        self._check(self._l.spand_set_coords(self._h, dim, N, np.ascontiguousarray(X.T).ravel()))

    def partition(self, A):
        N, cp, ri, _ = _csc(A)
        self._check(self._l.spand_partition(self._h, N, cp, ri))
        self.N = N
        return self.partition_ids()

    def partition_ids(self):
        a = [np.zeros(self.N, dtype=np.int32) for _ in range(6)]
        self._l.spand_get_partition(self._h, *a)
        return a

    def get_assembly_perm(self):
        p = np.zeros(self.N, dtype=np.int32)
        self._l.spand_get_perm(self._h, p)
        return p

    def assemble(self, A):
        N, cp, ri, v = _csc(A)
        self._check(self._l.spand_assemble(self._h, N, cp, ri, v))

    def factorize(self):
        self._check(self._l.spand_factorize(self._h))

    def solve(self, b):
        x = np.ascontiguousarray(b, dtype=np.float64).copy()
        self._check(self._l.spand_solve(self._h, x))
        return x

    def solve_device(self, ptr):
        self._check(self._l.spand_solve_device(self._h, ptr))

    def cg(self, A, b, iters=100, tol=1e-12, verb=False, x0=None):
        N, cp, ri, v = _csc(A)
        x = np.zeros(N) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64).copy()
        sec = _d(0.0)
        it = self._l.spand_cg(self._h, N, cp, ri, v, np.ascontiguousarray(b, dtype=np.float64), x, iters, tol,
                              int(verb), C.byref(sec))
        if it < 0:
            raise RuntimeError(self._l.spand_last_error(self._h).decode())
        self.t_cg = sec.value
        return it, x

    def gmres(self, A, b, iters=100, restart=100, tol=1e-12, verb=False, x0=None):
        """gmres(A, rhs, x, precond, iters, restart, tol, verb), include/is.h:13 — on the GPU; (iterations, x)."""
        N, cp, ri, v = _csc(A)
        x = np.zeros(N) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64).copy()
        sec = _d(0.0)
        it = self._l.spand_gmres(self._h, N, cp, ri, v, np.ascontiguousarray(b, dtype=np.float64), x, iters, restart,
                                 tol, int(verb), C.byref(sec))
        if it < 0:
            raise RuntimeError(self._l.spand_last_error(self._h).decode())
        self.t_gmres = sec.value
        return it, x

    def nnz(self): return self._l.spand_nnz(self._h)
    def get_stop(self): return self._l.spand_get_stop(self._h)
    def get_N(self): return self._l.spand_get_N(self._h)
    def factorize_seconds(self): return self._l.spand_factorize_seconds(self._h)
    def analyze_seconds(self): return self._l.spand_analyze_seconds(self._h)

    # ---- sub-tree sharding over the GPUs of one node (one process per GPU) ----
    def mg_init(self, dist, device=None, arena_gb=None):
        """Collective over the default torch.distributed group: allocates this rank's shared arena and maps the
        peers' arenas (CUDA IPC). Call before partition/assemble; afterwards every rank makes the same calls."""
        rank, world = dist.get_rank(), dist.get_world_size()
        if device is not None:
            self.set_device(device)
        if arena_gb is None:
            arena_gb = float(os.environ.get("SPAND_MG_ARENA_GB", "40"))
        self._check(self._l.spand_mg_setup(self._h, rank, world, int(arena_gb * (1 << 30))))
        if world == 1:
            return
        buf = C.create_string_buffer(64)
        self._check(self._l.spand_mg_get_handle(self._h, buf))
        handles = [None] * world
        dist.all_gather_object(handles, buf.raw)
        self._check(self._l.spand_mg_set_peers(self._h, b"".join(handles)))
        dist.barrier()

    def mg_owner_map(self, nranks):
        """Rank owning every cluster (order of stats()); host only, valid after partition()."""
        out = np.zeros(self._l.spand_num_clusters(self._h), dtype=np.int32)
        self._check(self._l.spand_mg_owner_map(self._h, nranks, out))
        return out

    # ---- symbolic plan on the host only (no device needed) ----
    def plan_analyze(self, A):
        N, cp, ri, _ = _csc(A)
        self._check(self._l.spand_plan_analyze(self._h, N, cp, ri))

    def plan_live_edges(self, level, phase):
        """(column cluster, row cluster) of the blocks alive after `phase` of `level` (level < 0: as assembled)."""
        n = self._l.spand_plan_live_edges(self._h, level, phase, None, None)
        if n < 0:
            raise RuntimeError(self._l.spand_last_error(self._h).decode())
        a, b = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
        self._l.spand_plan_live_edges(self._h, level, phase, a.ctypes.data, b.ctypes.data)
        return a, b

    PLAN_COUNTS = ("eliminated", "out_panels", "in_panels", "fill_blocks", "schur_targets", "schur_contribs",
                   "scaled_clusters", "scaled_blocks", "rrqr_tasks", "rrqr_wavefronts", "merged_blocks", "merge_copies")

    def plan_counts(self, level):
        out = np.zeros(12, dtype=np.int64)
        self._check(self._l.spand_plan_counts(self._h, level, out))
        return dict(zip(self.PLAN_COUNTS, (int(v) for v in out)))

    def cluster_layout(self):
        """(start, hierarchy level) of every cluster, same order as stats()."""
        n = self._l.spand_num_clusters(self._h)
        a, b = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
        self._l.spand_get_cluster_layout(self._h, a, b)
        return a, b
    def kernel_launches(self): return self._l.spand_kernel_launches(self._h)
    def arena_bytes(self): return self._l.spand_arena_bytes(self._h)

    def set_profile(self, on): self._l.spand_set_profile(self._h, int(on))

    def set_monitor_flops(self, on): self._l.spand_set_monitor_flops(self._h, int(on))

    def flops_log(self):
        """(n, 5) int64 array of (level, kind, rows, cols, inner): reference write_log_flops, src/tree.cpp:60-77."""
        fn = self._l.spand_get_flops_log
        fn.restype = C.c_longlong
        fn.argtypes = [_p, C.c_void_p]
        n = fn(self._h, None)
        if n < 0:
            raise RuntimeError(self._l.spand_last_error(self._h).decode())
        out = np.zeros((n, 5), dtype=np.int64)
        if n:
            fn(self._h, out.ctypes.data)
        return out

    def family_stats(self):
        """{family: (device ms, launches)} of the last factorize(); ms is 0 unless set_profile(True)."""
        nf = self._l.spand_num_families()
        ms = np.zeros(nf)
        ln = np.zeros(nf, dtype=np.int64)
        self._l.spand_get_family_stats(self._h, ms, ln)
        return {self._l.spand_family_name(i).decode(): (float(ms[i]), int(ln[i])) for i in range(nf)}

    def stats(self):
        n = self._l.spand_num_clusters(self._h)
        a = [np.zeros(n, dtype=np.int32) for _ in range(3)]
        self._l.spand_get_stats(self._h, *a)
        return a

    def log(self):
        nf = self._l.spand_log_fields()
        out = np.zeros(self.nlevels * nf)
        self._l.spand_get_log(self._h, out)
        out = out.reshape(self.nlevels, nf)
        return {self._l.spand_log_field_name(i).decode(): out[:, i].copy() for i in range(nf)}

    def get_trailing_mat(self):
        import scipy.sparse as sp
        nnz = self._l.spand_trailing(self._h, None, None, None)
        if nnz < 0:
            raise RuntimeError(self._l.spand_last_error(self._h).decode())
        cp = np.zeros(self.N + 1, dtype=np.int32)
        ri = np.zeros(nnz, dtype=np.int32)
        v = np.zeros(nnz)
        self._l.spand_trailing(self._h, cp.ctypes.data, ri.ctypes.data, v.ctypes.data)
        return sp.csc_matrix((v, ri, cp), shape=(self.N, self.N))
