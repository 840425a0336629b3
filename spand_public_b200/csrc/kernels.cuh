// Device task descriptors + launchers for the spaND factorization hot path on sm_100a.
//
// One launcher = one variable-size batch over all clusters / edges of a level. Which reference
// BLAS/LAPACK call each batch replaces (citations relative to /root/reference):
//   launch_potrf_step   LAPACKE_dpotrf    src/util.cpp:158-165   <- potf_cluster  src/tree.cpp:582
//   launch_trsm_step    cblas_dtrsm       src/util.cpp:256-274   <- trsm_potf_edgeIn/Out src/tree.cpp:640-663
//   launch_gemm         cblas_dgemm/dsyrk src/util.cpp:122-156   <- gemm_edges    src/tree.cpp:748-793
//   launch_rrqr         assemble_Asn + LAPACKE_dgeqp3 + choose_rank + scatter
//                                         src/tree.cpp:1189-1224, 1292-1347, 1004-1046; src/util.cpp:383-452
//   launch_copy         update_edges block copy + setZero           src/tree.cpp:1133-1184
//   launch_trsv / launch_gemv / launch_house / launch_xcopy
//                       Scaling*/Gemm*/Orthogonal/Merge ::fwd/bwd   src/operations.cpp:31-208
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "symbolic.hpp"

namespace spand {

constexpr int NB = 64;  // block size of the right-looking blocked POTRF / TRSM

struct PotrfTask {
    double* A;  // n x n, column-major, lower triangle referenced
    int ld, n;
};

enum TrsmMode {
    TRSM_RLT = 0,  // B (m x n) <- B T^-T, T lower n x n      (out-edge of an LLT pivot)
    TRSM_LLN = 1,  // B (n x m) <- T^-1 B, T lower n x n      (in-edge of an LLT / PLU pivot)
    TRSM_RUN = 2,  // B (m x n) <- B T^-1, T upper n x n      (out-edge of a PLU pivot)
    TRSM_LLU = 3,  // B (n x m) <- T^-1 B, T unit lower       (U12 inside the blocked GETRF)
};

struct TrsmTask {
    double* B;
    const double* T;
    int ldb, ldt;
    int m;  // free dimension of B
    int n;  // triangle dimension
    const double* diag;  // nullptr: diagonal of T; else the n diagonal entries (PLU keeps diag(U) beside the block)
    const double* inv;   // strip kernels only: inverses of the 64 x 64 diagonal blocks of T, 4096 doubles each
    int tri;             // strip kernel, TRSM_LLN only: B is block lower triangular (rows above the strip's first column
                         // block are zero and stay zero), e.g. B = I when the whole inverse of T is wanted
};

struct TrtriTask {  // inverses of the diagonal blocks of one triangle
    const double* T;
    int ldt, n;
    double* inv;
    int ldw;  // 0: 64 x 64 blocks of 4096 doubles each; > 0 (n <= 64 only): one n x n block with this leading dimension
};

struct UtTask {  // Lt (n x n, leading dimension n) <- U^T as a lower triangle: Lt(i, p) = U(p, i) for p < i, Lt(i, i) = ud[i]
    const double* U;
    const double* ud;
    double* Lt;
    int ldu, n;
};

struct EyeTask {  // W (n x n, leading dimension ld) <- I
    double* W;
    int n, ld;
};

// PLU pivot (src/util.cpp:183-227): on exit A holds L (lower, with its non-unit diagonal |d|^1/2) and the strictly
// upper part of U; ud = diag(U) = sign(d) |d|^1/2; ipiv = LAPACK swap sequence (0-based); perm = swap2perm(ipiv).
struct GetrfTask {
    double* A;
    int ld, n;
    double* ud;
    int* ipiv;
    int* perm;
};

struct RowPermTask {  // B (n x m) <- B[perm, :]
    double* B;
    int ldb, n, m;
    const int* perm;
};

struct GemmContrib {
    const double* A;  // m x k, column-major
    const double* B;  // NT: n x k (used transposed);  NN: k x n
    int lda, ldb, k;
};

// POS: C = (0|C) + sum. TRIB (NT form): B_c is lower triangular (n x n), the inner index of a tile stops at its last
// column; TRIA (NN form): A_c is lower triangular (m x m), the inner index of a tile stops at its last row.
enum GemmFlags { GEMM_LOWER = 1, GEMM_ZERO_INIT = 2, GEMM_NN = 4, GEMM_POS = 8, GEMM_TRIB = 16, GEMM_TRIA = 32 };

struct GemmTask {
    double* C;  // m x n, C = (ZERO_INIT ? 0 : C) - sum_c A_c op(B_c)
    int ldc, m, n;
    int c0, nc;  // contributions [c0, c0 + nc), applied in this order (deterministic)
    int flags;
};

struct QrSrc {
    double* blk;     // the edge block
    int ld;
    int nbr;         // neighbour cluster id: its current size is the free dimension of the block
    int transposed;  // 0: in-edge  A[s,n] (rows x w);  1: out-edge A[n,s] (w x rows)
};

constexpr int QR_NB = 16;   // block size of the blocked QRCP (dlaqps-like)
#ifndef SPAND_QR_NBS
#define SPAND_QR_NBS 8
#endif
constexpr int QR_NBS = SPAND_QR_NBS;   // block size of the streaming shape (panel in global memory, 64 registers per thread)
constexpr int QR_FLD = 17;  // row stride of the F block in shared memory (odd: conflict-free)

struct QrTask {
    int cluster;
    int rows;     // current size of the cluster (it has not been sparsified before)
    int src0, nsrc;
    double* W;    // scratch panel ld x maxcols in global memory (L2-resident), only when !in_smem
    int maxcols;  // upper bound of the gathered column count (sizes used for the launch layout)
    double* V;    // out: rows x rank Householder vectors (unit diagonal implicit), ld = rows
    double* tau;  // out: rank
    int ld;       // leading dimension of the panel (shared or global)
    int L;        // lanes per column (power of two <= 32)
    int nb;       // block size <= QR_NB
    int in_smem;  // panel resident in (distributed) shared memory
    int hcap;     // hot-set kernel (rrqr_hc2.cu): capacity of the shared-memory hot set, in columns
    double* X;    // hot-set kernel, G >= 8: exchange area of hc2_exchange_doubles(rows, G) doubles
};

struct CopyTask {
    const double* src;  // nullptr => identity block
    double* dst;
    int lds, ldd, rows, cols;
};

struct TrsvTask {
    const double* T;
    double* x;
    int ld, n;
    const double* diag;  // upper solve only: diag(U) kept beside the block (nullptr: diagonal of T)
    const int* perm;     // forward solve only: x <- x[perm] first (P^T of PLU), nullptr: none
};

struct GemvContrib {
    const double* A;
    const double* x;
    int lda, k;  // fwd: A is m x k, y -= A x ;  bwd (trans): A is k x m, y -= A^T x
};

struct GemvTask {
    double* y;
    int m, c0, nc;
};

struct HouseTask {
    const double* V;
    const double* tau;
    double* x;
    int rows, rank;
};

struct XCopyTask {
    const double* src;
    double* dst;
    int n;
};

// All launchers are asynchronous on `st`. `err` is a device int: bit 0 = non-SPD pivot.
void launch_potrf_step(const PotrfTask* t, int nt, int j0, int* err, cudaStream_t st);
void launch_trsm_step(int mode, const TrsmTask* t, int nt, int j0, int max_m, cudaStream_t st);
// Whole triangular solves as DMMA GEMMs (TRSM_RLT / TRSM_LLN): one CTA per 64-wide strip of the free dimension sweeps
// the 64-wide blocks of the triangle, X_j = (B_j - sum_{p<j} X_p T_jp^T) inv(T_jj)^T. strip_prefix: nt + 1 exclusive
// prefix of ceil(m / 64).
void launch_trtri(const TrtriTask* t, int nt, int max_n, cudaStream_t st);
void launch_eye(const EyeTask* t, int nt, cudaStream_t st);
void launch_ut(const UtTask* t, int nt, cudaStream_t st);
void launch_trsm_strip(int mode, const TrsmTask* t, int nt, const int* strip_prefix, int total_strips, cudaStream_t st);
// GETRF with partial pivoting. n <= 64: one launch does everything (factor, split_LU, perm). Larger: right-looking
// over 64-wide panels: launch_getrf_panel (pivoting inside the panel) -> launch_getrf_laswp (row swaps outside the
// panel) -> TRSM_LLU + GEMM from the host driver -> launch_getrf_finish (split_LU + swap2perm) at the end.
// err bit 1 = exactly singular pivot.
void launch_getrf_small(const GetrfTask* t, int nt, int* err, cudaStream_t st);
void launch_getrf_panel(const GetrfTask* t, int nt, int j0, int* err, cudaStream_t st);
void launch_getrf_laswp(const GetrfTask* t, int nt, int j0, int max_n, cudaStream_t st);
void launch_getrf_finish(const GetrfTask* t, int nt, cudaStream_t st);
void launch_rowperm(const RowPermTask* t, int nt, cudaStream_t st);
// tile_prefix: nt + 1 exclusive prefix of ceil(m/64)*ceil(n/64); total_tiles = tile_prefix[nt]; tile_task (optional):
// task of every tile (replaces the binary search in tile_prefix when a launch has 10^5 tasks)
void launch_gemm_tiled(const GemmTask* t, int nt, const GemmContrib* c, const int* tile_prefix, int total_tiles,
                       cudaStream_t st, const int* tile_task = nullptr);
void launch_gemm_small(const GemmTask* t, int nt, const GemmContrib* c, cudaStream_t st);
// csize: device array of current cluster sizes (read for neighbours, written with the rank)
// One thread-block cluster of G CTAs per task. Launch shapes: (128 threads, G = 1, panel in shared memory),
// (256 threads, G = 1..16, panel in distributed shared memory), (512 threads, G = 8, panel in global scratch).
// smem = dynamic shared memory per CTA >= rrqr_smem_bytes(...) of every task of the launch.
// theta > 0 (global panels only): hot / cold kernel, columns below theta x the pivot's norm at the start of a block
// are refreshed once per block on the tensor cores instead of being swept every step; theta = 0: every column swept.
void launch_rrqr(const QrTask* t, int nt, const QrSrc* s, int* csize, double tol, int G, int nthreads, bool in_smem,
                 int smem, cudaStream_t st, double theta = 0.0);
// Hot-set kernel (rrqr_hc2.cu): pivot search on a shared-memory copy of the columns that can win it, cold columns
// refreshed once per block on the tensor cores. Panel in global memory (t.W), rows <= 640.
constexpr int HC2_NB = 16;                 // largest block (reflectors between two refreshes)
int hc2_row_pairs(int rows);               // template selector, 0: too tall for this kernel
int hc2_threads(int rows);
size_t hc2_smem_bytes(int rows, int maxcols, int G, int hcap, int nsrc);
size_t hc2_exchange_doubles(int rows, int G);
void launch_rrqr_hc2(const QrTask* t, int nt, const QrSrc* s, int* csize, double tol, int G, int row_pairs, int smem,
                     double theta, cudaStream_t st);
void hc2_stats(unsigned long long* out16, bool reset);  // -DSPAND_RRQR_TIMING builds
// Column kernel (rrqr_hc2.cu): panels of at most 64 rows resident in the shared memory of one CTA, one thread per
// column (no reductions over the short rows). t.ld = rrqr_col_ld(rows).
int rrqr_col_ld(int rows);
int rrqr_col_max_rows();
size_t rrqr_col_smem_bytes(int rows, int maxcols, int nsrc, bool global_panel);
void launch_rrqr_col(const QrTask* t, int nt, const QrSrc* s, int* csize, double tol, int nthreads, int smem,
                     bool global_panel, cudaStream_t st);
int rrqr_max_smem();
// one dense matrix through the batch kernels (kernel-level tests); see rrqr.cu
int rrqr_single(int rows, int cols, const double* A_host, int nsrc, int transposed, double tol, int G, int nthreads,
                int in_smem, int nb, double theta, int* rank_out, double* R_host, double* V_host, double* tau_host,
                std::string& err);
// debug builds (-DSPAND_RRQR_TIMING): accumulated clock64 cycles per kernel phase, [10] = CTAs counted
void rrqr_phase_cycles(unsigned long long* out48, bool reset);
size_t rrqr_smem_bytes(int rows, int maxcols, int G, int nb, int ld, bool in_smem);
void launch_copy(const CopyTask* t, int nt, cudaStream_t st);
// max_n / max_m: upper bound of the task sizes of the launch (selects the kernel shapes and the grid)
void launch_trsv(const TrsvTask* t, int nt, int trans, int max_n, cudaStream_t st);
void launch_gemv(const GemvTask* t, int nt, const GemvContrib* c, int trans, int max_m, cudaStream_t st);
void launch_house(const HouseTask* t, int nt, int trans, int max_rows, cudaStream_t st);
void launch_xcopy(const XCopyTask* t, int nt, cudaStream_t st);
void launch_fill(double* p, size_t n, double v, cudaStream_t st);

// ------------------------------------------------------------------------------------------------
// Plan-driven ("sym") batches: the task arrays of the symbolic plan (symbolic.hpp) are resident on the device and
// hold ids only; the kernels resolve them through these tables at run time. They serve every task whose
// dimensions are all <= 64 (one warp for <= 32, one 64-thread CTA through a device-side "mid" list otherwise) and
// skip the rest, which the host drives through the pointer-based blocked launchers above.
// ------------------------------------------------------------------------------------------------
struct DevTables {
    const int* csize;       // current size of every cluster
    double* const* eptr;    // block of every edge
    const int* eld;         // its leading dimension
    const int* en1;         // column cluster of every edge
    const int* en2;         // row cluster of every edge
    double* const* xptr;    // solution segment of every cluster
    const int* pos;         // offset of a cluster inside its parent (valid after the merge planning of its level)
    const int* parent;
    // sub-tree sharding over several GPUs: rank owning every cluster (a block belongs to the owner of its column
    // cluster, a task to the owner of the block / segment it writes); rank < 0: single GPU, nothing is filtered
    const int* owner;
    int rank;
};
constexpr int SMALL_DIM = 64;      // largest dimension served by the plan-driven kernels
constexpr int COPY_SMALL = 4096;   // largest block (elements) copied by one warp in the merge

// mid / cnt: device work list (capacity nt) and its counter (zeroed by the caller) used inside the launcher
void launch_potrf_sym(const DevTables& T, const int* clusters, const int* piv, int nt, int* mid, int* cnt, int* err,
                      cudaStream_t st);
void launch_trsm_sym(int mode, const DevTables& T, const SymTrsm* tasks, int nt, int* mid, int* cnt, cudaStream_t st);
// two-sided scaling of a block: B <- L_row^-1 (B L_col^-T); right[i] / left[i] describe the same block.
// warp_only: blocks with a dimension above 32 are left to the caller (explicit-inverse GEMM path of the host driver)
void launch_scale_sym(const DevTables& T, const SymTrsm* right, const SymTrsm* left, int nt, int* mid, int* cnt,
                      cudaStream_t st, bool warp_only = false);
void launch_gemm_sym(const DevTables& T, const SymGemm* tasks, int nt, const SymCon* con, int* mid, int* cnt,
                     cudaStream_t st);
// GEN / PLU blocks with both dimensions <= 64 (src/tree.cpp:668-689, :838-853), plan-driven like the LLT batches.
// ud[c] / pperm[c]: diag(U) and the row permutation of the current pivot of cluster c (device pointer tables).
//   PLU_BOTH  (scale):           B <- L_row^-1 P_row^T (B U_col^-1); right[i] / left[i] describe the same block
//   PLU_RIGHT (out-edge panels): B <- B U^-1                        (tasks in `right`, `left` unused)
//   PLU_LEFT  (in-edge panels):  B <- L^-1 P^T B                    (tasks in `left`, `right` unused)
enum PluMode { PLU_BOTH = 0, PLU_RIGHT = 1, PLU_LEFT = 2 };
void launch_plu_sym(int mode, const DevTables& T, const SymTrsm* right, const SymTrsm* left, int nt,
                    const double* const* ud, const int* const* pperm, int* mid, int* cnt, cudaStream_t st);
void launch_copy_sym(const DevTables& T, const SymCopy* tasks, int nt, int pivots_identity, cudaStream_t st);
void launch_expand_qsrc(const DevTables& T, const SymQrSrc* s, int n, QrSrc* out, cudaStream_t st);
// recorded operations -> solve batches (sizes are captured now)
void launch_expand_trsv(const DevTables& T, const int* clusters, const int* piv, int nt, TrsvTask* out, cudaStream_t st);
void launch_expand_gemv(const DevTables& T, const SymGemv* t, int nt, const SymGemvCon* c, int ncon, GemvTask* out,
                        GemvContrib* outc, cudaStream_t st);
void launch_expand_xcopy(const DevTables& T, const int* children, int n, XCopyTask* fwd, XCopyTask* bwd, cudaStream_t st);
void launch_expand_house(const DevTables& T, const QrTask* q, int n, HouseTask* out, cudaStream_t st);
// Multi-GPU plumbing over peer-mapped memory (NVLink): `flags[r]` is rank r's array of nranks epoch counters.
// launch_peer_barrier: every rank stores `epoch` into its slot on every peer, then waits until all of its own slots
// have reached `epoch` (system-scope fences on both sides): all kernels enqueued before it on every rank are complete
// and visible when the kernels enqueued after it start.
constexpr int MG_MAX_RANKS = 16;
struct PeerPtrs {
    void* p[MG_MAX_RANKS];
};
void launch_peer_barrier(const PeerPtrs& flags, int rank, int nranks, unsigned epoch, cudaStream_t st);
// Error flag of a sharded factorization: every rank publishes its flag word in its shared arena, then ORs the words
// of all ranks into its own flag (between two peer barriers), so that all ranks report the same error together.
void launch_err_publish(const int* err, int* shared_word, cudaStream_t st);
void launch_err_or(const PeerPtrs& words, int nranks, int* err, cudaStream_t st);
// csize[first, first + n) <- min over ranks (sizes only shrink: the replica that ran the RRQR holds the rank)
void launch_csize_min(const PeerPtrs& csize, int rank, int nranks, int first, int n, cudaStream_t st);
// x[idx[i]] = leaf_r[i] with r = dof_owner[i]: collects the solution segments from their owners
void launch_scatter_owned(int n, const int* idx, const PeerPtrs& leaf, const signed char* dof_owner, double* dst,
                          cudaStream_t st);
// assembly: dst[map[k]] = val[k] for map[k] != 0xffffffff
// (dst0: offset subtracted from the map, n: number of values; entries outside [0, dst_len) are skipped)
void launch_scatter_values(const double* val, const unsigned* map, size_t n, double* dst, cudaStream_t st);

// PCG building blocks (src/is.cpp:39-121)
void launch_spmv(int n, const int* rowptr, const int* colind, const double* val, const double* x, double* y,
                 cudaStream_t st);
void launch_axpy(int n, double alpha, const double* x, double* y, cudaStream_t st);      // y += alpha x
void launch_xpay(int n, const double* x, double beta, double* y, cudaStream_t st);       // y = x + beta y
void launch_gather(int n, const int* idx, const double* src, double* dst, cudaStream_t st);   // dst[i] = src[idx[i]]
void launch_scatter(int n, const int* idx, const double* src, double* dst, cudaStream_t st);  // dst[idx[i]] = src[i]

// Householder GMRES building blocks (src/is.cpp:123-300), krylov.cu. `partial` is a scratch of KR_GRID + 1 doubles;
// tau / beta are device scalars. Two launches per call, deterministic reductions, no host synchronisation.
constexpr int KR_GRID = 148 * 4;
void launch_kr_house_apply(int n, double* v, const double* ess, const double* tau, double* partial, cudaStream_t st);
void launch_kr_house_make(int n, const double* x, double* ess, double* tau, double* beta, double* partial,
                          cudaStream_t st);
// out[0] = a . b, reduced in a fixed order (bit-reproducible, unlike launch_dot's atomics)
void launch_dot_det(int n, const double* a, const double* b, double* partial, double* out, cudaStream_t st);
void launch_kr_unit(int n, int k, double* v, cudaStream_t st);            // v = e_k
void launch_kr_residual(int n, const double* b, double* y, cudaStream_t st);  // y = b - y

}  // namespace spand
