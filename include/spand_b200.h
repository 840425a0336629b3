/* spand_b200.h — C ABI of the B200-native spaND factorization path.
 *
 * This is the drop-in boundary: every entry point replaces one public member of the reference's
 * spaND::Tree (include/tree.h, citations relative to /root/reference) or one free function taking a
 * Tree (include/is.h). Plain pointers and sizes only; all matrices are CSC with int32 indices and
 * FP64 values; vectors are host pointers unless the name says _device. Return value 0 = success,
 * 1 = "Error: Non-SPD Pivot" (src/tree.cpp:587-590), 2 = "Error: Singular Pivot" (src/tree.cpp:625-628),
 * -1 = any other failure (message via spand_last_error). There is no CPU fallback: calls that need the
 * GPU fail with -1 when no CUDA device is present.
 */
#ifndef SPAND_B200_H
#define SPAND_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct spand_tree spand_tree;

/* enums of include/spaND.h:19-21 */
enum { SPAND_SPD = 0, SPAND_SYM = 1, SPAND_GEN = 2 };
enum { SPAND_LLT = 0, SPAND_PLU = 3 };

/* Tree::Tree(int lvl)                               include/tree.h:170, src/tree.cpp:86-113 */
spand_tree* spand_create(int nlevels);
void spand_destroy(spand_tree* t);
const char* spand_last_error(spand_tree* t);

/* Tree::set_tol / set_skip / set_symm_kind / set_scaling_kind / set_use_geo / set_verb / set_use_sparsify
 *                                                   include/tree.h:133-147, src/tree.cpp:114-131 */
int spand_set_tol(spand_tree* t, double tol);
int spand_set_skip(spand_tree* t, int skip);
int spand_set_symm_kind(spand_tree* t, int kind);
int spand_set_scaling_kind(spand_tree* t, int kind);
int spand_set_use_geo(spand_tree* t, int geo);
int spand_set_verb(spand_tree* t, int verb);
int spand_set_use_sparsify(spand_tree* t, int use);
int spand_set_device(spand_tree* t, int device);
/* Tree::set_Xcoo — dim x N, column-major; copied (the reference borrows the pointer, tree.h:56) */
int spand_set_coords(spand_tree* t, int dim, int N, const double* X);
/* test hook: stop after (level, phase) with phase 0=eliminate 1=scale 2=sparsify 3=merge; -1 = off */
int spand_set_stop(spand_tree* t, int level, int phase);

/* Tree::partition(SpMat&)                           include/tree.h:176, src/tree.cpp:306-419
 * A: symmetric pattern (values ignored). */
int spand_partition(spand_tree* t, int N, const int* colptr, const int* rowind);
/* the returned vector<ClusterID>, natural ordering; each array has N entries */
int spand_get_partition(spand_tree* t, int* self_lvl, int* self_sep, int* l_lvl, int* l_sep, int* r_lvl, int* r_sep);
/* Tree::get_assembly_perm                           include/tree.h:152 */
int spand_get_perm(spand_tree* t, int* perm);
int spand_get_N(spand_tree* t);

/* Tree::assemble(SpMat&)                            include/tree.h:182, src/tree.cpp:505-575
 * Builds the dense blocks and uploads them to HBM. May be called again to restart from the same partition. */
int spand_assemble(spand_tree* t, int N, const int* colptr, const int* rowind, const double* val);

/* Tree::factorize()                                 include/tree.h:186, src/tree.cpp:1447-1551 */
int spand_factorize(spand_tree* t);

/* Tree::solve(VectorXd&) — in place                 include/tree.h:192, src/tree.cpp:1610-1635 */
int spand_solve(spand_tree* t, double* x);
int spand_solve_device(spand_tree* t, double* x_device);

/* cg(A, rhs, x, precond, iters, tol, verb)          include/is.h:12, src/is.cpp:39-121
 * returns the iteration count of the reference (i+1) or a negative error; x is the initial guess / result */
int spand_cg(spand_tree* t, int N, const int* colptr, const int* rowind, const double* val, const double* rhs,
             double* x, int iters, double tol, int verb, double* seconds);

/* gmres(A, rhs, x, precond, iters, restart, tol, verb)   include/is.h:13, src/is.cpp:123-300
 * Householder GMRES, left preconditioned; returns the reference's iteration count or a negative error */
int spand_gmres(spand_tree* t, int N, const int* colptr, const int* rowind, const double* val, const double* rhs,
                double* x, int iters, int restart, double tol, int verb, double* seconds);

/* geqp3(A, jpvt, tau) + choose_rank(diag(R), tol) + triu(R[:rank, :]) P^T on ONE dense matrix through the batched
 * sparsification kernel                              src/util.cpp:383-394, :434-452, src/tree.cpp:1334-1335
 * A: rows x cols column-major, given as nsrc column blocks of equal width (transposed != 0: every block is stored
 * transposed, as the out-edges of a cluster are, src/tree.cpp:1211-1217). On return *rank is the truncated rank and, when
 * *rank < rows, the first *rank rows of every block of R (same layout as A) hold triu(R[:rank, :]) P^T; V (rows x
 * min(rows, cols), unit diagonal implicit) and tau hold the Householder reflectors. G / nthreads / in_smem / nb select
 * the launch shape (cluster width, CTA size, panel in shared (1) or global (0) memory, block size); theta > 0 selects the
 * hot / cold variant for global panels; in_smem = 2 selects the hot-set kernel (nb = capacity of its shared-memory hot
 * set, in columns); in_smem = 3 selects the column kernel (rows <= 64, one thread per column, nthreads = 128 or 256).
 * Kernel-level parity tests drive every shape through this entry. */
int spand_geqp3_truncated(int rows, int cols, const double* A, int nsrc, int transposed, double tol, int G,
                          int nthreads, int in_smem, int nb, double theta, int* rank, double* R, double* V, double* tau);

/* profiling aid: per-phase clock64 cycles of the RRQR kernel, non-zero only in -DSPAND_RRQR_TIMING builds */
void spand_debug_rrqr_phases(unsigned long long* out48, int reset);
/* same for the hot-set kernel: [0] steps [1] blocks [2] blocks closed early [3] hot columns [4] unpivoted columns
 * (both summed over the block starts) [5] threshold retries [6..9] clock64 cycles of the first CTA of every task:
 * block boundary / hot loop / block end / gather + scatter */
void spand_debug_hc2_stats(unsigned long long* out16, int reset);

/* Tree::nnz / get_stop / get_nlevels                include/tree.h:158-160 */
long long spand_nnz(spand_tree* t);
int spand_get_stop(spand_tree* t);
int spand_get_nlevels(spand_tree* t);

/* write_stats triple (order id, original size, final size) for every hierarchy cluster
 *                                                   include/tree.h:242-251, src/tree.cpp:44-57 */
int spand_num_clusters(spand_tree* t);
int spand_get_stats(spand_tree* t, int* id, int* size, int* rank);
/* first row/column (permuted ordering) and hierarchy level of every cluster, same order as spand_get_stats
 * (Cluster::get_start(), include/cluster.h:16-116) */
int spand_get_cluster_layout(spand_tree* t, int* start, int* hlevel);

/* Tree::log / Tree::tprof                           include/tree.h:163-165, include/util.h:262-401
 * spand_log_fields() doubles per level, see spand_log_field_name(i) */
int spand_log_fields(void);
const char* spand_log_field_name(int i);
int spand_get_log(spand_tree* t, double* out);
/* device time of the last factorize(), CUDA events on the factorization stream, seconds */
double spand_factorize_seconds(spand_tree* t);
/* host seconds of the symbolic analysis done by the first spand_assemble() of a (partition, pattern) pair; later
 * spand_assemble() calls with the same pattern reuse the plan (no counterpart in the reference, whose list
 * manipulations src/tree.cpp:748-793, :1133-1184 are replayed inside every factorize) */
double spand_analyze_seconds(spand_tree* t);
/* The symbolic plan on the host only (no device needed): replays the list manipulations of Tree::factorize
 * (src/tree.cpp:748-793 gemm_edges fill-in, :1133-1184 update_edges) on integers. spand_plan_live_edges returns the
 * blocks alive after `phase` (0 eliminate, 1 scale, 2 sparsify, 3 merge) of `level` (< 0: as assembled) as
 * (column cluster, row cluster) ids; pass NULL to get the count. spand_plan_counts fills 12 numbers per level:
 * eliminated clusters, out-panels, in-panels, fill-in blocks, Schur targets, Schur contributions, scaled clusters,
 * scaled blocks, RRQR tasks, RRQR wavefronts, merged blocks, merge copies. */
int spand_plan_analyze(spand_tree* t, int N, const int* colptr, const int* rowind);

/* Sub-tree sharding over the GPUs of one node, one process per GPU (nothing of this exists in the reference, which is
 * a single sequential process). Every rank calls, in this order: spand_set_device, spand_mg_setup
 * (allocates the rank's shared arena: blocks, solution segments and cluster sizes that peers reach through NVLink),
 * spand_mg_get_handle (64 bytes, a cudaIpcMemHandle_t), an all-gather of the handles in the caller's process group,
 * spand_mg_set_peers (nranks * 64 bytes in rank order), then the usual partition / assemble / factorize / solve with
 * identical arguments on every rank. Every rank ends up with the full solution. nranks must be a power of two.
 * spand_mg_owner_map is host-only: the rank owning every cluster (order of spand_get_stats). SPD/LLT only. */
int spand_mg_setup(spand_tree* t, int rank, int nranks, long long arena_bytes);
int spand_mg_get_handle(spand_tree* t, void* out64);
int spand_mg_set_peers(spand_tree* t, const void* handles);
int spand_mg_owner_map(spand_tree* t, int nranks, int* owner);
int spand_plan_live_edges(spand_tree* t, int level, int phase, int* n1, int* n2);
int spand_plan_counts(spand_tree* t, int level, long long* out);
long long spand_kernel_launches(spand_tree* t);
long long spand_arena_bytes(spand_tree* t);

/* Tree::set_monitor_flops + write_log_flops       include/tree.h:147, src/tree.cpp:60-77, pushes at :592,648,662,792,1312
 * With monitoring on, spand_get_flops_log returns one tuple (level, kind, rows, cols, inner) per BLAS call the
 * reference would have made: kind 0 pivot (rows), 1 panel (free dimension, triangle dimension), 2 gemm (target rows,
 * target cols, inner), 3 rrqr (rows, cols). The reference's sixth column (seconds of that call) has no counterpart:
 * the calls of a level run as batches; per-family device time is spand_get_family_stats. Two-call protocol: out = NULL
 * returns the number of tuples, else 5 long long per tuple are written. Call after spand_factorize. */
int spand_set_monitor_flops(spand_tree* t, int on);
long long spand_get_flops_log(spand_tree* t, long long* out);

/* Per-kernel-family device time of the last factorize(): CUDA events recorded around every launch of the
 * family on the factorization stream (the role of the reference's per-routine Profile counters potf/trsm/
 * gemm/geqp3/merge_copy, include/util.h:262-309). Enable with spand_set_profile(t, 1) before factorize(). */
int spand_set_profile(spand_tree* t, int on);
int spand_num_families(void);
const char* spand_family_name(int f);
int spand_get_family_stats(spand_tree* t, double* ms, long long* launches);

/* Tree::get_trailing_mat                            include/tree.h:153, src/tree.cpp:1730-1763
 * permuted ordering, CSC; call with null pointers to get nnz first */
int spand_trailing(spand_tree* t, int* colptr, int* rowind, double* val);

/* host utilities of src/util.cpp used by the drivers */
void spand_util_random(int size, int seed, double* out);              /* util.cpp:549-558 */
void spand_util_linspace_nd(int n, int dim, double* out);             /* util.cpp:488-517 */
int spand_util_neglapl(int n, int d, int* colptr, int* rowind, double* val); /* mats/neglapl_d_n.mm generator */
int spand_util_aniso(int n, int* colptr, int* rowind, double* val);  /* config C5, SURVEY.md 8(d) */
int spand_util_mm_read(const char* fn, int* rows, int* cols, int* colptr, int* rowind, double* val); /* mmio.hpp:160 */
int spand_util_mm_read_dense(const char* fn, int* rows, int* cols, double* out);                      /* mmio.hpp:225 */

#ifdef __cplusplus
}
#endif
#endif
