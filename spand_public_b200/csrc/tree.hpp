// spand::Tree — B200-native counterpart of spaND::Tree (reference include/tree.h:130-198).
//
// Same call contract: setters -> partition -> assemble -> factorize -> solve*. The block-sparse
// trailing matrix (reference Cluster/Edge graph, include/cluster.h:16-116, include/edge.h:32-44) is
// held as flat host arrays describing structure only; every dense block lives in a device arena
// (HBM). factorize() walks the levels like src/tree.cpp:1447-1551 but emits one variable-size batch
// per phase instead of one BLAS call per block. The recorded operations (include/operations.h) are
// kept as per-level batch descriptors on the device so that solve() replays them with a fixed
// sequence of launches.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "host/partition.hpp"
#include "host/sparse.hpp"
#include "kernels.cuh"

namespace spand {

enum SymmKind { SPD = 0, SYM = 1, GEN = 2 };
enum ScalingKind { SVD = 1, PLU = 3, PLUQ = 4, EVD = 2, LLT = 0, LDLT = 5 };

struct LevelLog {
    int dofs_nd = 0, dofs_left_nd = 0, dofs_left_elim = 0, dofs_left_spars = 0;
    long long fact_nnz = 0;
    long long rank_before = 0, rank_after = 0;
    int nspars = 0, ignored = 0;
    long long nbrs = 0;
    // device time per phase (CUDA events on the factorization stream), seconds
    double t_elim = 0, t_scale = 0, t_spars = 0, t_merge = 0;
    // flop / byte model of SURVEY.md section 8(d)
    double fl_pivot = 0, fl_panel = 0, fl_schur = 0, fl_rrqr_rank = 0, fl_rrqr_full = 0;
    double by_scale = 0, by_rrqr = 0, by_merge = 0;
    // host planning time and launch counts
    double t_host = 0;
    double t_plan_elim = 0, t_plan_scale = 0, t_plan_spars = 0, t_plan_merge = 0;  // host planning only (no waits)
    int launches = 0;
    int wavefronts = 0;
};

// Bump allocator over cudaMalloc'ed chunks. Blocks never move; nothing is freed before reset().
class DeviceArena {
   public:
    explicit DeviceArena(size_t chunk_bytes) : chunk_(chunk_bytes) {}
    ~DeviceArena() { release(); }
    void* alloc(size_t bytes);
    template <class T>
    T* alloc_n(size_t n) { return (T*)alloc(n * sizeof(T)); }
    void reset();    // keep chunks, rewind
    void release();  // cudaFree everything
    size_t used() const { return used_; }
    size_t capacity() const;

   private:
    struct Chunk { char* p; size_t cap, top; };
    std::vector<Chunk> chunks_;
    size_t cur_ = 0, chunk_, used_ = 0;
};

// Pinned host staging + device mirror for descriptor arrays.
class Stager {
   public:
    ~Stager();
    void reserve(size_t bytes);
    void reset() { top_ = 0; }
    // copies `bytes` from src into pinned staging and enqueues H2D into dst
    void upload(void* dst, const void* src, size_t bytes, cudaStream_t st);

   private:
    char* pinned_ = nullptr;
    size_t cap_ = 0, top_ = 0;
};

class Tree {
   public:
    explicit Tree(int nlevels);
    ~Tree();

    // include/tree.h:133-147
    bool verb = false;
    bool use_geo = false;
    double tol = 10.0;
    int skip = 0;
    int symm_kind = SPD;
    int scale_kind = LLT;
    bool use_want_sparsify = true;
    bool monitor_flops = false;
    // per-call flop tuples (level, kind 0 pivot / 1 panel / 2 gemm / 3 rrqr, rows, cols, inner): what the reference
    // pushes at every BLAS call when set_monitor_flops is on (tree.cpp:592,648,662,792,1312, written by
    // write_log_flops tree.cpp:60-77); here they are listed from the plan after the factorization (finalize_logs)
    struct FlopTuple { long long lvl, kind, rows, cols, inner; };
    const std::vector<FlopTuple>& flop_log() { finalize_logs(); return flop_log_; }
    int stop_level = -1, stop_phase = -1;  // parity-test hook (see oracle)
    int device = 0;
    bool device_in_use() const { return st_ != nullptr; }
    // per-kernel-family device timing (CUDA events around every launch of the family on the factorization
    // stream); off by default because the extra events serialise nothing but cost host time
    bool profile_families = false;
    enum Family { F_POTRF = 0, F_TRSM, F_GEMM, F_RRQR, F_COPY, F_COUNT };
    double family_ms[F_COUNT] = {0, 0, 0, 0, 0};
    long long family_launches[F_COUNT] = {0, 0, 0, 0, 0};
    static const char* family_name(int f);

    void set_coords(int dim, int N, const double* X);
    void partition(const SpMat& A);
    void assemble(const SpMat& A);
    void assemble_csc(int n, const int* colptr, const int* rowind, const double* val);  // borrowed arrays, no host copy
    void factorize();                 // throws std::runtime_error("Error: Non-SPD Pivot\n") etc.
    void solve(double* x_host);       // in place, host vector of length N
    void solve_device(double* x_dev); // in place, device vector of length N (natural ordering)
    int cg(const SpMat& A, const double* rhs, double* x, int iters, double tol, bool verb, double* seconds);
    // src/is.cpp:123-300 (Householder GMRES, left preconditioned by this tree); same return value as the reference
    int gmres(const SpMat& A, const double* rhs, double* x, int iters, int restart, double tol, bool verb,
              double* seconds);
    long long nnz();
    int get_stop() const;
    SpMat trailing_mat();
    void stats(std::vector<int>& id, std::vector<int>& size, std::vector<int>& rank) const;
    void cluster_layout(std::vector<int>& start, std::vector<int>& hlevel) const;  // same order as stats()

    int nlevels;
    int N = 0;
    Ordering ord;
    std::vector<LevelLog> log;   // call logs() for the completed flop / byte / nnz model
    const std::vector<LevelLog>& logs();
    double analyze_seconds() const { return t_analyze_; }
    // ---- sub-tree sharding over the GPUs of one node, one process per GPU ----
    // mg_setup allocates this rank's shared arena (blocks, solution segments, cluster sizes: everything a peer may
    // touch) and must be followed by an exchange of the IPC handles (mg_get_handle / mg_set_peers) before assemble().
    int mg_rank = 0, mg_nranks = 1;
    void mg_setup(int rank, int nranks, size_t arena_bytes);
    void mg_get_handle(void* out64) const;
    void mg_set_peers(const void* handles);
    std::vector<int> owner_map(int nranks) const;  // host only: rank owning every cluster (valid after partition)
    // symbolic plan without a device (CPU tests of the planner)
    void analyze_only(const SpMat& A);
    void plan_live_edges(int level, int phase, std::vector<int>& n1, std::vector<int>& n2) const;
    void plan_counts(int level, long long out[12]) const;
    double t_factorize_device = 0;  // seconds, CUDA events
    // device bytes holding blocks and factors: the private arena plus, when sharded, this rank's share of the
    // peer-mapped arena (where the blocks it owns live)
    size_t arena_bytes() const { return (arena_ ? arena_->used() : 0) + (mg_nranks > 1 ? mg_top_[mg_rank] : 0); }
    long long launches_total = 0;

   private:
    struct Cluster {
        int start, size, orig_size, level;
        bool sparsify, eliminated;
        int parent;                  // cluster id or -1
        int child_begin, child_end;  // cluster ids [begin,end)
        int hlevel;                  // hierarchy level at which it lives
    };
    struct SolveLevel {
        // forward order: elim trsv -> elim gemv -> scale trsv -> house -> merge copy
        TrsvTask* e_trsv = nullptr; int n_e_trsv = 0;
        GemvTask* e_gemv_f = nullptr; GemvContrib* e_gemv_fc = nullptr; int n_e_gemv_f = 0;
        GemvTask* e_gemv_b = nullptr; GemvContrib* e_gemv_bc = nullptr; int n_e_gemv_b = 0;
        TrsvTask* s_trsv = nullptr; int n_s_trsv = 0;
        HouseTask* house = nullptr; int n_house = 0;
        XCopyTask* m_fwd = nullptr; XCopyTask* m_bwd = nullptr; int n_merge = 0;
        int max_e = 0, max_s = 0;  // largest cluster at eliminate / scale time: sizes the solve launches
    };
    // device copies of the id arrays of one SymLevel (symbolic.hpp), resident in sym_arena_
    struct DevLevel {
        int *E = nullptr, *e_piv = nullptr, *S = nullptr, *s_piv = nullptr, *children = nullptr;
        SymTrsm *e_out = nullptr, *e_in = nullptr, *s_right = nullptr, *s_left = nullptr;
        SymGemm* e_gemm = nullptr;
        SymCon* e_con = nullptr;
        SymGemv *e_gf = nullptr, *e_gb = nullptr;
        SymGemvCon *e_gfc = nullptr, *e_gbc = nullptr;
        SymQrSrc* qs = nullptr;
        SymCopy* m_copy = nullptr;
        int n_children = 0;
    };

    DenseMat Xcoo_;
    bool have_coords_ = false;
    std::vector<Cluster> cl_;               // index == order id
    std::vector<std::vector<int>> bottoms_; // cluster ids per hierarchy level
    int current_bottom_ = 0, ilvl_ = 0;
    long long nnz_ = 0;
    bool factorized_ = false;
    bool assembled_ = false;
    int state_level_ = -1, state_phase_ = -1;  // last completed (level, phase) of factorize(); -1: just assembled
    std::vector<int> phases_done_;             // per level, bit p = phase p ran

    // symbolic plan: built by the first assemble() of a (partition, pattern) pair, reused afterwards
    SymbolicPlan plan_;
    bool plan_valid_ = false;
    int ord_serial_ = 0, plan_ord_serial_ = -1;
    void assemble_impl(const SpMat* Afull, int n, const int* colptr, const int* rowind, const double* val);
    std::vector<int> pat_colptr_, pat_rowind_;  // pattern the plan and the value map were built for
    std::vector<size_t> leaf_off_;              // element offset of every leaf block in the assembled buffer
    size_t leaf_total_ = 0;
    double t_analyze_ = 0;
    DeviceArena* sym_arena_ = nullptr;          // plan arrays + value map (persistent across assemble calls)
    std::vector<DevLevel> dplan_;
    unsigned* d_valmap_ = nullptr;
    int *d_en1_ = nullptr, *d_en2_ = nullptr, *d_parent_ = nullptr;
    void analyze(const SpMat& A);
    void analyze_host(const SpMat& A, std::vector<unsigned>& valmap);
    void build_clusters();
    bool plan_host_valid_ = false;

    struct DeviceGuard;
    std::vector<FlopTuple> flop_log_;
    cudaStream_t st_ = nullptr;
    static constexpr int kSide = 4;  // side streams for independent launches of one wavefront
    cudaStream_t side_[kSide] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork_ = nullptr, ev_join_[kSide] = {nullptr, nullptr, nullptr, nullptr};
    DeviceArena* arena_ = nullptr;    // blocks, factors, solve descriptors, x
    DeviceArena* scratch_ = nullptr;  // per-level descriptors + RRQR workspaces
    Stager stager_;
    int* d_err_ = nullptr;
    double* d_xnat_ = nullptr;  // N, natural-order staging for solve
    double* d_xleaf_ = nullptr;
    int* d_perm_ = nullptr;
    std::vector<SolveLevel> solve_;
    // numeric tables (host mirror + device): current cluster sizes, block pointers / leading dimensions, solution
    // segments, offsets of children inside their parents
    std::vector<int> h_csize_, h_eld_, h_pos_;
    std::vector<double*> h_eptr_, h_xptr_;
    int *d_csize_ = nullptr, *d_eld_ = nullptr, *d_pos_ = nullptr;
    double **d_eptr_ = nullptr, **d_xptr_ = nullptr;
    int *d_mid_ = nullptr, *d_cnt_ = nullptr;  // device work lists of the plan-driven kernels
    int cnt_next_ = 0;
    DevTables tab_{};
    // multi-GPU state: peer-mapped bases of the shared arenas, bump pointers of every rank's arena (all ranks
    // simulate all allocations, so every pointer table holds valid peer addresses), barrier epoch
    char* mg_base_[MG_MAX_RANKS] = {};
    size_t mg_top_[MG_MAX_RANKS] = {};
    size_t mg_size_ = 0, mg_off_csize_ = 0, mg_off_leaf_ = 0, mg_off_blocks_ = 0;
    bool mg_peers_set_ = false;
    unsigned mg_epoch_ = 0;
    PeerPtrs mg_flags_{}, mg_csize_{}, mg_leaf_{};
    std::vector<int> h_owner_;
    int* d_owner_ = nullptr;
    signed char* d_dof_owner_ = nullptr;
    bool mg() const { return mg_nranks > 1; }
    bool mine(int cluster) const { return mg_nranks == 1 || h_owner_[cluster] == mg_rank; }
    double* alloc_block(int owner, size_t doubles);
    void mg_barrier();
    // PLU: diag(U), swap sequence and permutation of the current pivot of every cluster
    std::vector<double*> h_ud_;
    std::vector<int*> h_ipiv_, h_pperm_;
    double** d_ud_ = nullptr;  // device tables of the two (plan-driven PLU kernels)
    int** d_pperm_ = nullptr;
    void upload_plu_tables();
    // cluster sizes before the elimination and after the sparsification of every level: the flop / byte / nnz
    // model is evaluated from these after the factorization (finalize_logs), off the critical path
    std::vector<std::vector<int>> size_pre_, size_post_;
    std::vector<std::vector<int>> qr_cols_;  // columns seen by every RRQR task (reference: Asn.cols())
    bool logs_final_ = true;
    void finalize_logs();

    bool symmetry() const { return symm_kind == SPD || symm_kind == SYM; }
    void ensure_device();
    void free_device();
    template <class T>
    T* to_device(const std::vector<T>& v, DeviceArena* where);
    int* next_counter();

    void phase_eliminate(LevelLog& lg, SolveLevel& sl);
    void phase_scale(LevelLog& lg, SolveLevel& sl);
    void phase_sparsify(LevelLog& lg, SolveLevel& sl);
    void phase_merge(LevelLog& lg, SolveLevel& sl);
    void phase_eliminate_plu(LevelLog& lg, SolveLevel& sl);
    void phase_scale_plu(LevelLog& lg, SolveLevel& sl);
    void alloc_edges(int e0, int e1, bool zero, LevelLog& lg);
    void run_getrf(std::vector<GetrfTask>& tasks, LevelLog& lg);
    void run_rowperm(std::vector<RowPermTask>& tasks, LevelLog& lg);
    void alloc_plu(int c);
    void run_potrf(std::vector<PotrfTask>& tasks, LevelLog& lg);
    void run_trsm(int mode, std::vector<TrsmTask>& tasks, LevelLog& lg);
    // two-sided scaling through explicit inverses of the pivots' Cholesky factors (two grouped GEMM launches)
    void run_scale_inv(int min_dim, LevelLog& lg);
    int scale_inv_mode_ = -1;  // SPAND_SCALE_INV: 0 triangular solves, 1 blocks with a dimension > 64, 2 (default) > 32
    void run_gemm(std::vector<GemmTask>& tasks, std::vector<GemmContrib>& contribs, LevelLog& lg);
    void check_error();
    int ndofs_left() const;
    int level_max_size() const;
    TrsmTask host_trsm(const SymTrsm& t, const double* diag) const;
    struct FamEvent { int fam; cudaEvent_t a, b; };
    std::vector<FamEvent> fam_events_;
    std::vector<cudaEvent_t> ev_pool_;
    cudaEvent_t fam_begin(int fam);
    void fam_end(int fam, cudaEvent_t a);
    void fam_resolve();
};

}  // namespace spand
