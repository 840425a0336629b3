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
    int stop_level = -1, stop_phase = -1;  // parity-test hook (see oracle)
    int device = 0;
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
    void factorize();                 // throws std::runtime_error("Error: Non-SPD Pivot\n") etc.
    void solve(double* x_host);       // in place, host vector of length N
    void solve_device(double* x_dev); // in place, device vector of length N (natural ordering)
    int cg(const SpMat& A, const double* rhs, double* x, int iters, double tol, bool verb, double* seconds);
    long long nnz() const { return nnz_; }
    int get_stop() const;
    SpMat trailing_mat();
    void stats(std::vector<int>& id, std::vector<int>& size, std::vector<int>& rank) const;

    int nlevels;
    int N = 0;
    Ordering ord;
    std::vector<LevelLog> log;
    double t_factorize_device = 0;  // seconds, CUDA events
    size_t arena_bytes() const { return arena_ ? arena_->used() : 0; }
    long long launches_total = 0;

   private:
    struct Cluster {
        int start, size, orig_size, level;
        bool sparsify, eliminated;
        int parent;                  // cluster id or -1
        int child_begin, child_end;  // cluster ids [begin,end)
        int hlevel;                  // hierarchy level at which it lives
        std::vector<int> out;        // edge ids, pivot first
        std::vector<int> in;         // edge ids
        double* x = nullptr;         // device solution segment (orig_size)
        double* ud = nullptr;        // PLU: diag(U), swap sequence and permutation of the current pivot
        int* ipiv = nullptr;
        int* perm = nullptr;
    };
    struct Edge {
        int n1, n2;       // block A[rows of n2, cols of n1]
        double* A;        // device
        int ld;
        bool original, alive, identity;
    };
    struct SolveLevel {
        // forward order: elim trsv -> elim gemv -> scale trsv -> house -> merge copy
        TrsvTask* e_trsv = nullptr; int n_e_trsv = 0;
        GemvTask* e_gemv_f = nullptr; GemvContrib* e_gemv_fc = nullptr; int n_e_gemv_f = 0;
        GemvTask* e_gemv_b = nullptr; GemvContrib* e_gemv_bc = nullptr; int n_e_gemv_b = 0;
        TrsvTask* s_trsv = nullptr; int n_s_trsv = 0;
        HouseTask* house = nullptr; int n_house = 0;
        XCopyTask* m_fwd = nullptr; XCopyTask* m_bwd = nullptr; int n_merge = 0;
    };

    DenseMat Xcoo_;
    bool have_coords_ = false;
    std::vector<Cluster> cl_;               // index == order id
    std::vector<std::vector<int>> bottoms_; // cluster ids per hierarchy level
    std::vector<Edge> ed_;
    int current_bottom_ = 0, ilvl_ = 0;
    long long nnz_ = 0;
    bool factorized_ = false;

    cudaStream_t st_ = nullptr;
    static constexpr int kSide = 4;  // side streams for independent launches of one wavefront
    cudaStream_t side_[kSide] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork_ = nullptr, ev_join_[kSide] = {nullptr, nullptr, nullptr, nullptr};
    DeviceArena* arena_ = nullptr;    // blocks, factors, solve descriptors, x
    DeviceArena* scratch_ = nullptr;  // per-level descriptors + RRQR workspaces
    Stager stager_;
    int* d_csize_ = nullptr;
    int* d_err_ = nullptr;
    double* d_xnat_ = nullptr;  // N, natural-order staging for solve
    int* d_perm_ = nullptr;
    std::vector<SolveLevel> solve_;
    std::vector<int> h_csize_;

    bool symmetry() const { return symm_kind == SPD || symm_kind == SYM; }
    void ensure_device();
    void free_device();
    int find_out(int c, int n2) const;
    int new_edge(int n1, int n2, double* A, int ld, bool original);
    template <class T>
    T* to_device(const std::vector<T>& v, DeviceArena* where);

    void phase_eliminate(LevelLog& lg, SolveLevel& sl);
    void phase_scale(LevelLog& lg, SolveLevel& sl);
    void phase_sparsify(LevelLog& lg, SolveLevel& sl);
    void phase_merge(LevelLog& lg, SolveLevel& sl);
    void phase_eliminate_plu(LevelLog& lg, SolveLevel& sl);
    void phase_scale_plu(LevelLog& lg, SolveLevel& sl);
    void run_getrf(std::vector<GetrfTask>& tasks, LevelLog& lg);
    void run_rowperm(std::vector<RowPermTask>& tasks, LevelLog& lg);
    void alloc_plu(Cluster& cs);
    void run_potrf(std::vector<PotrfTask>& tasks, LevelLog& lg);
    void run_trsm(int mode, std::vector<TrsmTask>& tasks, LevelLog& lg);
    void run_gemm(std::vector<GemmTask>& tasks, std::vector<GemmContrib>& contribs, LevelLog& lg);
    void check_error();
    int ndofs_left() const;
    struct FamEvent { int fam; cudaEvent_t a, b; };
    std::vector<FamEvent> fam_events_;
    std::vector<cudaEvent_t> ev_pool_;
    cudaEvent_t fam_begin(int fam);
    void fam_end(int fam, cudaEvent_t a);
    void fam_resolve();
};

}  // namespace spand
