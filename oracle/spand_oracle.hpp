// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the reference's factorization hot path (spaND, leopoldcambier/spaND_public),
// used (a) as the parity checker for the CUDA path and (b) as the timed CPU baseline of bench.py.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library; nothing under spand_public_b200/ links or calls it.
//
// It issues the same BLAS/LAPACK call sequence as the reference, one call per block, sequentially
// over clusters in list order (dpotrf, dtrsm, dgemm/dsyrk, dgetrf, dgeqp3, dormqr, dtrsv, dgemv),
// against OpenBLAS from the scipy wheel (symbols prefixed scipy_).
//
// Pinning: the reference cannot be compiled in this image (Eigen, metis.h, cblas.h/lapacke.h and
// googletest are absent, no network), so the oracle is pinned against every known-answer test the
// reference holds on this path (tests/tests.cpp): Util.ChooseRank :254-262, Util.swap2perm :349-357,
// Util.Block2Dense :264-305, Util.LinspaceNd :307-316, PartitionTest.Square :378-412,
// PartitionTest.Consistency :417-481, Assembly.Consistency :488-547, ApproxTest.Exact :562-609,
// ApproxTest.Approx :799-856, ApproxTest.Repro :917-994 — see tests/test_oracle_*.py.
// The reference holds no golden vectors for factor entries, ranks or iteration counts.
//
// The integer front end (modified ND, ordering, hierarchy) is shared with the product
// (spand_public_b200/csrc/host/partition.cpp) because north_star requires it bit-exact on both sides.
#pragma once
#include <array>
#include <list>
#include <memory>
#include <string>
#include <vector>

#include "partition.hpp"
#include "sparse.hpp"

namespace spand_oracle {

using spand::DenseMat;
using spand::SpMat;

enum SymmKind { SPD = 0, SYM = 1, GEN = 2 };
enum ScalingKind { LLT = 0, PLU = 3 };

struct OCluster;

// include/edge.h:32-44 — block A[rows of n2, cols of n1]
struct OEdge {
    OCluster* n1;
    OCluster* n2;
    bool original;
    DenseMat A;
};

// include/cluster.h:16-116
struct OCluster {
    int start, size, level, order;
    bool eliminated = false, sparsify = false;
    OCluster* parent = nullptr;
    std::vector<OCluster*> children;
    std::list<std::unique_ptr<OEdge>> out;  // pivot first
    std::list<OEdge*> in;
    std::vector<double> x;  // length = original size
    int original_size() const { return (int)x.size(); }
    OEdge* pivot() const { return out.front().get(); }
};

// include/operations.h:32-163
struct OOp {
    enum Kind { ScalingLLT, ScalingPLU, GemmSymmOut, GemmSymmIn, GemmOut, GemmIn, Orthogonal, Merge, Split } kind;
    OCluster* self = nullptr;
    OCluster* nbr = nullptr;
    int nself = 0, nnbr = 0;  // segment lengths captured at record time
    DenseMat A;               // L / panel / V
    DenseMat U;               // PLU only
    std::vector<int> p, q;    // PLU only
    std::vector<double> tau;  // Orthogonal only
    long long nnz() const;
};

struct LevelLog {
    int dofs_nd = 0, dofs_left_nd = 0, dofs_left_elim = 0, dofs_left_spars = 0;
    long long fact_nnz = 0;
    long long rank_before = 0, rank_after = 0;
    int nspars = 0, ignored = 0;
    long long nbrs = 0;
    double t_elim = 0, t_scale = 0, t_spars = 0, t_merge = 0;
    // flop model of SURVEY.md section 8(d)
    double fl_pivot = 0, fl_panel = 0, fl_schur = 0, fl_rrqr_rank = 0, fl_rrqr_full = 0;
    double by_scale = 0, by_rrqr = 0, by_merge = 0;
};

class OTree {
   public:
    explicit OTree(int nlevels);
    bool verb = false;
    bool use_geo = false;
    double tol = 10.0;
    int skip = 0;
    int symm_kind = SPD;
    int scale_kind = LLT;
    bool use_want_sparsify = true;
    // debug hook for parity tests: stop after (level, phase); phase 0=elim 1=scale 2=sparsify 3=merge
    int stop_level = -1, stop_phase = -1;

    void set_coords(int dim, int N, const double* X);
    void partition(const SpMat& A);
    void assemble(const SpMat& A);
    void factorize();  // throws std::runtime_error like tree.cpp:587-590 / :625-628
    void solve(double* x);
    long long nnz() const;
    int get_stop() const;
    SpMat trailing_mat() const;

    spand::Ordering ord;
    std::vector<LevelLog> log;
    int nlevels;
    int N = 0;
    int ilvl = 0;
    int current_bottom = 0;
    std::vector<std::list<std::unique_ptr<OCluster>>> bottoms;
    std::vector<std::unique_ptr<OCluster>> others;
    std::list<OOp> ops;
    // per-call flop tuples (level, kind 0 pivot / 1 panel / 2 gemm / 3 rrqr, rows, cols, inner) exactly where the
    // reference pushes them when set_monitor_flops is on (tree.cpp:592,648,662,792,1312; written by write_log_flops
    // tree.cpp:60-77)
    bool monitor_flops = false;
    std::vector<std::array<long, 5>> flop_log;
    void flop(int kind, long rows, long cols, long inner) {
        if (monitor_flops) flop_log.push_back({(long)ilvl, (long)kind, rows, cols, inner});
    }
    // (order, original size, final size) for every hierarchy cluster — the write_stats triple (tree.h:242-251)
    void stats(std::vector<int>& id, std::vector<int>& size, std::vector<int>& rank) const;

   private:
    DenseMat Xcoo;
    bool have_coords = false;
    int max_order = 0;
    bool symmetry() const { return symm_kind == SPD || symm_kind == SYM; }
    void eliminate_cluster(OCluster* s);
    void scale_cluster(OCluster* s);
    void sparsify_cluster(OCluster* s);
    void merge_all();
    void schur_update(OEdge* e1, OEdge* e2, bool t1, bool t2);
    void set_eliminated(OCluster* s);
    void add_edge(OCluster* c, std::unique_ptr<OEdge> e);
    void sort_edges(OCluster* c);
    int ndofs_left() const;
};

int choose_rank(const double* s, int n, double tol);             // src/util.cpp:434-452
void swap2perm(const std::vector<int>& swap, std::vector<int>& perm);  // src/util.cpp:76-88
// src/util.cpp:454-486
void block2dense(const std::vector<int>& rowval, const std::vector<int>& colptr, const std::vector<double>& nnzval,
                 int i, int j, int li, int lj, DenseMat* dst, bool transpose);
// src/is.cpp:39-121; returns the iteration count exactly as the reference does (i+1)
int cg(const SpMat& A, const double* rhs, double* x, OTree& precond, int iters, double tol, bool verb);
// src/is.cpp:123-300 (Householder GMRES, left-preconditioned)
int gmres(const SpMat& A, const double* rhs, double* x, OTree& precond, int iters, int restart, double tol, bool verb);

void set_blas_threads(int n);

}  // namespace spand_oracle
