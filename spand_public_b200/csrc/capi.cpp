// C ABI of the B200 path (include/spand_b200.h). Thin: argument marshalling + error mapping only.
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

#include "../../include/spand_b200.h"
#include "tree.hpp"

using namespace spand;

struct spand_tree {
    Tree t;
    std::string err;
    explicit spand_tree(int nl) : t(nl) {}
};

namespace {
template <class F>
int guarded(spand_tree* h, F&& f) {
    try {
        f();
        return 0;
    } catch (std::exception& e) {
        h->err = e.what();
        if (h->err.find("Non-SPD") != std::string::npos) return 1;
        if (h->err.find("Singular") != std::string::npos) return 2;
        return -1;
    }
}
const char* kLogNames[] = {"dofs_nd", "dofs_left_nd", "dofs_left_elim", "dofs_left_spars", "fact_nnz", "rank_before",
                           "rank_after", "nspars", "ignored", "nbrs", "t_elim", "t_scale", "t_spars", "t_merge",
                           "fl_pivot", "fl_panel", "fl_schur", "fl_rrqr_rank", "fl_rrqr_full", "by_scale", "by_rrqr",
                           "by_merge", "t_host", "launches", "wavefronts", "t_plan_elim", "t_plan_scale", "t_plan_spars",
                           "t_plan_merge"};
constexpr int kLogFields = sizeof(kLogNames) / sizeof(kLogNames[0]);
}  // namespace

extern "C" {

spand_tree* spand_create(int nlevels) {
    try {
        return new spand_tree(nlevels);
    } catch (...) {
        return nullptr;
    }
}
void spand_destroy(spand_tree* t) { delete t; }
const char* spand_last_error(spand_tree* t) { return t->err.c_str(); }

int spand_set_tol(spand_tree* t, double tol) { t->t.tol = tol; return 0; }
int spand_set_skip(spand_tree* t, int skip) { t->t.skip = skip; return 0; }
int spand_set_symm_kind(spand_tree* t, int kind) { t->t.symm_kind = kind; return 0; }
int spand_set_scaling_kind(spand_tree* t, int kind) { t->t.scale_kind = kind; return 0; }
int spand_set_use_geo(spand_tree* t, int geo) { t->t.use_geo = geo != 0; return 0; }
int spand_set_verb(spand_tree* t, int verb) { t->t.verb = verb != 0; return 0; }
int spand_set_use_sparsify(spand_tree* t, int use) { t->t.use_want_sparsify = use != 0; return 0; }
int spand_set_device(spand_tree* t, int device) {
    return guarded(t, [&] {
        if (t->t.device_in_use() && device != t->t.device)
            throw std::runtime_error("set_device: the tree already holds memory on another device");
        t->t.device = device;
    });
}
int spand_set_monitor_flops(spand_tree* t, int on) { t->t.monitor_flops = on != 0; return 0; }
long long spand_get_flops_log(spand_tree* t, long long* out) {
    long long n = -1;
    guarded(t, [&] {
        const auto& fl = t->t.flop_log();
        if (out)
            for (size_t i = 0; i < fl.size(); i++) {
                out[5 * i] = fl[i].lvl;
                out[5 * i + 1] = fl[i].kind;
                out[5 * i + 2] = fl[i].rows;
                out[5 * i + 3] = fl[i].cols;
                out[5 * i + 4] = fl[i].inner;
            }
        n = (long long)fl.size();
    });
    return n;
}
int spand_set_profile(spand_tree* t, int on) { t->t.profile_families = on != 0; return 0; }
int spand_num_families(void) { return Tree::F_COUNT; }
const char* spand_family_name(int f) { return Tree::family_name(f); }
int spand_get_family_stats(spand_tree* t, double* ms, long long* launches) {
    for (int f = 0; f < Tree::F_COUNT; f++) {
        ms[f] = t->t.family_ms[f];
        launches[f] = t->t.family_launches[f];
    }
    return 0;
}
int spand_set_stop(spand_tree* t, int level, int phase) {
    t->t.stop_level = level;
    t->t.stop_phase = phase;
    return 0;
}
int spand_set_coords(spand_tree* t, int dim, int N, const double* X) {
    return guarded(t, [&] { t->t.set_coords(dim, N, X); });
}
int spand_partition(spand_tree* t, int N, const int* colptr, const int* rowind) {
    return guarded(t, [&] { t->t.partition(from_csc(N, colptr, rowind, nullptr)); });
}
int spand_get_partition(spand_tree* t, int* self_lvl, int* self_sep, int* l_lvl, int* l_sep, int* r_lvl, int* r_sep) {
    auto& part = t->t.ord.part;
    for (size_t i = 0; i < part.size(); i++) {
        self_lvl[i] = part[i].self.lvl;
        self_sep[i] = part[i].self.sep;
        l_lvl[i] = part[i].l.lvl;
        l_sep[i] = part[i].l.sep;
        r_lvl[i] = part[i].r.lvl;
        r_sep[i] = part[i].r.sep;
    }
    return 0;
}
int spand_get_perm(spand_tree* t, int* perm) {
    std::memcpy(perm, t->t.ord.perm.data(), sizeof(int) * t->t.ord.perm.size());
    return 0;
}
int spand_get_N(spand_tree* t) { return t->t.N; }
int spand_assemble(spand_tree* t, int N, const int* colptr, const int* rowind, const double* val) {
    return guarded(t, [&] { t->t.assemble_csc(N, colptr, rowind, val); });
}
int spand_factorize(spand_tree* t) {
    return guarded(t, [&] { t->t.factorize(); });
}
int spand_solve(spand_tree* t, double* x) {
    return guarded(t, [&] { t->t.solve(x); });
}
int spand_solve_device(spand_tree* t, double* x) {
    return guarded(t, [&] { t->t.solve_device(x); });
}
int spand_cg(spand_tree* t, int N, const int* colptr, const int* rowind, const double* val, const double* rhs, double* x,
             int iters, double tol, int verb, double* seconds) {
    int it = -1;
    int rc = guarded(t, [&] { it = t->t.cg(from_csc(N, colptr, rowind, val), rhs, x, iters, tol, verb != 0, seconds); });
    return rc == 0 ? it : -1;
}
int spand_gmres(spand_tree* t, int N, const int* colptr, const int* rowind, const double* val, const double* rhs,
                double* x, int iters, int restart, double tol, int verb, double* seconds) {
    int it = -1;
    int rc = guarded(t, [&] {
        it = t->t.gmres(from_csc(N, colptr, rowind, val), rhs, x, iters, restart, tol, verb != 0, seconds);
    });
    return rc == 0 ? it : -1;
}
int spand_geqp3_truncated(int rows, int cols, const double* A, int nsrc, int transposed, double tol, int G,
                          int nthreads, int in_smem, int nb, double theta, int* rank, double* R, double* V, double* tau) {
    static thread_local std::string err;
    const int rc = spand::rrqr_single(rows, cols, A, nsrc, transposed, tol, G, nthreads, in_smem, nb, theta, rank, R, V,
                                      tau, err);
    if (rc != 0) fprintf(stderr, "spand_geqp3_truncated: %s\n", err.c_str());
    return rc;
}
void spand_debug_hc2_stats(unsigned long long* out16, int reset) { spand::hc2_stats(out16, reset != 0); }
void spand_debug_rrqr_phases(unsigned long long* out48, int reset) { spand::rrqr_phase_cycles(out48, reset != 0); }
long long spand_nnz(spand_tree* t) { return t->t.nnz(); }
int spand_get_stop(spand_tree* t) { return t->t.get_stop(); }
int spand_get_nlevels(spand_tree* t) { return t->t.nlevels; }
int spand_num_clusters(spand_tree* t) {
    int n = 0;
    for (auto& l : t->t.ord.levels) n += (int)l.size();
    return n;
}
int spand_get_stats(spand_tree* t, int* id, int* size, int* rank) {
    std::vector<int> a, b, c;
    t->t.stats(a, b, c);
    std::memcpy(id, a.data(), sizeof(int) * a.size());
    std::memcpy(size, b.data(), sizeof(int) * b.size());
    std::memcpy(rank, c.data(), sizeof(int) * c.size());
    return 0;
}
int spand_get_cluster_layout(spand_tree* t, int* start, int* hlevel) {
    std::vector<int> a, b;
    t->t.cluster_layout(a, b);
    std::memcpy(start, a.data(), sizeof(int) * a.size());
    std::memcpy(hlevel, b.data(), sizeof(int) * b.size());
    return 0;
}
int spand_log_fields(void) { return kLogFields; }
const char* spand_log_field_name(int i) { return (i >= 0 && i < kLogFields) ? kLogNames[i] : ""; }
int spand_get_log(spand_tree* t, double* out) {
    for (int l = 0; l < t->t.nlevels; l++) {
        const LevelLog& g = t->t.logs()[l];
        double v[kLogFields] = {(double)g.dofs_nd, (double)g.dofs_left_nd, (double)g.dofs_left_elim,
                                (double)g.dofs_left_spars, (double)g.fact_nnz, (double)g.rank_before,
                                (double)g.rank_after, (double)g.nspars, (double)g.ignored, (double)g.nbrs, g.t_elim,
                                g.t_scale, g.t_spars, g.t_merge, g.fl_pivot, g.fl_panel, g.fl_schur, g.fl_rrqr_rank,
                                g.fl_rrqr_full, g.by_scale, g.by_rrqr, g.by_merge, g.t_host, (double)g.launches,
                                (double)g.wavefronts, g.t_plan_elim, g.t_plan_scale, g.t_plan_spars, g.t_plan_merge};
        std::memcpy(out + (size_t)kLogFields * l, v, sizeof(v));
    }
    return 0;
}
double spand_factorize_seconds(spand_tree* t) { return t->t.t_factorize_device; }
double spand_analyze_seconds(spand_tree* t) { return t->t.analyze_seconds(); }
int spand_mg_setup(spand_tree* t, int rank, int nranks, long long arena_bytes) {
    return guarded(t, [&] { t->t.mg_setup(rank, nranks, (size_t)arena_bytes); });
}
int spand_mg_get_handle(spand_tree* t, void* out64) {
    return guarded(t, [&] { t->t.mg_get_handle(out64); });
}
int spand_mg_set_peers(spand_tree* t, const void* handles) {
    return guarded(t, [&] { t->t.mg_set_peers(handles); });
}
int spand_mg_owner_map(spand_tree* t, int nranks, int* owner) {
    return guarded(t, [&] {
        std::vector<int> o = t->t.owner_map(nranks);
        std::memcpy(owner, o.data(), sizeof(int) * o.size());
    });
}
int spand_plan_analyze(spand_tree* t, int N, const int* colptr, const int* rowind) {
    return guarded(t, [&] { t->t.analyze_only(from_csc(N, colptr, rowind, nullptr)); });
}
int spand_plan_live_edges(spand_tree* t, int level, int phase, int* n1, int* n2) {
    int n = -1;
    guarded(t, [&] {
        std::vector<int> a, b;
        t->t.plan_live_edges(level, phase, a, b);
        if (n1) {
            std::memcpy(n1, a.data(), sizeof(int) * a.size());
            std::memcpy(n2, b.data(), sizeof(int) * b.size());
        }
        n = (int)a.size();
    });
    return n;
}
int spand_plan_counts(spand_tree* t, int level, long long* out) {
    return guarded(t, [&] { t->t.plan_counts(level, out); });
}
long long spand_kernel_launches(spand_tree* t) { return t->t.launches_total; }
long long spand_arena_bytes(spand_tree* t) { return (long long)t->t.arena_bytes(); }
int spand_trailing(spand_tree* t, int* colptr, int* rowind, double* val) {
    int nnz = -1;
    guarded(t, [&] {
        SpMat T = t->t.trailing_mat();
        if (colptr) {
            std::memcpy(colptr, T.colptr.data(), sizeof(int) * T.colptr.size());
            std::memcpy(rowind, T.rowind.data(), sizeof(int) * T.rowind.size());
            std::memcpy(val, T.val.data(), sizeof(double) * T.val.size());
        }
        nnz = T.nnz();
    });
    return nnz;
}

void spand_util_random(int size, int seed, double* out) {
    std::vector<double> v = random_vec(size, seed);
    std::memcpy(out, v.data(), sizeof(double) * size);
}
void spand_util_linspace_nd(int n, int dim, double* out) {
    DenseMat X = linspace_nd(n, dim);
    std::memcpy(out, X.a.data(), sizeof(double) * X.a.size());
}
static int export_csc(const SpMat& A, int* colptr, int* rowind, double* val) {
    if (colptr) {
        std::memcpy(colptr, A.colptr.data(), sizeof(int) * A.colptr.size());
        std::memcpy(rowind, A.rowind.data(), sizeof(int) * A.rowind.size());
        std::memcpy(val, A.val.data(), sizeof(double) * A.val.size());
    }
    return A.nnz();
}
int spand_util_neglapl(int n, int d, int* colptr, int* rowind, double* val) {
    return export_csc(neglapl(n, d), colptr, rowind, val);
}
int spand_util_aniso(int n, int* colptr, int* rowind, double* val) {
    return export_csc(aniso_convdiff(n), colptr, rowind, val);
}
int spand_util_mm_read(const char* fn, int* rows, int* cols, int* colptr, int* rowind, double* val) {
    try {
        SpMat A = mm_read_sparse(fn);
        *rows = A.rows;
        *cols = A.cols;
        return export_csc(A, colptr, rowind, val);
    } catch (...) {
        return -1;
    }
}
int spand_util_mm_read_dense(const char* fn, int* rows, int* cols, double* out) {
    try {
        DenseMat A = mm_read_dense(fn);
        *rows = A.rows;
        *cols = A.cols;
        if (out) std::memcpy(out, A.a.data(), sizeof(double) * A.a.size());
        return 0;
    } catch (...) {
        return -1;
    }
}
}
