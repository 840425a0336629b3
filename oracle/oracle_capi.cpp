// ORACLE — TEST INFRASTRUCTURE ONLY. Plain-C entry points so tests/ and bench.py can drive the CPU
// restatement through ctypes. Mirrors the product's C ABI (include/spand_b200.h) one to one.
#include <cstring>
#include <stdexcept>
#include <string>

#include "spand_oracle.hpp"

using namespace spand_oracle;

namespace {
struct Handle {
    OTree t;
    std::string err;
    explicit Handle(int nl) : t(nl) {}
};
template <class F>
int guarded(Handle* h, F&& f) {
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
}  // namespace

extern "C" {

void* orc_create(int nlevels) {
    try {
        return new Handle(nlevels);
    } catch (...) {
        return nullptr;
    }
}
void orc_destroy(void* h) { delete (Handle*)h; }
const char* orc_last_error(void* h) { return ((Handle*)h)->err.c_str(); }

void orc_set_params(void* h_, double tol, int skip, int symm_kind, int scale_kind, int use_geo, int verb,
                    int use_want_sparsify) {
    OTree& t = ((Handle*)h_)->t;
    t.tol = tol;
    t.skip = skip;
    t.symm_kind = symm_kind;
    t.scale_kind = scale_kind;
    t.use_geo = use_geo != 0;
    t.verb = verb != 0;
    t.use_want_sparsify = use_want_sparsify != 0;
}
void orc_set_stop(void* h_, int level, int phase) {
    OTree& t = ((Handle*)h_)->t;
    t.stop_level = level;
    t.stop_phase = phase;
}
int orc_set_coords(void* h_, int dim, int N, const double* X) {
    Handle* h = (Handle*)h_;
    return guarded(h, [&] { h->t.set_coords(dim, N, X); });
}
int orc_partition(void* h_, int N, const int* colptr, const int* rowind) {
    Handle* h = (Handle*)h_;
    return guarded(h, [&] { h->t.partition(spand::from_csc(N, colptr, rowind, nullptr)); });
}
int orc_assemble(void* h_, int N, const int* colptr, const int* rowind, const double* val) {
    Handle* h = (Handle*)h_;
    return guarded(h, [&] { h->t.assemble(spand::from_csc(N, colptr, rowind, val)); });
}
int orc_factorize(void* h_) {
    Handle* h = (Handle*)h_;
    return guarded(h, [&] { h->t.factorize(); });
}
int orc_solve(void* h_, double* x) {
    Handle* h = (Handle*)h_;
    return guarded(h, [&] { h->t.solve(x); });
}
int orc_cg(void* h_, int N, const int* colptr, const int* rowind, const double* val, const double* rhs, double* x,
           int iters, double tol, int verb) {
    Handle* h = (Handle*)h_;
    int it = -1;
    guarded(h, [&] { it = cg(spand::from_csc(N, colptr, rowind, val), rhs, x, h->t, iters, tol, verb != 0); });
    return it;
}
int orc_gmres(void* h_, int N, const int* colptr, const int* rowind, const double* val, const double* rhs, double* x,
              int iters, int restart, double tol, int verb) {
    Handle* h = (Handle*)h_;
    int it = -1;
    guarded(h, [&] { it = gmres(spand::from_csc(N, colptr, rowind, val), rhs, x, h->t, iters, restart, tol, verb != 0); });
    return it;
}
long long orc_nnz(void* h_) { return ((Handle*)h_)->t.nnz(); }
int orc_get_stop(void* h_) { return ((Handle*)h_)->t.get_stop(); }
int orc_get_N(void* h_) { return ((Handle*)h_)->t.N; }
void orc_get_perm(void* h_, int* perm) {
    OTree& t = ((Handle*)h_)->t;
    std::memcpy(perm, t.ord.perm.data(), sizeof(int) * t.ord.perm.size());
}
void orc_get_partition(void* h_, int* self_lvl, int* self_sep, int* l_lvl, int* l_sep, int* r_lvl, int* r_sep) {
    OTree& t = ((Handle*)h_)->t;
    for (size_t i = 0; i < t.ord.part.size(); i++) {
        auto& p = t.ord.part[i];
        self_lvl[i] = p.self.lvl;
        self_sep[i] = p.self.sep;
        l_lvl[i] = p.l.lvl;
        l_sep[i] = p.l.sep;
        r_lvl[i] = p.r.lvl;
        r_sep[i] = p.r.sep;
    }
}
int orc_num_clusters(void* h_) {
    OTree& t = ((Handle*)h_)->t;
    int n = 0;
    for (auto& b : t.bottoms) n += (int)b.size();
    return n;
}
void orc_get_stats(void* h_, int* id, int* size, int* rank) {
    std::vector<int> a, b, c;
    ((Handle*)h_)->t.stats(a, b, c);
    std::memcpy(id, a.data(), sizeof(int) * a.size());
    std::memcpy(size, b.data(), sizeof(int) * b.size());
    std::memcpy(rank, c.data(), sizeof(int) * c.size());
}
// per level: dofs_nd, dofs_left_nd, dofs_left_elim, dofs_left_spars, fact_nnz, rank_before, rank_after, nspars,
// ignored, nbrs, t_elim, t_scale, t_spars, t_merge, fl_pivot, fl_panel, fl_schur, fl_rrqr_rank, fl_rrqr_full,
// by_scale, by_rrqr, by_merge  (22 doubles)
int orc_log_fields() { return 22; }
void orc_set_monitor_flops(void* h_, int on) { ((Handle*)h_)->t.monitor_flops = on != 0; }
// two-call protocol: out == nullptr returns the number of tuples; else fills 5 long long per tuple
long long orc_get_flops_log(void* h_, long long* out) {
    OTree& t = ((Handle*)h_)->t;
    if (out)
        for (size_t i = 0; i < t.flop_log.size(); i++)
            for (int k = 0; k < 5; k++) out[5 * i + k] = t.flop_log[i][k];
    return (long long)t.flop_log.size();
}
void orc_get_log(void* h_, double* out) {
    OTree& t = ((Handle*)h_)->t;
    for (int l = 0; l < t.nlevels; l++) {
        const LevelLog& g = t.log[l];
        double v[22] = {(double)g.dofs_nd, (double)g.dofs_left_nd, (double)g.dofs_left_elim, (double)g.dofs_left_spars,
                        (double)g.fact_nnz, (double)g.rank_before, (double)g.rank_after, (double)g.nspars,
                        (double)g.ignored, (double)g.nbrs, g.t_elim, g.t_scale, g.t_spars, g.t_merge, g.fl_pivot,
                        g.fl_panel, g.fl_schur, g.fl_rrqr_rank, g.fl_rrqr_full, g.by_scale, g.by_rrqr, g.by_merge};
        std::memcpy(out + 22 * l, v, sizeof(v));
    }
}
// Trailing matrix (tree.h get_trailing_mat) in permuted ordering, CSC. Two-call protocol.
int orc_trailing(void* h_, int* colptr, int* rowind, double* val) {
    OTree& t = ((Handle*)h_)->t;
    SpMat T = t.trailing_mat();
    if (colptr) {
        std::memcpy(colptr, T.colptr.data(), sizeof(int) * T.colptr.size());
        std::memcpy(rowind, T.rowind.data(), sizeof(int) * T.rowind.size());
        std::memcpy(val, T.val.data(), sizeof(double) * T.val.size());
    }
    return T.nnz();
}

// Known-answer hooks (tests/tests.cpp:254-262, :349-357, :264-305)
int orc_choose_rank(const double* s, int n, double tol) { return choose_rank(s, n, tol); }
void orc_swap2perm(const int* swap, int n, int* perm) {
    std::vector<int> s(swap, swap + n), p;
    swap2perm(s, p);
    std::memcpy(perm, p.data(), sizeof(int) * n);
}
void orc_block2dense(int ncols, const int* colptr, const int* rowval, const double* nnzval, int i, int j, int li, int lj,
                     double* dst, int transpose) {
    std::vector<int> cp(colptr, colptr + ncols + 1), rv(rowval, rowval + colptr[ncols]);
    std::vector<double> nv(nnzval, nnzval + colptr[ncols]);
    DenseMat D = transpose ? DenseMat(lj, li) : DenseMat(li, lj);
    block2dense(rv, cp, nv, i, j, li, lj, &D, transpose != 0);
    std::memcpy(dst, D.a.data(), sizeof(double) * D.a.size());
}
void orc_set_threads(int n) { set_blas_threads(n); }
}
