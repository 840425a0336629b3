// spaND_b200.hpp — header-only C++ facade over the C ABI (spand_b200.h) with the reference's spaND::Tree surface
// (reference include/tree.h:130-198, include/is.h:12) so that tests/spaND.cpp-style drivers switch by changing
// the namespace. Works without Eigen (std::vector based CSC); Eigen overloads appear when <Eigen/SparseCore> exists.
#pragma once
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "spand_b200.h"

#if defined(__has_include)
#if __has_include(<Eigen/SparseCore>)
#include <Eigen/Core>
#include <Eigen/SparseCore>
#define SPAND_B200_HAVE_EIGEN 1
#endif
#endif

namespace spaND_b200 {

enum class ScalingKind { LLT = SPAND_LLT, PLU = SPAND_PLU };   // include/spaND.h:19 (SVD/EVD/PLUQ/LDLT: out of scope)
enum class SymmKind { SPD = SPAND_SPD, SYM = SPAND_SYM, GEN = SPAND_GEN };  // include/spaND.h:21

// column-compressed matrix with int32 indices, the layout of Eigen::SparseMatrix<double,0,int> (include/util.h SpMat)
struct CscView {
    int rows = 0, cols = 0;
    const int* colptr = nullptr;
    const int* rowind = nullptr;
    const double* val = nullptr;
};

struct ClusterID6 { int self_lvl, self_sep, l_lvl, l_sep, r_lvl, r_sep; };  // include/partition.h:52-71

class Tree {
   public:
    explicit Tree(int nlevels) : h_(spand_create(nlevels)) {
        if (!h_) throw std::runtime_error("spand_create failed");
    }
    ~Tree() { spand_destroy(h_); }
    Tree(const Tree&) = delete;
    Tree& operator=(const Tree&) = delete;

    void set_verb(bool v) { spand_set_verb(h_, v); }
    void set_tol(double t) { spand_set_tol(h_, t); }
    void set_skip(int s) { spand_set_skip(h_, s); }
    void set_use_geo(bool g) { spand_set_use_geo(h_, g); }
    void set_use_sparsify(bool u) { spand_set_use_sparsify(h_, u); }
    void set_scaling_kind(ScalingKind k) { spand_set_scaling_kind(h_, (int)k); }
    void set_symm_kind(SymmKind k) { spand_set_symm_kind(h_, (int)k); }
    void set_device(int d) { spand_set_device(h_, d); }
    // dim x N column-major (the reference borrows an Eigen::MatrixXd*, include/tree.h:56; here the data is copied)
    void set_Xcoo(int dim, int N, const double* X) { check(spand_set_coords(h_, dim, N, X)); }

    std::vector<ClusterID6> partition(const CscView& A) {
        check(spand_partition(h_, A.rows, A.colptr, A.rowind));
        int N = A.rows;
        std::vector<int> a(N), b(N), c(N), d(N), e(N), f(N);
        spand_get_partition(h_, a.data(), b.data(), c.data(), d.data(), e.data(), f.data());
        std::vector<ClusterID6> out(N);
        for (int i = 0; i < N; i++) out[i] = {a[i], b[i], c[i], d[i], e[i], f[i]};
        return out;
    }
    void assemble(const CscView& A) { check(spand_assemble(h_, A.rows, A.colptr, A.rowind, A.val)); }
    void factorize() { check(spand_factorize(h_)); }  // throws "Error: Non-SPD Pivot\n" / "Error: Singular Pivot\n"
    void solve(double* x) const { check(spand_solve(h_, x)); }
    void solve(std::vector<double>& x) const { solve(x.data()); }

    int get_N() const { return spand_get_N(h_); }
    int get_nlevels() const { return spand_get_nlevels(h_); }
    int get_stop() const { return spand_get_stop(h_); }
    long long nnz() const { return spand_nnz(h_); }
    std::vector<int> get_assembly_perm() const {
        std::vector<int> p(get_N());
        spand_get_perm(h_, p.data());
        return p;
    }
    double factorize_seconds() const { return spand_factorize_seconds(h_); }
    spand_tree* handle() const { return h_; }

    // Tree::get_trailing_mat (include/tree.h:153): permuted ordering, CSC
    struct Csc {
        int n = 0;
        std::vector<int> colptr, rowind;
        std::vector<double> val;
    };
    Csc get_trailing_mat() const {
        Csc T;
        T.n = get_N();
        const int nnz = spand_trailing(h_, nullptr, nullptr, nullptr);
        if (nnz < 0) throw std::runtime_error(spand_last_error(h_));
        T.colptr.assign(T.n + 1, 0);
        T.rowind.assign(nnz, 0);
        T.val.assign(nnz, 0.0);
        if (spand_trailing(h_, T.colptr.data(), T.rowind.data(), T.val.data()) < 0)
            throw std::runtime_error(spand_last_error(h_));
        return T;
    }
    // Tree::log (include/tree.h:163, include/util.h:366-401): one row of spand_log_fields() doubles per level
    std::vector<std::vector<double>> log() const {
        const int nf = spand_log_fields(), nl = get_nlevels();
        std::vector<double> flat((size_t)nf * nl);
        spand_get_log(h_, flat.data());
        std::vector<std::vector<double>> out(nl, std::vector<double>(nf));
        for (int l = 0; l < nl; l++)
            for (int f = 0; f < nf; f++) out[l][f] = flat[(size_t)l * nf + f];
        return out;
    }
    static std::vector<std::string> log_field_names() {
        std::vector<std::string> n;
        for (int f = 0; f < spand_log_fields(); f++) n.push_back(spand_log_field_name(f));
        return n;
    }
    // Tree::print_log (include/tree.h:164): the per-level table of the reference's verbose mode
    void print_log(FILE* out = stdout) const {
        const auto names = log_field_names();
        const auto lg = log();
        std::fprintf(out, "lvl");
        for (const auto& n : names) std::fprintf(out, " %s", n.c_str());
        std::fprintf(out, "\n");
        for (size_t l = 0; l < lg.size(); l++) {
            std::fprintf(out, "%zu", l);
            for (double v : lg[l]) std::fprintf(out, " %.6g", v);
            std::fprintf(out, "\n");
        }
    }
    // set_monitor_flops + write_log_flops (include/tree.h:147, src/tree.cpp:60-77): lvl;kind;rows;cols;inner;time
    void set_monitor_flops(bool on) { spand_set_monitor_flops(h_, on); }
    void write_log_flops(const std::string& fn) const {
        const long long n = spand_get_flops_log(h_, nullptr);
        if (n < 0) throw std::runtime_error(spand_last_error(h_));
        std::vector<long long> t((size_t)5 * n);
        if (n) spand_get_flops_log(h_, t.data());
        static const char* kind[4] = {"pivot", "panel", "gemm", "rrqr"};
        FILE* f = std::fopen(fn.c_str(), "w");
        if (!f) throw std::runtime_error("write_log_flops: cannot open " + fn);
        for (long long i = 0; i < n; i++)
            std::fprintf(f, "%lld;%s;%lld;%lld;%lld;0\n", t[5 * i], kind[t[5 * i + 1] & 3], t[5 * i + 2], t[5 * i + 3],
                         t[5 * i + 4]);
        std::fclose(f);
    }

#ifdef SPAND_B200_HAVE_EIGEN
    using SpMat = Eigen::SparseMatrix<double, 0, int>;
    static CscView view(const SpMat& A) {
        return {(int)A.rows(), (int)A.cols(), A.outerIndexPtr(), A.innerIndexPtr(), A.valuePtr()};
    }
    void set_Xcoo(Eigen::MatrixXd* X) { set_Xcoo((int)X->rows(), (int)X->cols(), X->data()); }
    std::vector<ClusterID6> partition(SpMat& A) { A.makeCompressed(); return partition(view(A)); }
    void assemble(SpMat& A) { A.makeCompressed(); assemble(view(A)); }
    void solve(Eigen::VectorXd& x) const { solve(x.data()); }
#endif

   private:
    spand_tree* h_;
    void check(int rc) const {
        if (rc != 0) throw std::runtime_error(spand_last_error(h_));
    }
};

// include/is.h:12 — returns the reference's iteration count (i + 1)
inline int cg(const CscView& A, const double* rhs, double* x, const Tree& precond, int iters, double tol, bool verb,
              double* seconds = nullptr) {
    int it = spand_cg(precond.handle(), A.rows, A.colptr, A.rowind, A.val, rhs, x, iters, tol, verb, seconds);
    if (it < 0) throw std::runtime_error(spand_last_error(precond.handle()));
    return it;
}
// include/is.h:13 — Householder GMRES, left preconditioned; returns the reference's iteration count
inline int gmres(const CscView& A, const double* rhs, double* x, const Tree& precond, int iters, int restart,
                 double tol, bool verb, double* seconds = nullptr) {
    int it = spand_gmres(precond.handle(), A.rows, A.colptr, A.rowind, A.val, rhs, x, iters, restart, tol, verb, seconds);
    if (it < 0) throw std::runtime_error(spand_last_error(precond.handle()));
    return it;
}
#ifdef SPAND_B200_HAVE_EIGEN
inline int gmres(const Tree::SpMat& A, const Eigen::VectorXd& rhs, Eigen::VectorXd& x, const Tree& precond, int iters,
                 int restart, double tol, bool verb) {
    return gmres(Tree::view(A), rhs.data(), x.data(), precond, iters, restart, tol, verb);
}
inline int cg(const Tree::SpMat& A, const Eigen::VectorXd& rhs, Eigen::VectorXd& x, const Tree& precond, int iters,
              double tol, bool verb) {
    return cg(Tree::view(A), rhs.data(), x.data(), precond, iters, tol, verb);
}
#endif

}  // namespace spaND_b200
