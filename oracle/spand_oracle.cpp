// ORACLE — TEST INFRASTRUCTURE ONLY (see spand_oracle.hpp for scope and pinning).
// Every function cites the reference file:line (relative to /root/reference) that it restates.
#include "spand_oracle.hpp"

#include <sys/time.h>

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <set>
#include <stdexcept>

// OpenBLAS (LP64) from the scipy wheel; symbols carry a scipy_ prefix.
extern "C" {
void scipy_cblas_dgemm(int order, int ta, int tb, int m, int n, int k, double alpha, const double* A, int lda,
                       const double* B, int ldb, double beta, double* C, int ldc);
void scipy_cblas_dsyrk(int order, int uplo, int trans, int n, int k, double alpha, const double* A, int lda,
                       double beta, double* C, int ldc);
void scipy_cblas_dtrsm(int order, int side, int uplo, int trans, int diag, int m, int n, double alpha,
                       const double* A, int lda, double* B, int ldb);
void scipy_cblas_dtrsv(int order, int uplo, int trans, int diag, int n, const double* A, int lda, double* x, int incx);
void scipy_cblas_dgemv(int order, int trans, int m, int n, double alpha, const double* A, int lda, const double* x,
                       int incx, double beta, double* y, int incy);
int scipy_LAPACKE_dpotrf(int layout, char uplo, int n, double* a, int lda);
int scipy_LAPACKE_dgetrf(int layout, int m, int n, double* a, int lda, int* ipiv);
int scipy_LAPACKE_dgeqp3(int layout, int m, int n, double* a, int lda, int* jpvt, double* tau);
int scipy_LAPACKE_dormqr(int layout, char side, char trans, int m, int n, int k, const double* a, int lda,
                         const double* tau, double* c, int ldc);
void scipy_openblas_set_num_threads(int n);
}

namespace spand_oracle {

namespace {
enum { ColMajor = 102, NoTrans = 111, Trans = 112, Upper = 121, Lower = 122, NonUnit = 131, Left = 141, Right = 142 };
double wtime() {
    timeval t;
    gettimeofday(&t, nullptr);
    return t.tv_sec + 1e-6 * t.tv_usec;
}

// src/util.cpp:122-138
void gemm(const DenseMat& A, const DenseMat& B, DenseMat& C, int tA, int tB, double alpha, double beta) {
    int m = C.rows, n = C.cols;
    int k = (tA == NoTrans ? A.cols : A.rows);
    if (m == 0 || n == 0 || k == 0) return;
    scipy_cblas_dgemm(ColMajor, tA, tB, m, n, k, alpha, A.a.data(), A.rows, B.a.data(), B.rows, beta, C.a.data(), C.rows);
}
// src/util.cpp:148-156
void syrk(const DenseMat& A, DenseMat& C, int tA, double alpha, double beta) {
    int n = C.rows;
    int k = (tA == NoTrans ? A.cols : A.rows);
    if (n == 0 || k == 0) return;
    scipy_cblas_dsyrk(ColMajor, Lower, tA, n, k, alpha, A.a.data(), A.rows, beta, C.a.data(), n);
}
// src/util.cpp:158-165
int potf(DenseMat& A) {
    if (A.rows == 0) return 0;
    return scipy_LAPACKE_dpotrf(ColMajor, 'L', A.rows, A.a.data(), A.rows);
}
// src/util.cpp:256-264
void trsm_right(const DenseMat& L, DenseMat& B, int uplo, int trans) {
    if (B.rows == 0 || B.cols == 0) return;
    scipy_cblas_dtrsm(ColMajor, Right, uplo, trans, NonUnit, B.rows, B.cols, 1.0, L.a.data(), B.cols, B.a.data(), B.rows);
}
// src/util.cpp:266-274
void trsm_left(const DenseMat& L, DenseMat& B, int uplo, int trans) {
    if (B.rows == 0 || B.cols == 0) return;
    scipy_cblas_dtrsm(ColMajor, Left, uplo, trans, NonUnit, B.rows, B.cols, 1.0, L.a.data(), B.rows, B.a.data(), B.rows);
}
// src/util.cpp:276-283
void trsv(const DenseMat& LU, double* x, int uplo, int trans) {
    if (LU.rows == 0) return;
    scipy_cblas_dtrsv(ColMajor, uplo, trans, NonUnit, LU.rows, LU.a.data(), LU.rows, x, 1);
}
// src/util.cpp:307-316 : x2 -= A21 x1
void gemv_notrans(const DenseMat& A, const double* x1, double* x2) {
    if (A.rows == 0 || A.cols == 0) return;
    scipy_cblas_dgemv(ColMajor, NoTrans, A.rows, A.cols, -1.0, A.a.data(), A.rows, x1, 1, 1.0, x2, 1);
}
// src/util.cpp:318-326 : x2 -= A12^T x1
void gemv_trans(const DenseMat& A, const double* x1, double* x2) {
    if (A.rows == 0 || A.cols == 0) return;
    scipy_cblas_dgemv(ColMajor, Trans, A.rows, A.cols, -1.0, A.a.data(), A.rows, x1, 1, 1.0, x2, 1);
}
// src/util.cpp:183-200
int getf(DenseMat& A, std::vector<int>& p) {
    int n = A.rows;
    p.resize(n);
    if (n == 0) return 0;
    std::vector<int> swap(n);
    int info = scipy_LAPACKE_dgetrf(ColMajor, n, n, A.a.data(), n, swap.data());
    if (info != 0) return info;
    for (int i = 0; i < n; i++) swap[i] -= 1;
    swap2perm(swap, p);
    return 0;
}
// src/util.cpp:213-227
void split_LU(const DenseMat& A, DenseMat& L, DenseMat& U) {
    int n = A.rows;
    L = DenseMat(n, n);
    U = DenseMat(n, n);
    for (int j = 0; j < n; j++) {
        double dj = A(j, j);
        double sj = std::sqrt(std::fabs(dj));
        for (int i = j + 1; i < n; i++) L(i, j) = A(i, j) * sj;
        L(j, j) = sj;
    }
    for (int i = 0; i < n; i++) {
        double di = A(i, i);
        double si = (di > 0 ? 1.0 : (di < 0 ? -1.0 : 0.0)) * std::sqrt(std::fabs(di));
        double inv = 1.0 / di;
        for (int j = i + 1; j < n; j++) U(i, j) = si * (inv * A(i, j));
        U(i, i) = si;
    }
}
}  // namespace

void set_blas_threads(int n) { scipy_openblas_set_num_threads(n); }

int choose_rank(const double* s, int n, double tol) {
    if (tol == 0) return n;
    if (tol >= 1.0) return 0;
    if (n <= 1) return n;
    double sref = std::fabs(s[0]);
    int rank = 1;
    while (rank < n && std::fabs(s[rank]) / sref >= tol) rank++;
    return rank;
}

void swap2perm(const std::vector<int>& swap, std::vector<int>& perm) {
    int n = (int)swap.size();
    perm.resize(n);
    for (int i = 0; i < n; i++) perm[i] = i;
    for (int i = 0; i < n; i++) std::swap(perm[swap[i]], perm[i]);
}

void block2dense(const std::vector<int>& rowval, const std::vector<int>& colptr, const std::vector<double>& nnzval,
                 int i, int j, int li, int lj, DenseMat* dst, bool transpose) {
    for (int col = 0; col < lj; col++) {
        int b = colptr[j + col], e = colptr[j + col + 1];
        auto it = std::lower_bound(rowval.begin() + b, rowval.begin() + e, i);
        for (int k = (int)(it - rowval.begin()); k < e; k++) {
            int row = rowval[k];
            if (row >= i + li) break;
            if (transpose) (*dst)(col, row - i) = nnzval[k];
            else (*dst)(row - i, col) = nnzval[k];
        }
    }
}

// src/operations.cpp:23-25,105-107,122-124,159-161,180-182,206-208
long long OOp::nnz() const {
    switch (kind) {
        case ScalingLLT: return (long long)nself * (nself + 1) / 2;
        case ScalingPLU: return (long long)nself * nself + 2LL * nself;
        case GemmSymmOut:
        case GemmSymmIn:
        case GemmOut:
        case GemmIn: return (long long)nself * nnbr;
        case Orthogonal: return (long long)nself * nself;
        default: return 0;
    }
}

// src/tree.cpp:86-111
OTree::OTree(int nlevels_) : nlevels(nlevels_) {
    if (nlevels <= 0) throw std::runtime_error("nlevels must be > 0");
    log.assign(nlevels, LevelLog());
    bottoms.resize(nlevels);
}

void OTree::set_coords(int dim, int N_, const double* X) {
    Xcoo = DenseMat(dim, N_);
    std::copy(X, X + (size_t)dim * N_, Xcoo.a.begin());
    have_coords = true;
}

// src/cluster.cpp:113-123
void OTree::add_edge(OCluster* c, std::unique_ptr<OEdge> e) {
    if (e->n1 == e->n2) {
        c->out.push_front(std::move(e));
    } else {
        e->n2->in.push_back(e.get());
        c->out.push_back(std::move(e));
    }
}

// src/cluster.cpp:162-176
void OTree::sort_edges(OCluster* c) {
    c->out.sort([](const std::unique_ptr<OEdge>& a, const std::unique_ptr<OEdge>& b) {
        if (a->n1 == a->n2) return true;
        if (b->n1 == b->n2) return false;
        return a->n2->order < b->n2->order;
    });
    c->in.sort([](const OEdge* a, const OEdge* b) { return a->n1->order < b->n1->order; });
}

// src/cluster.cpp:32-45
void OTree::set_eliminated(OCluster* s) {
    for (auto& e : s->out) e->n2->in.remove_if([s](OEdge* e2) { return e2->n1 == s; });
    for (auto* e : s->in) e->n1->out.remove_if([s](const std::unique_ptr<OEdge>& e2) { return e2->n2 == s; });
    s->in.clear();
    s->out.clear();
    s->eliminated = true;
}

// src/tree.cpp:306-419 (the integer part lives in build_ordering, shared with the product)
void OTree::partition(const SpMat& A) {
    if (use_geo && !have_coords) throw std::runtime_error("use_geo set without coordinates");
    ord = spand::build_ordering(A, nlevels, use_geo ? &Xcoo : nullptr, verb);
    N = A.rows;
    for (auto& b : bottoms) b.clear();
    others.clear();
    ops.clear();
    log.assign(nlevels, LevelLog());
    for (auto& p : ord.part) log[p.self.lvl].dofs_nd += 1;
    for (int l = nlevels - 2; l >= 0; l--) log[l].dofs_left_nd = log[l + 1].dofs_nd + log[l + 1].dofs_left_nd;
    std::vector<std::vector<OCluster*>> ptr(nlevels);
    for (int h = 0; h < nlevels; h++) {
        for (auto& cn : ord.levels[h]) {
            auto c = std::make_unique<OCluster>();
            c->start = cn.start;
            c->size = cn.size;
            c->level = cn.level;
            c->order = cn.order;
            c->sparsify = cn.sparsify;
            c->x.assign(cn.size, 0.0);
            ptr[h].push_back(c.get());
            if (h > 0)
                for (int k = cn.child_begin; k < cn.child_end; k++) {
                    ptr[h - 1][k]->parent = c.get();
                    c->children.push_back(ptr[h - 1][k]);
                }
            bottoms[h].push_back(std::move(c));
        }
    }
    max_order = ord.norders;
    ilvl = 0;
    current_bottom = 0;
}

// src/tree.cpp:505-575
void OTree::assemble(const SpMat& A) {
    SpMat App = spand::symm_perm(A, ord.perm);
    std::vector<OCluster*> cmap(N);
    for (auto& self : bottoms[0])
        for (int k = self->start; k < self->start + self->size; k++) cmap[k] = self.get();
    for (auto& self : bottoms[0]) {
        // neighbours in this block column, de-duplicated, in increasing order
        std::map<int, OCluster*> nbrs;
        for (int j = self->start; j < self->start + self->size; j++)
            for (int k = App.colptr[j]; k < App.colptr[j + 1]; k++) {
                int row = App.rowind[k];
                if (symmetry() && row < j) continue;
                nbrs[cmap[row]->order] = cmap[row];
            }
        nbrs[self->order] = self.get();
        for (auto& kv : nbrs) {
            OCluster* nbr = kv.second;
            auto e = std::make_unique<OEdge>();
            e->n1 = self.get();
            e->n2 = nbr;
            e->original = true;
            e->A = DenseMat(nbr->size, self->size);
            block2dense(App.rowind, App.colptr, App.val, nbr->start, self->start, nbr->size, self->size, &e->A, false);
            add_edge(self.get(), std::move(e));
        }
        sort_edges(self.get());
    }
}

int OTree::ndofs_left() const {
    int n = 0;
    for (auto& s : bottoms[current_bottom])
        if (!s->eliminated) n += s->size;
    return n;
}

// src/tree.cpp:748-793 (diag == nullptr branch)
void OTree::schur_update(OEdge* e1, OEdge* e2, bool t1, bool t2) {
    OCluster* s = t1 ? e1->n2 : e1->n1;
    OCluster* n1 = t1 ? e1->n1 : e1->n2;
    OCluster* n2 = t2 ? e2->n2 : e2->n1;
    OEdge* target = nullptr;
    for (auto& e : n2->out)
        if (e->n2 == n1) {
            target = e.get();
            break;
        }
    if (!target) {  // fill-in
        auto e = std::make_unique<OEdge>();
        e->n1 = n2;
        e->n2 = n1;
        e->original = false;
        e->A = DenseMat(n1->size, n2->size);
        target = e.get();
        add_edge(n2, std::move(e));
    }
    LevelLog& lg = log[ilvl];
    if (n1 == n2 && symmetry()) {
        syrk(e1->A, target->A, t1 ? Trans : NoTrans, -1.0, 1.0);
        lg.fl_schur += (double)n1->size * (n1->size + 1) * s->size;
        flop(2, n1->size, n2->size, s->size);
    } else {
        gemm(e1->A, e2->A, target->A, t1 ? Trans : NoTrans, t2 ? Trans : NoTrans, -1.0, 1.0);
        lg.fl_schur += 2.0 * n1->size * n2->size * s->size;
        flop(2, n1->size, n2->size, s->size);
    }
}

// src/tree.cpp:895-967 (LLT :897-910 and PLU :929-956 branches)
void OTree::eliminate_cluster(OCluster* self) {
    LevelLog& lg = log[ilvl];
    double n = self->size;
    if (scale_kind == LLT) {
        DenseMat& Ass = self->pivot()->A;
        if (potf(Ass) != 0) throw std::runtime_error("Error: Non-SPD Pivot\n");  // tree.cpp:587-590
        lg.fl_pivot += n * n * n / 3.0;
        flop(0, (long)n, 0, 0);
        for (auto* e : self->in) {  // tree.cpp:717-724
            trsm_left(Ass, e->A, Lower, NoTrans);
            lg.fl_panel += (double)e->A.cols * n * n;
            flop(1, e->A.cols, (long)n, 0);
        }
        bool first = true;
        for (auto& e : self->out) {
            if (first) { first = false; continue; }
            trsm_right(Ass, e->A, Lower, Trans);
            lg.fl_panel += (double)e->A.rows * n * n;
            flop(1, e->A.rows, (long)n, 0);
        }
        // tree.cpp:862-883
        std::vector<OEdge*> outs;
        for (auto& e : self->out)
            if (e->n1 != e->n2) outs.push_back(e.get());
        std::vector<OEdge*> ins(self->in.begin(), self->in.end());
        for (auto* e1 : outs)
            for (auto* e2 : outs)
                if (e1->n2->order >= e2->n2->order) schur_update(e1, e2, false, true);
        for (auto* e1 : ins)
            for (auto* e2 : ins)
                if (e1->n1->order >= e2->n1->order) schur_update(e1, e2, true, false);
        for (auto* e1 : outs)
            for (auto* e2 : ins) schur_update(e1, e2, false, false);
        // record (tree.cpp:909-910, :885-892)
        OOp op;
        op.kind = OOp::ScalingLLT;
        op.self = self;
        op.nself = self->size;
        op.A = std::move(self->pivot()->A);
        ops.push_back(std::move(op));
        for (auto* e : outs) {
            OOp g;
            g.kind = OOp::GemmSymmOut;
            g.self = self;
            g.nbr = e->n2;
            g.nself = self->size;
            g.nnbr = e->n2->size;
            g.A = std::move(e->A);
            ops.push_back(std::move(g));
        }
        for (auto* e : ins) {
            OOp g;
            g.kind = OOp::GemmSymmIn;
            g.self = self;
            g.nbr = e->n1;
            g.nself = self->size;
            g.nnbr = e->n1->size;
            g.A = std::move(e->A);
            ops.push_back(std::move(g));
        }
    } else {  // PLU
        OOp op;
        op.kind = OOp::ScalingPLU;
        op.self = self;
        op.nself = self->size;
        DenseMat& Ass = self->pivot()->A;
        if (getf(Ass, op.p) != 0) throw std::runtime_error("Error: Singular Pivot\n");  // tree.cpp:625-628
        op.q.resize(self->size);
        for (int i = 0; i < self->size; i++) op.q[i] = i;
        split_LU(Ass, op.A, op.U);
        lg.fl_pivot += 2.0 * n * n * n / 3.0;
        flop(0, (long)n, 0, 0);
        std::vector<OEdge*> outs;
        for (auto& e : self->out)
            if (e->n1 != e->n2) outs.push_back(e.get());
        std::vector<OEdge*> ins(self->in.begin(), self->in.end());
        for (auto* e : ins) {  // tree.cpp:668-676 : L^-1 P^T Asn
            DenseMat tmp = e->A;
            for (int j = 0; j < tmp.cols; j++)
                for (int i = 0; i < tmp.rows; i++) e->A(i, j) = tmp(op.p[i], j);
            trsm_left(op.A, e->A, Lower, NoTrans);
            lg.fl_panel += (double)e->A.cols * n * n;
            flop(1, e->A.cols, (long)n, 0);
        }
        for (auto* e : outs) {  // tree.cpp:681-689 : Ans Q^T U^-1 (Q = I for PLU)
            trsm_right(op.U, e->A, Upper, NoTrans);
            lg.fl_panel += (double)e->A.rows * n * n;
            flop(1, e->A.rows, (long)n, 0);
        }
        for (auto* e1 : outs)
            for (auto* e2 : ins) schur_update(e1, e2, false, false);  // tree.cpp:943-947
        ops.push_back(std::move(op));
        for (auto* e : outs) {
            OOp g;
            g.kind = OOp::GemmOut;
            g.self = self;
            g.nbr = e->n2;
            g.nself = self->size;
            g.nnbr = e->n2->size;
            g.A = std::move(e->A);
            ops.push_back(std::move(g));
        }
        for (auto* e : ins) {
            OOp g;
            g.kind = OOp::GemmIn;
            g.self = self;
            g.nbr = e->n1;
            g.nself = self->size;
            g.nnbr = e->n1->size;
            g.A = std::move(e->A);
            ops.push_back(std::move(g));
        }
    }
    set_eliminated(self);
}

// src/tree.cpp:796-856
void OTree::scale_cluster(OCluster* self) {
    LevelLog& lg = log[ilvl];
    double n = self->size;
    lg.by_scale += 16.0 * n * n;  // pivot read + L write
    if (scale_kind == LLT) {
        DenseMat& Ass = self->pivot()->A;
        if (potf(Ass) != 0) throw std::runtime_error("Error: Non-SPD Pivot\n");
        lg.fl_pivot += n * n * n / 3.0;
        flop(0, (long)n, 0, 0);
        for (auto* e : self->in) {
            trsm_left(Ass, e->A, Lower, NoTrans);
            lg.fl_panel += (double)e->A.cols * n * n;
            flop(1, e->A.cols, (long)n, 0);
            lg.by_scale += 8.0 * e->A.rows * e->A.cols;  // each block is visited from both ends: 2 x 8 = 16 B/elem
        }
        bool first = true;
        for (auto& e : self->out) {
            if (first) { first = false; continue; }
            trsm_right(Ass, e->A, Lower, Trans);
            lg.fl_panel += (double)e->A.rows * n * n;
            flop(1, e->A.rows, (long)n, 0);
            lg.by_scale += 8.0 * e->A.rows * e->A.cols;
        }
        OOp op;
        op.kind = OOp::ScalingLLT;
        op.self = self;
        op.nself = self->size;
        op.A = std::move(self->pivot()->A);
        ops.push_back(std::move(op));
    } else {
        OOp op;
        op.kind = OOp::ScalingPLU;
        op.self = self;
        op.nself = self->size;
        DenseMat& Ass = self->pivot()->A;
        if (getf(Ass, op.p) != 0) throw std::runtime_error("Error: Singular Pivot\n");
        op.q.resize(self->size);
        for (int i = 0; i < self->size; i++) op.q[i] = i;
        split_LU(Ass, op.A, op.U);
        lg.fl_pivot += 2.0 * n * n * n / 3.0;
        flop(0, (long)n, 0, 0);
        for (auto* e : self->in) {
            DenseMat tmp = e->A;
            for (int j = 0; j < tmp.cols; j++)
                for (int i = 0; i < tmp.rows; i++) e->A(i, j) = tmp(op.p[i], j);
            trsm_left(op.A, e->A, Lower, NoTrans);
            lg.fl_panel += (double)e->A.cols * n * n;
            flop(1, e->A.cols, (long)n, 0);
            lg.by_scale += 8.0 * e->A.rows * e->A.cols;
        }
        bool first = true;
        for (auto& e : self->out) {
            if (first) { first = false; continue; }
            trsm_right(op.U, e->A, Upper, NoTrans);
            lg.fl_panel += (double)e->A.rows * n * n;
            flop(1, e->A.rows, (long)n, 0);
            lg.by_scale += 8.0 * e->A.rows * e->A.cols;
        }
        ops.push_back(std::move(op));
    }
    // pivot := I (tree.cpp:811,848,853)
    DenseMat I(self->size, self->size);
    for (int i = 0; i < self->size; i++) I(i, i) = 1.0;
    self->pivot()->A = std::move(I);
}

// src/tree.cpp:1417-1433 -> sparsify_adaptive_only :1292-1347 (pred == true, pivot == I)
//   with assemble_Asn :1189-1224 and shrink_split_scatter_phi :980-1104
void OTree::sparsify_cluster(OCluster* self) {
    LevelLog& lg = log[ilvl];
    bool want = use_want_sparsify ? self->sparsify : true;  // tree.cpp:287-294
    if (!want) {
        lg.ignored++;
        return;
    }
    lg.rank_before += self->size;
    lg.nspars++;
    int rows = self->size;
    int cols = 0;
    for (auto* e : self->in) cols += e->n1->size;
    bool first = true;
    for (auto& e : self->out) {
        if (first) { first = false; continue; }
        cols += e->n2->size;
    }
    lg.nbrs += cols;
    DenseMat Asn(rows, cols);
    int c = 0;
    for (auto* e : self->in) {
        int w = e->n1->size;
        for (int j = 0; j < w; j++)
            for (int i = 0; i < rows; i++) Asn(i, c + j) = e->A(i, j);
        c += w;
    }
    first = true;
    for (auto& e : self->out) {
        if (first) { first = false; continue; }
        int w = e->n2->size;
        for (int j = 0; j < w; j++)
            for (int i = 0; i < rows; i++) Asn(i, c + j) = e->A(j, i);
        c += w;
    }
    int mn = std::min(rows, cols);
    std::vector<int> jpvt(cols, 0);
    std::vector<double> tau(mn, 0.0);
    if (rows > 0 && cols > 0) {  // src/util.cpp:383-394
        int info = scipy_LAPACKE_dgeqp3(ColMajor, rows, cols, Asn.a.data(), rows, jpvt.data(), tau.data());
        if (info != 0) throw std::runtime_error("dgeqp3 failed");
        for (auto& j : jpvt) j--;
    }
    std::vector<double> diag(mn);
    for (int i = 0; i < mn; i++) diag[i] = Asn(i, i);
    int rank = choose_rank(diag.data(), mn, tol);
    if (const char* qlog = getenv("SPAND_ORACLE_QRLOG")) {
        // analysis hook (not part of the restatement): one line per RRQR "level rows cols rank" for traffic models
        static FILE* qf = fopen(qlog, "w");
        if (qf) { fprintf(qf, "%d %d %d %d\n", ilvl, rows, cols, rank); fflush(qf); }
    }
    if (getenv("SPAND_ORACLE_DEADCOLS") && rank > 0) {
        // analysis hook (not part of the restatement): how much of the per-step panel sweep of a truncated QRCP
        // touches columns whose residual norm is already below the stopping threshold (they can never be pivots)
        const double margin = atof(getenv("SPAND_ORACLE_DEADCOLS"));
        const double thr2 = margin * tol * diag[0] * margin * tol * diag[0];
        std::vector<double> res2(cols, 0.0);  // residual norm^2 of column p below row k, built bottom-up
        std::vector<std::vector<double>> tail(cols);
        double all = 0, live = 0, live_blk = 0;
        for (int p = 0; p < cols; p++) {
            tail[p].assign(mn + 1, 0.0);
            for (int k = std::min(mn, p + 1) - 1; k >= 0; k--) tail[p][k] = tail[p][k + 1] + Asn(k, p) * Asn(k, p);
        }
        for (int k = 0; k < rank; k++) {
            const int kb = k - k % 16;  // liveness refreshed at block boundaries only
            for (int p = k + 1; p < cols; p++) {
                all += rows - k;
                if (tail[p][std::min(k, mn)] >= thr2) live += rows - k;
                if (tail[p][std::min(kb, mn)] >= thr2) live_blk += rows - k;
            }
        }
        if (getenv("SPAND_ORACLE_HOTCOLS")) {
            // hot/cold model: at a block boundary kb a column is "hot" when its residual norm is at least theta times
            // the pivot level |R(kb,kb)|; only hot columns are swept during the block
            const int nbh = atoi(getenv("SPAND_ORACLE_HOTCOLS"));
            static double h_all[64], h_hot[3][64];
            const double th[3] = {0.25, 0.5, 0.75};
            for (int k = 0; k < rank; k++) {
                const int kb = k - k % nbh;
                const double lev2 = Asn(kb, kb) * Asn(kb, kb);
                for (int p = k + 1; p < cols; p++) {
                    h_all[ilvl] += rows - k;
                    for (int q = 0; q < 3; q++)
                        if (tail[p][std::min(kb, mn)] >= th[q] * th[q] * lev2) h_hot[q][ilvl] += rows - k;
                }
            }
            fprintf(stderr, "HOTCOLS lvl %d rows %d cols %d rank %d | level so far hot(.25) %.3f hot(.5) %.3f hot(.75) %.3f (%.3g)\n",
                    ilvl, rows, cols, rank, h_hot[0][ilvl] / h_all[ilvl], h_hot[1][ilvl] / h_all[ilvl],
                    h_hot[2][ilvl] / h_all[ilvl], h_all[ilvl]);
        }
        static double g_all[64], g_live[64], g_blk[64];
        g_all[ilvl] += all; g_live[ilvl] += live; g_blk[ilvl] += live_blk;
        fprintf(stderr, "DEADCOLS lvl %d rows %d cols %d rank %d live %.3f live_blk %.3f | level so far %.3f %.3f (%.3g)\n", ilvl, rows, cols,
                rank, all > 0 ? live / all : 1.0, all > 0 ? live_blk / all : 1.0, g_live[ilvl] / g_all[ilvl], g_blk[ilvl] / g_all[ilvl], g_all[ilvl]);
    }
    {
        double r = rows, cc = cols, rf = mn, rk = rank;
        flop(3, rows, cols, 0);
        lg.fl_rrqr_full += 4 * r * cc * rf - 2 * (r + cc) * rf * rf + (4.0 / 3.0) * rf * rf * rf;
        lg.fl_rrqr_rank += 4 * r * cc * rk - 2 * (r + cc) * rk * rk + (4.0 / 3.0) * rk * rk * rk;
        lg.by_rrqr += 8 * r * cc + 8 * rk * cc + 8 * r * rk;
    }
    if (rank >= rows) {  // tree.cpp:1317-1319
        lg.rank_after += self->size;
        return;
    }
    // Orthogonal op: v = Asn[:, :rank], h = tau[:rank]; captured segment length = rows (tree.cpp:1322-1331)
    OOp q;
    q.kind = OOp::Orthogonal;
    q.self = self;
    q.nself = rows;
    q.A = DenseMat(rows, rank);
    for (int j = 0; j < rank; j++)
        for (int i = 0; i < rows; i++) q.A(i, j) = Asn(i, j);
    q.tau.assign(tau.begin(), tau.begin() + rank);
    ops.push_back(std::move(q));
    // AsnP = triu(Asn[:rank,:]) * P^T  (tree.cpp:1334-1335)
    DenseMat AsnP(rank, cols);
    for (int j = 0; j < cols; j++)
        for (int i = 0; i < rank && i <= j; i++) AsnP(i, jpvt[j]) = Asn(i, j);
    // shrink + scatter (tree.cpp:980-1067, pred == true branches)
    int sibling_size = rows - rank;
    auto sib = std::make_unique<OCluster>();
    sib->start = self->start + rank;
    sib->size = sibling_size;
    sib->level = self->level;
    sib->order = max_order++;
    sib->sparsify = self->sparsify;
    sib->x.assign(sibling_size, 0.0);
    self->size = rank;
    c = 0;
    for (auto* e : self->in) {
        int w = e->n1->size;
        DenseMat B(rank, w);
        for (int j = 0; j < w; j++)
            for (int i = 0; i < rank; i++) B(i, j) = AsnP(i, c + j);
        e->A = std::move(B);
        c += w;
    }
    first = true;
    for (auto& e : self->out) {
        if (first) { first = false; continue; }
        int w = e->n2->size;
        DenseMat B(w, rank);
        for (int j = 0; j < w; j++)
            for (int i = 0; i < rank; i++) B(j, i) = AsnP(i, c + j);
        e->A = std::move(B);
        c += w;
    }
    // pivot I -> I_rank (+) I_rest (tree.cpp:1076-1099); sibling owns only its identity pivot
    {
        DenseMat Aoo(rank, rank);
        for (int i = 0; i < rank; i++) Aoo(i, i) = 1.0;
        self->pivot()->A = std::move(Aoo);
        auto e = std::make_unique<OEdge>();
        e->n1 = e->n2 = sib.get();
        e->original = true;
        e->A = DenseMat(sibling_size, sibling_size);
        for (int i = 0; i < sibling_size; i++) e->A(i, i) = 1.0;
        add_edge(sib.get(), std::move(e));
    }
    OCluster* sibling = sib.get();
    others.push_back(std::move(sib));
    OOp sp;
    sp.kind = OOp::Split;
    sp.self = self;
    sp.nbr = sibling;
    ops.push_back(std::move(sp));
    eliminate_cluster(sibling);  // tree.cpp:1342 : factor of I, no edges
    lg.rank_after += self->size;
}

// src/tree.cpp:1435-1445 with reset_size :1106-1131 and update_edges :1133-1184
void OTree::merge_all() {
    LevelLog& lg = log[ilvl];
    current_bottom++;
    std::map<OCluster*, int> posparent;
    for (auto& snew : bottoms[current_bottom]) {
        int size = 0;
        for (auto* sold : snew->children) {
            posparent[sold] = size;
            size += sold->size;
        }
        snew->size = size;
        snew->x.assign(size, 0.0);
        OOp m;
        m.kind = OOp::Merge;
        m.self = snew.get();
        ops.push_back(std::move(m));
    }
    for (auto& snew : bottoms[current_bottom]) {
        std::map<int, std::pair<OCluster*, bool>> merged;  // keyed by order => deterministic
        for (auto* sold : snew->children)
            for (auto& eold : sold->out) {
                OCluster* nnew = eold->n2->parent;
                auto it = merged.find(nnew->order);
                if (it == merged.end()) it = merged.emplace(nnew->order, std::make_pair(nnew, false)).first;
                if (!it->second.second) it->second.second = eold->original;
            }
        for (auto& kv : merged) {
            OCluster* nnew = kv.second.first;
            auto e = std::make_unique<OEdge>();
            e->n1 = snew.get();
            e->n2 = nnew;
            e->original = kv.second.second;
            e->A = DenseMat(nnew->size, snew->size);
            lg.by_merge += 8.0 * nnew->size * snew->size;
            add_edge(snew.get(), std::move(e));
        }
        sort_edges(snew.get());
        for (auto* sold : snew->children) {
            for (auto& eold : sold->out) {
                OCluster* nold = eold->n2;
                OCluster* nnew = nold->parent;
                OEdge* found = nullptr;
                for (auto& e : snew->out)
                    if (e->n2 == nnew) {
                        found = e.get();
                        break;
                    }
                int r0 = posparent[nold], c0 = posparent[sold];
                for (int j = 0; j < sold->size; j++)
                    for (int i = 0; i < nold->size; i++) found->A(r0 + i, c0 + j) = eold->A(i, j);
                lg.by_merge += 8.0 * nold->size * sold->size;
            }
            sold->out.clear();
            sold->in.clear();
        }
    }
}

// src/tree.cpp:1447-1551
void OTree::factorize() {
    flop_log.clear();
    if (symm_kind == SPD && scale_kind != LLT) throw std::runtime_error("SPD requires LLT");
    if (symm_kind == GEN && scale_kind != PLU) throw std::runtime_error("GEN requires PLU");
    if (symm_kind == SYM) throw std::runtime_error("SYM/LDLT is out of scope (SURVEY.md section 2)");
    for (ilvl = 0; ilvl < nlevels; ilvl++) {
        LevelLog& lg = log[ilvl];
        if (verb) printf("Level %d, %d dofs left\n", ilvl, ndofs_left());
        double t0 = wtime();
        for (auto& self : bottoms[current_bottom])
            if (self->level == ilvl) eliminate_cluster(self.get());
        double t1 = wtime();
        lg.t_elim += t1 - t0;
        lg.dofs_left_elim = ndofs_left();
        if (ilvl == stop_level && stop_phase == 0) return;
        if (ilvl >= skip) {
            for (auto& self : bottoms[current_bottom])
                if (self->level > ilvl) scale_cluster(self.get());
            double t2 = wtime();
            lg.t_scale += t2 - t1;
            if (ilvl == stop_level && stop_phase == 1) return;
            for (auto& self : bottoms[current_bottom])
                if (self->level > ilvl) sparsify_cluster(self.get());
            double t3 = wtime();
            lg.t_spars += t3 - t2;
            if (ilvl == stop_level && stop_phase == 2) return;
        }
        if (ilvl < nlevels - 1) {
            double t4 = wtime();
            merge_all();
            lg.t_merge += wtime() - t4;
        }
        lg.dofs_left_spars = ndofs_left();
        lg.fact_nnz = nnz();
        if (ilvl == stop_level && stop_phase == 3) return;
        if (verb)
            printf("  lvl %d: elim %.2e scale %.2e spars %.2e merge %.2e | left %d -> %d\n", ilvl, lg.t_elim, lg.t_scale,
                   lg.t_spars, lg.t_merge, lg.dofs_left_elim, lg.dofs_left_spars);
    }
}

// src/tree.cpp:1637-1643
long long OTree::nnz() const {
    long long n = 0;
    for (auto& op : ops) n += op.nnz();
    return n;
}

// src/tree.cpp:158-169
int OTree::get_stop() const {
    int stop = N;
    for (auto& l : log) {
        if (l.dofs_left_elim > 0) stop = std::min(stop, l.dofs_left_elim);
        if (l.dofs_left_spars > 0) stop = std::min(stop, l.dofs_left_spars);
    }
    return stop;
}

// src/tree.cpp:1610-1635 with src/operations.cpp:31-208 and src/cluster.cpp:82-98
void OTree::solve(double* xio) {
    std::vector<double> b(N);
    for (int i = 0; i < N; i++) b[i] = xio[ord.perm[i]];
    for (int h = 0; h < nlevels; h++)
        for (auto& c : bottoms[h]) {
            if (h == 0) std::copy(b.begin() + c->start, b.begin() + c->start + c->x.size(), c->x.begin());
            else std::fill(c->x.begin(), c->x.end(), 0.0);
        }
    auto fwd = [&](OOp& op) {
        switch (op.kind) {
            case OOp::ScalingLLT: trsv(op.A, op.self->x.data(), Lower, NoTrans); break;
            case OOp::ScalingPLU: {
                std::vector<double> t(op.self->x.begin(), op.self->x.begin() + op.nself);
                for (int i = 0; i < op.nself; i++) op.self->x[i] = t[op.p[i]];
                trsv(op.A, op.self->x.data(), Lower, NoTrans);
                break;
            }
            case OOp::GemmSymmOut:
            case OOp::GemmOut:
                if (op.nself > 0 && op.nnbr > 0) gemv_notrans(op.A, op.self->x.data(), op.nbr->x.data());
                break;
            case OOp::GemmSymmIn:
                if (op.nself > 0 && op.nnbr > 0) gemv_trans(op.A, op.self->x.data(), op.nbr->x.data());
                break;
            case OOp::GemmIn: break;
            case OOp::Orthogonal:
                if (op.nself > 0 && op.A.cols > 0)
                    scipy_LAPACKE_dormqr(ColMajor, 'L', 'T', op.nself, 1, op.A.cols, op.A.a.data(), op.nself,
                                         op.tau.data(), op.self->x.data(), op.nself);
                break;
            case OOp::Merge: {
                int k = 0;
                for (auto* c : op.self->children)
                    for (int i = 0; i < c->size; i++) op.self->x[k++] = c->x[i];
                break;
            }
            case OOp::Split:
                for (int i = 0; i < op.nbr->size; i++) op.nbr->x[i] = op.self->x[op.self->size + i];
                break;
        }
    };
    auto bwd = [&](OOp& op) {
        switch (op.kind) {
            case OOp::ScalingLLT: trsv(op.A, op.self->x.data(), Lower, Trans); break;
            case OOp::ScalingPLU: {
                trsv(op.U, op.self->x.data(), Upper, NoTrans);
                std::vector<double> t(op.self->x.begin(), op.self->x.begin() + op.nself);
                for (int i = 0; i < op.nself; i++) op.self->x[i] = t[op.q[i]];
                break;
            }
            case OOp::GemmSymmOut:
                if (op.nself > 0 && op.nnbr > 0) gemv_trans(op.A, op.nbr->x.data(), op.self->x.data());
                break;
            case OOp::GemmSymmIn:
            case OOp::GemmIn:
                if (op.nself > 0 && op.nnbr > 0) gemv_notrans(op.A, op.nbr->x.data(), op.self->x.data());
                break;
            case OOp::GemmOut: break;
            case OOp::Orthogonal:
                if (op.nself > 0)
                    scipy_LAPACKE_dormqr(ColMajor, 'L', 'N', op.nself, 1, op.A.cols, op.A.a.data(), op.nself,
                                         op.tau.data(), op.self->x.data(), op.nself);
                break;
            case OOp::Merge: {
                int k = 0;
                for (auto* c : op.self->children)
                    for (int i = 0; i < c->size; i++) c->x[i] = op.self->x[k++];
                break;
            }
            case OOp::Split:
                for (int i = 0; i < op.nbr->size; i++) op.self->x[op.self->size + i] = op.nbr->x[i];
                break;
        }
    };
    for (auto it = ops.begin(); it != ops.end(); ++it) fwd(*it);
    for (auto it = ops.rbegin(); it != ops.rend(); ++it) bwd(*it);
    for (auto& c : bottoms[0]) std::copy(c->x.begin(), c->x.end(), b.begin() + c->start);
    for (int i = 0; i < N; i++) xio[ord.perm[i]] = b[i];
}

// src/tree.cpp:1730-1763 — returned in the permuted ordering, full (both triangles for symmetric kinds)
SpMat OTree::trailing_mat() const {
    std::vector<spand::Triplet> t;
    for (auto& s : bottoms[current_bottom]) {
        for (auto& e : s->out) {
            int r0 = e->n2->start, c0 = s->start;
            for (int j = 0; j < s->size; j++)
                for (int i = 0; i < e->n2->size; i++) {
                    int gi = r0 + i, gj = c0 + j;
                    double v = e->A(i, j);
                    if (symmetry()) {
                        if (gi > gj) {
                            t.push_back({gj, gi, v});
                            t.push_back({gi, gj, v});
                        } else if (gi == gj) {
                            t.push_back({gi, gi, v});
                        }
                    } else {
                        t.push_back({gi, gj, v});
                    }
                }
        }
    }
    return spand::from_triplets(N, N, t);
}

void OTree::stats(std::vector<int>& id, std::vector<int>& size, std::vector<int>& rank) const {
    id.clear();
    size.clear();
    rank.clear();
    for (int h = 0; h < nlevels; h++)
        for (auto& c : bottoms[h]) {
            id.push_back(c->order);
            size.push_back(c->original_size());
            rank.push_back(c->size);
        }
}

// src/is.cpp:39-121
int cg(const SpMat& A, const double* rhs, double* x, OTree& precond, int iters, double tol, bool verb) {
    int n = A.cols;
    std::vector<double> residual(n), p(n), z(n), tmp(n);
    spand::spmv(A, x, tmp.data());
    for (int i = 0; i < n; i++) residual[i] = rhs[i] - tmp[i];
    double rhsNorm2 = 0;
    for (int i = 0; i < n; i++) rhsNorm2 += rhs[i] * rhs[i];
    if (rhsNorm2 == 0) {
        std::fill(x, x + n, 0.0);
        return 0;
    }
    double threshold = tol * tol * rhsNorm2;
    double residualNorm2 = 0;
    for (int i = 0; i < n; i++) residualNorm2 += residual[i] * residual[i];
    if (residualNorm2 < threshold) return 0;
    p = residual;
    precond.solve(p.data());
    double absNew = 0;
    for (int i = 0; i < n; i++) absNew += residual[i] * p[i];
    int i = 0;
    while (i < iters) {
        spand::spmv(A, p.data(), tmp.data());
        double ptmp = 0;
        for (int k = 0; k < n; k++) ptmp += p[k] * tmp[k];
        double alpha = absNew / ptmp;
        residualNorm2 = 0;
        for (int k = 0; k < n; k++) {
            x[k] += alpha * p[k];
            residual[k] -= alpha * tmp[k];
            residualNorm2 += residual[k] * residual[k];
        }
        if (verb) printf("%d: |Ax-b|/|b| = %3.2e <? %3.2e\n", i, std::sqrt(residualNorm2 / rhsNorm2), tol);
        if (residualNorm2 < threshold) break;
        z = residual;
        precond.solve(z.data());
        double absOld = absNew;
        absNew = 0;
        for (int k = 0; k < n; k++) absNew += residual[k] * z[k];
        double beta = absNew / absOld;
        for (int k = 0; k < n; k++) p[k] = z[k] + beta * p[k];
        i++;
    }
    return i + 1;
}


// ---- Householder GMRES, src/is.cpp:123-300. The reference leans on Eigen 3.3 for three primitives; they are
// restated here from Eigen's published definitions (Householder.h makeHouseholder / applyHouseholderOnTheLeft,
// Jacobi.h makeGivens for real scalars).
namespace {

// x = [c0; tail] (length n): H x = beta e_1 with H = I - tau [1; ess][1; ess]^T
void make_householder(const double* x, int n, double* ess, double& tau, double& beta) {
    double tail2 = 0;
    for (int i = 1; i < n; i++) tail2 += x[i] * x[i];
    const double c0 = x[0];
    if (n == 1 || tail2 <= std::numeric_limits<double>::min()) {
        tau = 0;
        beta = c0;
        for (int i = 1; i < n; i++) ess[i - 1] = 0;
    } else {
        beta = std::sqrt(c0 * c0 + tail2);
        if (c0 >= 0) beta = -beta;
        for (int i = 1; i < n; i++) ess[i - 1] = x[i] / (c0 - beta);
        tau = (beta - c0) / beta;
    }
}

void apply_householder_left(double* v, int n, const double* ess, double tau) {
    if (n == 1) {
        v[0] *= 1 - tau;
    } else if (tau != 0) {
        double tmp = 0;
        for (int i = 1; i < n; i++) tmp += ess[i - 1] * v[i];
        tmp += v[0];
        v[0] -= tau * tmp;
        for (int i = 1; i < n; i++) v[i] -= tau * ess[i - 1] * tmp;
    }
}

struct Givens {
    double c = 1, s = 0;
    void make(double p, double q) {
        if (q == 0) {
            c = p < 0 ? -1 : 1;
            s = 0;
        } else if (p == 0) {
            c = 0;
            s = q < 0 ? 1 : -1;
        } else if (std::fabs(p) > std::fabs(q)) {
            double t = q / p, u = std::sqrt(1 + t * t);
            if (p < 0) u = -u;
            c = 1 / u;
            s = -t * c;
        } else {
            double t = p / q, u = std::sqrt(1 + t * t);
            if (q < 0) u = -u;
            s = -1 / u;
            c = -t * s;
        }
    }
    // v.applyOnTheLeft(p, q, G.adjoint()): [x_p; x_q] <- [c -s; s c] [x_p; x_q]
    void apply_adjoint(double& xp, double& xq) const {
        double a = c * xp - s * xq, b = s * xp + c * xq;
        xp = a;
        xq = b;
    }
};

}  // namespace

int gmres(const SpMat& A, const double* rhs, double* x, OTree& precond, int iters, int restart, double tol, bool verb) {
    const int m = A.rows;
    double rn = 0;
    for (int i = 0; i < m; i++) rn += rhs[i] * rhs[i];
    if (std::sqrt(rn) <= std::numeric_limits<double>::min()) {
        std::fill(x, x + m, 0.0);
        return 1;  // the reference returns `true`
    }
    const int maxIters = iters;
    iters = 0;
    std::vector<double> r0(m), t(m), v(m), x_new(m);
    auto residual = [&]() {
        spand::spmv(A, x, t.data());
        for (int i = 0; i < m; i++) r0[i] = rhs[i] - t[i];
        precond.solve(r0.data());
    };
    residual();
    double r0Norm = 0;
    for (int i = 0; i < m; i++) r0Norm += r0[i] * r0[i];
    r0Norm = std::sqrt(r0Norm);
    if (r0Norm == 0) return 1;
    std::vector<double> H((size_t)m * (restart + 1), 0.0), w(restart + 1, 0.0), tau(restart + 1, 0.0);
    std::vector<Givens> G(restart);
    auto Hcol = [&](int i) { return H.data() + (size_t)i * m; };
    double beta;
    make_householder(r0.data(), m, Hcol(0) + 1, tau[0], beta);
    w[0] = beta;
    for (int k = 1; k <= restart; ++k) {
        ++iters;
        std::fill(v.begin(), v.end(), 0.0);
        v[k - 1] = 1.0;
        for (int i = k - 1; i >= 0; --i) apply_householder_left(v.data() + i, m - i, Hcol(i) + i + 1, tau[i]);
        spand::spmv(A, v.data(), t.data());
        v = t;
        precond.solve(v.data());
        for (int i = 0; i < k; ++i) apply_householder_left(v.data() + i, m - i, Hcol(i) + i + 1, tau[i]);
        double tn = 0;
        for (int i = k; i < m; i++) tn += v[i] * v[i];
        if (m - k > 0 && std::sqrt(tn) != 0.0) {
            make_householder(v.data() + k, m - k, Hcol(k) + k + 1, tau[k], beta);
            apply_householder_left(v.data() + k, m - k, Hcol(k) + k + 1, tau[k]);
        }
        for (int i = 0; i < k - 1; ++i) G[i].apply_adjoint(v[i], v[i + 1]);
        if (k < m && v[k] != 0.0) {
            G[k - 1].make(v[k - 1], v[k]);
            G[k - 1].apply_adjoint(v[k - 1], v[k]);
            G[k - 1].apply_adjoint(w[k - 1], w[k]);
        }
        for (int i = 0; i < k; i++) Hcol(k - 1)[i] = v[i];
        double tol_error = std::fabs(w[k]) / r0Norm;
        bool stop = (k == m || tol_error < tol || iters == maxIters);
        if (verb) printf("%d: |Ax-b|/|b| = %3.2e <? %3.2e\n", iters, tol_error, tol);
        if (stop || k == restart) {
            std::vector<double> y(w.begin(), w.begin() + k);
            for (int i = k - 1; i >= 0; --i) {  // upper-triangular solve with H(0:k, 0:k)
                for (int j = i + 1; j < k; j++) y[i] -= Hcol(j)[i] * y[j];
                y[i] /= Hcol(i)[i];
            }
            std::fill(x_new.begin(), x_new.end(), 0.0);
            for (int i = k - 1; i >= 0; --i) {
                x_new[i] += y[i];
                apply_householder_left(x_new.data() + i, m - i, Hcol(i) + i + 1, tau[i]);
            }
            for (int i = 0; i < m; i++) x[i] += x_new[i];
            if (stop) return iters;
            k = 0;
            residual();
            std::fill(H.begin(), H.end(), 0.0);
            std::fill(w.begin(), w.end(), 0.0);
            std::fill(tau.begin(), tau.end(), 0.0);
            make_householder(r0.data(), m, Hcol(0) + 1, tau[0], beta);
            w[0] = beta;
        }
    }
    return iters;
}

}  // namespace spand_oracle
