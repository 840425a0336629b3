// Host-side sparse/dense containers and generators (no Eigen).
//
// These replace the pieces of Eigen + include/mmio.hpp + src/util.cpp that the
// reference uses *around* the factorization hot path (reference citations are
// relative to /root/reference):
//   symmetric_graph      src/util.cpp:47-63
//   symm_perm            src/util.cpp:520-547
//   linspace_nd          src/util.cpp:488-517
//   random               src/util.cpp:549-558
//   sp_mmread            include/mmio.hpp:160-220
//   dense_mmread         include/mmio.hpp:225-262
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace spand {

// Compressed sparse column, int32 indices, rows sorted inside each column.
struct SpMat {
    int rows = 0, cols = 0;
    std::vector<int> colptr;   // cols + 1
    std::vector<int> rowind;   // nnz
    std::vector<double> val;   // nnz
    int nnz() const { return (int)rowind.size(); }
};

struct Triplet {
    int r, c;
    double v;
};

// Column-major dense matrix.
struct DenseMat {
    int rows = 0, cols = 0;
    std::vector<double> a;
    DenseMat() {}
    DenseMat(int r, int c) : rows(r), cols(c), a((size_t)r * c, 0.0) {}
    double& operator()(int i, int j) { return a[(size_t)j * rows + i]; }
    double operator()(int i, int j) const { return a[(size_t)j * rows + i]; }
};

// Duplicates are summed, like Eigen's setFromTriplets.
SpMat from_triplets(int rows, int cols, const std::vector<Triplet>& t);
SpMat from_csc(int n, const int* colptr, const int* rowind, const double* val);

// |A| + |A^T| + I
// A^T by one counting pass (rows of every column come out sorted)
SpMat transpose(const SpMat& A);
SpMat symmetric_graph(const SpMat& A);
// A[p,p]: entry (i,j) goes to (pinv[i], pinv[j])
SpMat symm_perm(const SpMat& A, const std::vector<int>& p);
// y = A x
void spmv(const SpMat& A, const double* x, double* y);

// d-dimensional Dirichlet stencil Laplacian, diag 2d, off-diag -1, dof index
// x + n*y (+ n^2*z); same matrices as mats/neglapl_d_n.mm once read (full storage).
SpMat neglapl(int n, int d);
// Config C5 of BASELINE.json (definition in SURVEY.md section 8d): 7-point finite-volume
// convection-diffusion with anisotropic variable coefficient; non-symmetric M-matrix.
SpMat aniso_convdiff(int n);
// dim x n^dim coordinates, column-major, first coordinate slowest.
DenseMat linspace_nd(int n, int dim);

// mt19937 + uniform_real_distribution(-1,1)
std::vector<double> random_vec(int size, int seed);

SpMat mm_read_sparse(const std::string& fn);
DenseMat mm_read_dense(const std::string& fn);
void mm_write_sparse(const std::string& fn, const SpMat& A, bool lower_hermitian);

}  // namespace spand
