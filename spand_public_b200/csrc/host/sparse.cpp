#include "sparse.hpp"

#include <algorithm>
#include <cmath>
#include <fstream>
#include <iomanip>
#include <numeric>
#include <random>
#include <sstream>
#include <stdexcept>

namespace spand {

SpMat from_triplets(int rows, int cols, const std::vector<Triplet>& t) {
    SpMat A;
    A.rows = rows;
    A.cols = cols;
    // Counting sort by column, then sort rows inside each column and sum duplicates.
    std::vector<int> cnt(cols + 1, 0);
    for (auto& e : t) cnt[e.c + 1]++;
    for (int j = 0; j < cols; j++) cnt[j + 1] += cnt[j];
    std::vector<int> pos(cnt.begin(), cnt.end() - 1);
    std::vector<int> ri(t.size());
    std::vector<double> vv(t.size());
    for (auto& e : t) {
        int p = pos[e.c]++;
        ri[p] = e.r;
        vv[p] = e.v;
    }
    A.colptr.assign(cols + 1, 0);
    A.rowind.reserve(t.size());
    A.val.reserve(t.size());
    std::vector<int> idx;
    for (int j = 0; j < cols; j++) {
        int b = cnt[j], e = cnt[j + 1];
        idx.resize(e - b);
        std::iota(idx.begin(), idx.end(), b);
        std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return ri[x] < ri[y]; });
        for (size_t k = 0; k < idx.size(); k++) {
            int r = ri[idx[k]];
            double v = vv[idx[k]];
            if (k > 0 && A.rowind.back() == r && (int)A.rowind.size() > A.colptr[j]) {
                A.val.back() += v;
            } else {
                A.rowind.push_back(r);
                A.val.push_back(v);
            }
        }
        A.colptr[j + 1] = (int)A.rowind.size();
    }
    return A;
}

SpMat from_csc(int n, const int* colptr, const int* rowind, const double* val) {
    // canonical input (rows strictly increasing inside every column, the normal case): plain copies
    bool canonical = colptr[0] == 0;
    for (int j = 0; j < n && canonical; j++) {
        canonical = colptr[j + 1] >= colptr[j];
        for (int k = colptr[j] + 1; k < colptr[j + 1] && canonical; k++) canonical = rowind[k] > rowind[k - 1];
        if (canonical && colptr[j + 1] > colptr[j])
            canonical = rowind[colptr[j]] >= 0 && rowind[colptr[j + 1] - 1] < n;
    }
    if (canonical) {
        SpMat A;
        A.rows = A.cols = n;
        const size_t nnz = (size_t)colptr[n];
        A.colptr.assign(colptr, colptr + n + 1);
        A.rowind.assign(rowind, rowind + nnz);
        if (val) A.val.assign(val, val + nnz);
        else A.val.assign(nnz, 1.0);
        return A;
    }
    // general input (unsorted rows, duplicates, empty columns): validate, then go through triplets
    if (colptr[0] != 0) throw std::runtime_error("from_csc: colptr[0] must be 0");
    for (int j = 0; j < n; j++)
        if (colptr[j + 1] < colptr[j]) throw std::runtime_error("from_csc: colptr must be non-decreasing");
    std::vector<Triplet> t;
    t.reserve(colptr[n]);
    for (int j = 0; j < n; j++)
        for (int k = colptr[j]; k < colptr[j + 1]; k++) {
            if (rowind[k] < 0 || rowind[k] >= n) throw std::runtime_error("from_csc: row index out of range");
            t.push_back({rowind[k], j, val ? val[k] : 1.0});
        }
    return from_triplets(n, n, t);
}

SpMat transpose(const SpMat& A) {
    SpMat T;
    T.rows = A.cols;
    T.cols = A.rows;
    const size_t nnz = A.rowind.size();
    T.colptr.assign(A.rows + 1, 0);
    for (size_t k = 0; k < nnz; k++) T.colptr[A.rowind[k] + 1]++;
    for (int i = 0; i < A.rows; i++) T.colptr[i + 1] += T.colptr[i];
    std::vector<int> pos(T.colptr.begin(), T.colptr.end() - 1);
    T.rowind.resize(nnz);
    T.val.resize(nnz);
    for (int j = 0; j < A.cols; j++)
        for (int k = A.colptr[j]; k < A.colptr[j + 1]; k++) {
            const int p = pos[A.rowind[k]]++;
            T.rowind[p] = j;
            T.val[p] = A.val[k];
        }
    return T;
}

SpMat symmetric_graph(const SpMat& A) {
    int n = A.rows;
    std::vector<Triplet> t;
    t.reserve(2 * (size_t)A.nnz() + n);
    for (int j = 0; j < n; j++) {
        t.push_back({j, j, 1.0});
        for (int k = A.colptr[j]; k < A.colptr[j + 1]; k++) {
            double v = std::fabs(A.val[k]);
            t.push_back({j, A.rowind[k], v});
            t.push_back({A.rowind[k], j, v});
        }
    }
    return from_triplets(n, n, t);
}

SpMat symm_perm(const SpMat& A, const std::vector<int>& p) {
    int n = A.rows;
    std::vector<int> pinv(n);
    for (int i = 0; i < n; i++) pinv[p[i]] = i;
    std::vector<Triplet> t;
    t.reserve(A.nnz());
    for (int j = 0; j < n; j++)
        for (int k = A.colptr[j]; k < A.colptr[j + 1]; k++) t.push_back({pinv[A.rowind[k]], pinv[j], A.val[k]});
    return from_triplets(n, n, t);
}

void spmv(const SpMat& A, const double* x, double* y) {
    std::fill(y, y + A.rows, 0.0);
    for (int j = 0; j < A.cols; j++) {
        double xj = x[j];
        for (int k = A.colptr[j]; k < A.colptr[j + 1]; k++) y[A.rowind[k]] += A.val[k] * xj;
    }
}

SpMat neglapl(int n, int d) {
    long N = 1;
    for (int i = 0; i < d; i++) N *= n;
    SpMat A;
    A.rows = A.cols = (int)N;
    A.colptr.assign(N + 1, 0);
    A.rowind.reserve(N * (2 * d + 1));
    A.val.reserve(N * (2 * d + 1));
    std::vector<long> stride(d);
    stride[0] = 1;
    for (int i = 1; i < d; i++) stride[i] = stride[i - 1] * n;
    std::vector<int> c(d);
    for (long j = 0; j < N; j++) {
        long r = j;
        for (int i = 0; i < d; i++) {
            c[i] = (int)(r % n);
            r /= n;
        }
        // rows ascending: -stride[d-1] ... -stride[0], diag, +stride[0] ... +stride[d-1]
        for (int i = d - 1; i >= 0; i--)
            if (c[i] > 0) {
                A.rowind.push_back((int)(j - stride[i]));
                A.val.push_back(-1.0);
            }
        A.rowind.push_back((int)j);
        A.val.push_back(2.0 * d);
        for (int i = 0; i < d; i++)
            if (c[i] < n - 1) {
                A.rowind.push_back((int)(j + stride[i]));
                A.val.push_back(-1.0);
            }
        A.colptr[j + 1] = (int)A.rowind.size();
    }
    return A;
}

SpMat aniso_convdiff(int n) {
    const double PI = 3.14159265358979323846;
    const double eps[3] = {1.0, 1e-1, 1e-2};
    const double conv = 1.0;  // cell Peclet 0.5 w.r.t. the unit reference diffusivity
    long N = (long)n * n * n;
    double h = 1.0 / n;
    auto kappa = [&](int i, int j, int k) {
        double x = (i + 0.5) * h, y = (j + 0.5) * h, z = (k + 0.5) * h;
        return std::pow(10.0, std::sin(2 * PI * x) * std::sin(2 * PI * y) * std::sin(2 * PI * z));
    };
    std::vector<Triplet> t;
    t.reserve(7 * N);
    const long stride[3] = {1, n, (long)n * n};
    for (int k = 0; k < n; k++)
        for (int j = 0; j < n; j++)
            for (int i = 0; i < n; i++) {
                long row = i + (long)n * j + (long)n * n * k;
                int c[3] = {i, j, k};
                double ka = kappa(i, j, k);
                double diag = 0.0;
                for (int d = 0; d < 3; d++) {
                    for (int s = -1; s <= 1; s += 2) {
                        int cc[3] = {c[0], c[1], c[2]};
                        cc[d] += s;
                        if (cc[d] < 0 || cc[d] >= n) {
                            diag += eps[d] * ka;  // Dirichlet ghost cell
                            continue;
                        }
                        double kb = kappa(cc[0], cc[1], cc[2]);
                        double kf = eps[d] * 2.0 * ka * kb / (ka + kb);
                        diag += kf;
                        double off = -kf;
                        if (s < 0) off -= conv;  // first-order upwind, beta_d > 0
                        t.push_back({(int)row, (int)(row + s * stride[d]), off});
                    }
                    diag += conv;
                }
                t.push_back({(int)row, (int)row, diag});
            }
    return from_triplets((int)N, (int)N, t);
}

DenseMat linspace_nd(int n, int dim) {
    long N = 1;
    for (int i = 0; i < dim; i++) N *= n;
    DenseMat X(dim, (int)N);
    for (long id = 0; id < N; id++) {
        long r = id;
        for (int d = dim - 1; d >= 0; d--) {
            X(d, (int)id) = (double)(r % n);
            r /= n;
        }
    }
    return X;
}

std::vector<double> random_vec(int size, int seed) {
    std::mt19937 rng;
    rng.seed(seed);
    std::uniform_real_distribution<double> dist(-1.0, 1.0);
    std::vector<double> x(size);
    for (int i = 0; i < size; i++) x[i] = dist(rng);
    return x;
}

namespace {
std::string lower(std::string s) {
    std::transform(s.begin(), s.end(), s.begin(), ::tolower);
    return s;
}
struct MMHeader {
    bool coordinate, pattern, integer_or_real;
    enum Prop { general, symmetric, hermitian, skew } prop;
};
MMHeader parse_header(const std::string& line) {
    std::istringstream is(line);
    std::string banner, object, format, type, prop;
    is >> banner >> object >> format >> type >> prop;
    if (banner != "%%MatrixMarket" || lower(object) != "matrix") throw std::runtime_error("mm: bad banner");
    MMHeader h;
    format = lower(format);
    type = lower(type);
    prop = lower(prop);
    if (format == "coordinate") h.coordinate = true;
    else if (format == "array") h.coordinate = false;
    else throw std::runtime_error("mm: bad format");
    h.pattern = (type == "pattern");
    h.integer_or_real = (type == "real" || type == "integer");
    if (!h.pattern && !h.integer_or_real) throw std::runtime_error("mm: unsupported value type " + type);
    if (prop == "general") h.prop = MMHeader::general;
    else if (prop == "symmetric") h.prop = MMHeader::symmetric;
    else if (prop == "hermitian") h.prop = MMHeader::hermitian;
    else if (prop == "skew-symmetric") h.prop = MMHeader::skew;
    else throw std::runtime_error("mm: bad property");
    return h;
}
bool next_data_line(std::ifstream& f, std::string& line) {
    while (std::getline(f, line)) {
        if (line.empty() || line[0] == '%') continue;
        return true;
    }
    return false;
}
}  // namespace

SpMat mm_read_sparse(const std::string& fn) {
    std::ifstream f(fn);
    if (!f.is_open()) throw std::runtime_error("Couldn't open " + fn);
    std::string line;
    std::getline(f, line);
    MMHeader h = parse_header(line);
    if (!h.coordinate || h.pattern) throw std::runtime_error("mm: need coordinate, non-pattern");
    if (!next_data_line(f, line)) throw std::runtime_error("mm: no size line");
    int M, N, K;
    {
        std::istringstream is(line);
        is >> M >> N >> K;
    }
    std::vector<Triplet> t;
    t.reserve(h.prop == MMHeader::general ? K : 2 * (size_t)K);
    int nread = 0;
    while (next_data_line(f, line)) {
        std::istringstream is(line);
        int i, j;
        double v;
        is >> i >> j >> v;
        i--;
        j--;
        if (h.prop != MMHeader::general && i < j) throw std::runtime_error("mm: upper entry in symmetric file");
        t.push_back({i, j, v});
        if (i != j && (h.prop == MMHeader::symmetric || h.prop == MMHeader::hermitian)) t.push_back({j, i, v});
        if (h.prop == MMHeader::skew) {
            if (i == j) throw std::runtime_error("mm: diagonal entry in skew-symmetric file");
            t.push_back({j, i, -v});
        }
        nread++;
    }
    if (nread != K) throw std::runtime_error("mm: entry count mismatch");
    return from_triplets(M, N, t);
}

DenseMat mm_read_dense(const std::string& fn) {
    std::ifstream f(fn);
    if (!f.is_open()) throw std::runtime_error("Couldn't open " + fn);
    std::string line;
    std::getline(f, line);
    MMHeader h = parse_header(line);
    if (h.coordinate || h.pattern || h.prop != MMHeader::general) throw std::runtime_error("mm: need array general");
    if (!next_data_line(f, line)) throw std::runtime_error("mm: no size line");
    int M, N;
    {
        std::istringstream is(line);
        is >> M >> N;
    }
    DenseMat A(M, N);
    long nread = 0;
    while (next_data_line(f, line)) {
        if (nread >= (long)M * N) throw std::runtime_error("mm: too many entries");
        A.a[nread++] = std::stod(line);
    }
    if (nread != (long)M * N) throw std::runtime_error("mm: entry count mismatch");
    return A;
}

void mm_write_sparse(const std::string& fn, const SpMat& A, bool lower_hermitian) {
    std::ofstream f(fn);
    if (!f.is_open()) throw std::runtime_error("Couldn't open " + fn);
    f << "%%MatrixMarket matrix coordinate real " << (lower_hermitian ? "hermitian" : "general") << "\n";
    long cnt = 0;
    for (int j = 0; j < A.cols; j++)
        for (int k = A.colptr[j]; k < A.colptr[j + 1]; k++)
            if (!lower_hermitian || A.rowind[k] >= j) cnt++;
    f << A.rows << " " << A.cols << " " << cnt << "\n";
    f << std::setprecision(17);
    for (int j = 0; j < A.cols; j++)
        for (int k = A.colptr[j]; k < A.colptr[j + 1]; k++)
            if (!lower_hermitian || A.rowind[k] >= j) f << A.rowind[k] + 1 << " " << j + 1 << " " << A.val[k] << "\n";
}

}  // namespace spand
