// sm_100a kernels for the spaND factorization hot path. See kernels.cuh for the mapping to the
// reference's BLAS/LAPACK call sites. All matrices are FP64, column-major.
#include <cfloat>
#include <climits>
#include <cstdio>

#include "kernels.cuh"

namespace spand {

namespace {

constexpr int LDS = NB + 1;  // padded leading dimension of 64x64 shared tiles

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------
// POTRF, one 64x64 diagonal block per CTA (left-looking, one thread per row).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NB) potrf_step_kernel(const PotrfTask* __restrict__ tasks, int j0, int* err) {
    PotrfTask t = tasks[blockIdx.x];
    int nb = min(NB, t.n - j0);
    if (nb <= 0) return;
    double* A = t.A + j0 + (size_t)j0 * t.ld;
    __shared__ double S[NB * LDS];
    int i = threadIdx.x;
    for (int j = 0; j < nb; j++)
        if (i < nb && i >= j) S[j * LDS + i] = A[i + (size_t)j * t.ld];
    __syncthreads();
    for (int k = 0; k < nb; k++) {
        double v = 0.0;
        if (i >= k && i < nb) {
            v = S[k * LDS + i];
            for (int p = 0; p < k; p++) v -= S[p * LDS + i] * S[p * LDS + k];
        }
        if (i == k) {
            if (!(v > 0.0)) atomicOr(err, 1);  // LAPACK: ajj <= 0 or NaN -> info > 0
            S[k * LDS + k] = sqrt(v);
        }
        __syncthreads();
        if (i > k && i < nb) S[k * LDS + i] = v / S[k * LDS + k];
        __syncthreads();
    }
    for (int j = 0; j < nb; j++)
        if (i < nb && i >= j) A[i + (size_t)j * t.ld] = S[j * LDS + i];
}

// ------------------------------------------------------------------------------------------------
// TRSM against one 64x64 diagonal block of the triangle; CTA = (task, 64-wide strip of the free dim).
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(NB) trsm_step_kernel(const TrsmTask* __restrict__ tasks, int j0) {
    TrsmTask t = tasks[blockIdx.x];
    int nb = min(NB, t.n - j0);
    int f0 = blockIdx.y * NB;  // offset in the free dimension
    if (nb <= 0 || f0 >= t.m) return;
    int fw = min(NB, t.m - f0);
    extern __shared__ double trsm_smem[];
    double* Ts = trsm_smem;
    double* Xs = trsm_smem + NB * LDS;
    int tid = threadIdx.x;
    const double* T = t.T + j0 + (size_t)j0 * t.ldt;
    // Ts[p*LDS + j] = coefficient multiplying unknown p in equation j (p <= j)
    if (MODE == TRSM_RUN) {
        // U[p][j], p <= j : column j contiguous in p
        for (int j = 0; j < nb; j++)
            if (tid <= j) Ts[tid * LDS + j] = T[tid + (size_t)j * t.ldt];
    } else {
        // L[j][p], p <= j : column p contiguous in j
        for (int p = 0; p < nb; p++)
            if (tid >= p && tid < nb) Ts[p * LDS + tid] = T[tid + (size_t)p * t.ldt];
    }
    if (MODE == TRSM_LLU) {
        if (tid < nb) Ts[tid * LDS + tid] = 1.0;
    } else if (t.diag != nullptr) {
        if (tid < nb) Ts[tid * LDS + tid] = t.diag[j0 + tid];
    }
    if (MODE == TRSM_LLN || MODE == TRSM_LLU) {
        // X = B[j0:j0+nb, f0:f0+fw]; Xs[c*LDS + i]
        double* B = t.B + j0 + (size_t)f0 * t.ldb;
        for (int c = 0; c < fw; c++)
            if (tid < nb) Xs[c * LDS + tid] = B[tid + (size_t)c * t.ldb];
        __syncthreads();
        if (tid < fw) {
            double* x = Xs + tid * LDS;
            for (int i = 0; i < nb; i++) {
                double v = x[i];
                for (int p = 0; p < i; p++) v -= Ts[p * LDS + i] * x[p];
                x[i] = v / Ts[i * LDS + i];
            }
        }
        __syncthreads();
        for (int c = 0; c < fw; c++)
            if (tid < nb) B[tid + (size_t)c * t.ldb] = Xs[c * LDS + tid];
    } else {
        // X = B[f0:f0+fw, j0:j0+nb]; Xs[j*LDS + r]
        double* B = t.B + f0 + (size_t)j0 * t.ldb;
        for (int j = 0; j < nb; j++)
            if (tid < fw) Xs[j * LDS + tid] = B[tid + (size_t)j * t.ldb];
        __syncthreads();
        if (tid < fw) {
            for (int j = 0; j < nb; j++) {
                double v = Xs[j * LDS + tid];
                for (int p = 0; p < j; p++) v -= Xs[p * LDS + tid] * Ts[p * LDS + j];
                Xs[j * LDS + tid] = v / Ts[j * LDS + j];
            }
            for (int j = 0; j < nb; j++) B[tid + (size_t)j * t.ldb] = Xs[j * LDS + tid];
        }
    }
}


// ------------------------------------------------------------------------------------------------
// GETRF with partial pivoting (dgetrf semantics: first maximal |a| wins) + split_LU (src/util.cpp:213-227).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_entry(double* A, int ld, int i, int j, const double* d) {
    // d = diag of the LAPACK factor. L(i,j) = l_ij |d_j|^1/2 ; U(i,j) = s_i (u_ij / d_i), s_i = sign(d_i) |d_i|^1/2
    double a = A[i + (size_t)j * ld];
    if (i > j) A[i + (size_t)j * ld] = a * sqrt(fabs(d[j]));
    else if (i < j) {
        double di = d[i];
        double si = (di > 0 ? 1.0 : (di < 0 ? -1.0 : 0.0)) * sqrt(fabs(di));
        A[i + (size_t)j * ld] = si * ((1.0 / di) * a);
    } else A[i + (size_t)j * ld] = sqrt(fabs(a));
}

__global__ void __launch_bounds__(NB) getrf_small_kernel(const GetrfTask* __restrict__ tasks, int* err) {
    GetrfTask t = tasks[blockIdx.x];
    const int n = t.n;
    if (n <= 0) return;
    __shared__ double S[NB * LDS];  // S[j * LDS + i] = A(i, j)
    __shared__ double dd[NB];
    __shared__ int piv_s, prm[NB];
    const int i = threadIdx.x;
    for (int j = 0; j < n; j++)
        if (i < n) S[j * LDS + i] = t.A[i + (size_t)j * t.ld];
    if (i < n) prm[i] = i;
    __syncthreads();
    for (int k = 0; k < n; k++) {
        // pivot search: first row with the largest |a_ik|, i >= k (two warps)
        double v = (i >= k && i < n) ? fabs(S[k * LDS + i]) : -1.0;
        int idx = i;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, v, o);
            int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (ov > v || (ov == v && oi < idx)) {
                v = ov;
                idx = oi;
            }
        }
        __shared__ double wv[2];
        __shared__ int wi[2];
        if ((i & 31) == 0) {
            wv[i >> 5] = v;
            wi[i >> 5] = idx;
        }
        __syncthreads();
        if (i == 0) {
            int p = (wv[1] > wv[0]) ? wi[1] : wi[0];  // ties: warp 0 holds the smaller indices
            piv_s = p;
            t.ipiv[k] = p;
            int tmp = prm[p];
            prm[p] = prm[k];
            prm[k] = tmp;
        }
        __syncthreads();
        const int p = piv_s;
        if (p != k && i < n) {  // swap rows k and p (thread i = column i)
            double a = S[i * LDS + k];
            S[i * LDS + k] = S[i * LDS + p];
            S[i * LDS + p] = a;
        }
        __syncthreads();
        const double akk = S[k * LDS + k];
        if (akk == 0.0) {
            if (i == 0) atomicOr(err, 2);  // dgetrf info > 0: exactly singular
        } else if (i > k && i < n) {
            double l = S[k * LDS + i] / akk;
            S[k * LDS + i] = l;
            for (int j = k + 1; j < n; j++) S[j * LDS + i] -= l * S[j * LDS + k];
        }
        __syncthreads();
    }
    if (i < n) {
        dd[i] = S[i * LDS + i];
        t.perm[i] = prm[i];
    }
    __syncthreads();
    if (i < n) {
        double di = dd[i];
        t.ud[i] = (di > 0 ? 1.0 : (di < 0 ? -1.0 : 0.0)) * sqrt(fabs(di));
    }
    for (int j = 0; j < n; j++)
        if (i < n) {
            double a = S[j * LDS + i], o;
            if (i > j) o = a * sqrt(fabs(dd[j]));
            else if (i < j) {
                double di = dd[i];
                double si = (di > 0 ? 1.0 : (di < 0 ? -1.0 : 0.0)) * sqrt(fabs(di));
                o = si * ((1.0 / di) * a);
            } else o = sqrt(fabs(a));
            t.A[i + (size_t)j * t.ld] = o;
        }
}

// Panel [j0, j0 + nbp) x rows [j0, n) of every matrix with n > j0, factored in place (global memory / L2).
constexpr int GP_T = 256;
__global__ void __launch_bounds__(GP_T) getrf_panel_kernel(const GetrfTask* __restrict__ tasks, int j0, int* err) {
    GetrfTask t = tasks[blockIdx.x];
    const int n = t.n;
    if (n <= j0) return;
    const int nbp = min(NB, n - j0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ double wv[GP_T / 32];
    __shared__ int wi[GP_T / 32];
    __shared__ int piv_s;
    __shared__ double rowk[NB];
    double* A = t.A;
    const size_t ld = t.ld;
    for (int c = 0; c < nbp; c++) {
        const int col = j0 + c;
        double v = -1.0;
        int idx = INT_MAX;
        for (int r = col + tid; r < n; r += GP_T) {
            double a = fabs(A[r + col * ld]);
            if (a > v) {  // rows visited in increasing order: keeps the first maximum
                v = a;
                idx = r;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, v, o);
            int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (ov > v || (ov == v && oi < idx)) {
                v = ov;
                idx = oi;
            }
        }
        if (lane == 0) {
            wv[warp] = v;
            wi[warp] = idx;
        }
        __syncthreads();
        if (tid == 0) {
            double bv = wv[0];
            int bi = wi[0];
            for (int w = 1; w < GP_T / 32; w++)
                if (wv[w] > bv || (wv[w] == bv && wi[w] < bi)) {
                    bv = wv[w];
                    bi = wi[w];
                }
            piv_s = bi;
            t.ipiv[col] = bi;
        }
        __syncthreads();
        const int p = piv_s;
        if (tid < nbp) {  // swap inside the panel, keep the (new) pivot row in shared memory
            double a = A[col + (j0 + tid) * ld], b = A[p + (j0 + tid) * ld];
            if (p != col) {
                A[col + (j0 + tid) * ld] = b;
                A[p + (j0 + tid) * ld] = a;
            } else b = a;
            rowk[tid] = b;
        }
        __syncthreads();
        const double akk = rowk[c];
        if (akk == 0.0) {
            if (tid == 0) atomicOr(err, 2);
            continue;
        }
        const int nr = n - col - 1, ncc = nbp - c - 1;
        for (int r = tid; r < nr; r += GP_T) A[col + 1 + r + col * ld] /= akk;
        __syncthreads();
        for (int e = tid; e < nr * ncc; e += GP_T) {
            int r = e % nr, cc = e / nr;
            A[col + 1 + r + (col + 1 + cc) * ld] -= A[col + 1 + r + col * ld] * rowk[c + 1 + cc];
        }
        __syncthreads();
    }
}

// Row swaps of panel j0 applied to the columns outside the panel; thread per column.
__global__ void __launch_bounds__(128) getrf_laswp_kernel(const GetrfTask* __restrict__ tasks, int j0) {
    GetrfTask t = tasks[blockIdx.x];
    const int n = t.n;
    if (n <= j0) return;
    const int nbp = min(NB, n - j0);
    int c = blockIdx.y * 128 + threadIdx.x;  // index among the n - nbp outside columns
    if (c >= n - nbp) return;
    if (c >= j0) c += nbp;
    double* col = t.A + (size_t)c * t.ld;
    for (int k = j0; k < j0 + nbp; k++) {
        int p = t.ipiv[k];
        if (p != k) {
            double a = col[k];
            col[k] = col[p];
            col[p] = a;
        }
    }
}

__global__ void __launch_bounds__(256) getrf_finish_kernel(const GetrfTask* __restrict__ tasks) {
    GetrfTask t = tasks[blockIdx.x];
    const int n = t.n;
    if (n <= NB) return;  // small pivots were finished by getrf_small_kernel
    extern __shared__ double dfin[];  // n: the LAPACK diagonal
    const int tid = threadIdx.x;
    for (int i = tid; i < n; i += 256) dfin[i] = t.A[i + (size_t)i * t.ld];
    if (tid == 0) {  // swap2perm (src/util.cpp:76-88)
        for (int i = 0; i < n; i++) t.perm[i] = i;
        for (int i = 0; i < n; i++) {
            int p = t.ipiv[i];
            int tmp = t.perm[p];
            t.perm[p] = t.perm[i];
            t.perm[i] = tmp;
        }
    }
    __syncthreads();
    for (int i = tid; i < n; i += 256) {
        double di = dfin[i];
        t.ud[i] = (di > 0 ? 1.0 : (di < 0 ? -1.0 : 0.0)) * sqrt(fabs(di));
    }
    for (size_t e = tid; e < (size_t)n * n; e += 256) split_entry(t.A, t.ld, (int)(e % n), (int)(e / n), dfin);
}

__global__ void __launch_bounds__(128) rowperm_kernel(const RowPermTask* __restrict__ tasks) {
    RowPermTask t = tasks[blockIdx.x];
    extern __shared__ double pbuf[];  // n x w strip
    const int n = t.n;
    if (n == 0 || t.m == 0) return;
    const int w = max(1, min(t.m, 6144 / n));
    for (int c0 = 0; c0 < t.m; c0 += w) {
        const int cw = min(w, t.m - c0);
        for (int e = threadIdx.x; e < n * cw; e += 128) {
            int i = e % n, c = e / n;
            pbuf[e] = t.B[t.perm[i] + (size_t)(c0 + c) * t.ldb];
        }
        __syncthreads();
        for (int e = threadIdx.x; e < n * cw; e += 128) {
            int i = e % n, c = e / n;
            t.B[i + (size_t)(c0 + c) * t.ldb] = pbuf[e];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Grouped GEMM  C = (0|C) - sum_c A_c op(B_c), FP64 tensor cores (DMMA m8n8k4), 64x64x16 CTA tiles.
// ------------------------------------------------------------------------------------------------
constexpr int GT = 64;         // tile edge
constexpr int GK = 16;         // k chunk
constexpr int GLD = GT + 4;    // (GLD mod 16) == 4 -> conflict-free fragment loads

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) gemm_tiled_kernel(const GemmTask* __restrict__ tasks, int nt,
                                                         const GemmContrib* __restrict__ contribs,
                                                         const int* __restrict__ tile_prefix) {
    int b = blockIdx.x;
    int lo = 0, hi = nt - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (tile_prefix[mid] <= b) lo = mid;
        else hi = mid - 1;
    }
    GemmTask t = tasks[lo];
    int local = b - tile_prefix[lo];
    int tm = (t.m + GT - 1) / GT;
    int tile_r = local % tm, tile_c = local / tm;
    if ((t.flags & GEMM_LOWER) && tile_c > tile_r) return;
    int row0 = tile_r * GT, col0 = tile_c * GT;
    bool nn = (t.flags & GEMM_NN) != 0;

    __shared__ double As[GK * GLD];
    __shared__ double Bs[GK * GLD];
    int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int wm = warp & 3, wn = warp >> 2;
    double acc[2][4][2];
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[a][c][0] = acc[a][c][1] = 0.0;

    int arow = tid & 63, ak = tid >> 6;  // A loader: rows contiguous
    int bk_nn = tid & 15, bj_nn = tid >> 4;

    for (int ci = 0; ci < t.nc; ci++) {
        GemmContrib c = contribs[t.c0 + ci];
        for (int k0 = 0; k0 < c.k; k0 += GK) {
            double ra[4], rb[4];
#pragma unroll
            for (int s = 0; s < 4; s++) {
                int kk = k0 + ak + 4 * s;
                int gi = row0 + arow;
                ra[s] = (gi < t.m && kk < c.k) ? c.A[gi + (size_t)kk * c.lda] : 0.0;
                if (!nn) {
                    int gj = col0 + arow;
                    rb[s] = (gj < t.n && kk < c.k) ? c.B[gj + (size_t)kk * c.ldb] : 0.0;
                } else {
                    int kk2 = k0 + bk_nn, gj = col0 + bj_nn + 16 * s;
                    rb[s] = (gj < t.n && kk2 < c.k) ? c.B[kk2 + (size_t)gj * c.ldb] : 0.0;
                }
            }
            __syncthreads();
#pragma unroll
            for (int s = 0; s < 4; s++) {
                As[(ak + 4 * s) * GLD + arow] = ra[s];
                if (!nn) Bs[(ak + 4 * s) * GLD + arow] = rb[s];
                else Bs[bk_nn * GLD + bj_nn + 16 * s] = rb[s];
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < GK; kk += 4) {
                double fa[2], fb[4];
                int kr = (kk + (lane & 3)) * GLD;
#pragma unroll
                for (int mi = 0; mi < 2; mi++) fa[mi] = As[kr + wm * 16 + mi * 8 + (lane >> 2)];
#pragma unroll
                for (int ni = 0; ni < 4; ni++) fb[ni] = Bs[kr + wn * 32 + ni * 8 + (lane >> 2)];
#pragma unroll
                for (int mi = 0; mi < 2; mi++)
#pragma unroll
                    for (int ni = 0; ni < 4; ni++) dmma(acc[mi][ni][0], acc[mi][ni][1], fa[mi], fb[ni]);
            }
        }
    }
    bool zero = (t.flags & GEMM_ZERO_INIT) != 0, lower = (t.flags & GEMM_LOWER) != 0;
#pragma unroll
    for (int mi = 0; mi < 2; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                int gi = row0 + wm * 16 + mi * 8 + (lane >> 2);
                int gj = col0 + wn * 32 + ni * 8 + 2 * (lane & 3) + e;
                if (gi < t.m && gj < t.n && (!lower || gi >= gj)) {
                    double* p = t.C + gi + (size_t)gj * t.ldc;
                    *p = (zero ? 0.0 : *p) - acc[mi][ni][e];
                }
            }
}

// Tiny targets: one warp per target, plain FMAs out of L1/L2.
__global__ void __launch_bounds__(128) gemm_small_kernel(const GemmTask* __restrict__ tasks, int nt,
                                                         const GemmContrib* __restrict__ contribs) {
    int ti = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (ti >= nt) return;
    int lane = threadIdx.x & 31;
    GemmTask t = tasks[ti];
    bool nn = (t.flags & GEMM_NN) != 0, zero = (t.flags & GEMM_ZERO_INIT) != 0, lower = (t.flags & GEMM_LOWER) != 0;
    int total = t.m * t.n;
    for (int e = lane; e < total; e += 32) {
        int i = e % t.m, j = e / t.m;
        if (lower && j > i) continue;
        double acc = 0.0;
        for (int ci = 0; ci < t.nc; ci++) {
            GemmContrib c = contribs[t.c0 + ci];
            const double* a = c.A + i;
            if (!nn) {
                const double* bb = c.B + j;
                for (int p = 0; p < c.k; p++) acc += a[(size_t)p * c.lda] * bb[(size_t)p * c.ldb];
            } else {
                const double* bb = c.B + (size_t)j * c.ldb;
                for (int p = 0; p < c.k; p++) acc += a[(size_t)p * c.lda] * bb[p];
            }
        }
        double* p = t.C + i + (size_t)j * t.ldc;
        *p = (zero ? 0.0 : *p) - acc;
    }
}

template <int NT>
__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < NT / 32; w++) s += red[w];
    return s;
}

// ------------------------------------------------------------------------------------------------
// Merge: block copies into the (pre-zeroed) parent blocks; one warp per task.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) copy_kernel(const CopyTask* __restrict__ tasks, int nt) {
    int ti = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (ti >= nt) return;
    int lane = threadIdx.x & 31;
    CopyTask t = tasks[ti];
    if (t.src == nullptr) {
        for (int i = lane; i < t.rows; i += 32) t.dst[i + (size_t)i * t.ldd] = 1.0;
        return;
    }
    if (t.rows >= 32) {
        for (int j = 0; j < t.cols; j++)
            for (int i = lane; i < t.rows; i += 32) t.dst[i + (size_t)j * t.ldd] = t.src[i + (size_t)j * t.lds];
    } else {
        int tot = t.rows * t.cols;
        for (int e = lane; e < tot; e += 32) {
            int i = e % t.rows, j = e / t.rows;
            t.dst[i + (size_t)j * t.ldd] = t.src[i + (size_t)j * t.lds];
        }
    }
}

__global__ void fill_kernel(double* p, size_t n, double v) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

// ------------------------------------------------------------------------------------------------
// Solve-phase kernels (operations.cpp fwd/bwd), one CTA per op.
// ------------------------------------------------------------------------------------------------
constexpr int SV_T = 128;
constexpr int SV_B = 32;

// trans == 0: x <- T^-1 x with T lower (forward) ; trans == 1: x <- T^-T x with T lower (backward)
// trans == 2: x <- T^-1 x with T upper (backward substitution, PLU's U)
__global__ void __launch_bounds__(SV_T) trsv_kernel(const TrsvTask* __restrict__ tasks, int trans) {
    TrsvTask t = tasks[blockIdx.x];
    int n = t.n;
    if (n == 0) return;
    int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ double xb[SV_B];
    const double* T = t.T;
    double* x = t.x;
    int nblk = (n + SV_B - 1) / SV_B;
    if (trans == 0 && t.perm != nullptr) {  // x <- x[perm]  (P^T of ScalingPLUQ::fwd, operations.cpp)
        double tmp[32];
        for (int r = 0; r < 32; r++) {
            int i = tid + r * SV_T;
            tmp[r] = (i < n) ? x[t.perm[i]] : 0.0;
        }
        __syncthreads();
        for (int r = 0; r < 32; r++) {
            int i = tid + r * SV_T;
            if (i < n) x[i] = tmp[r];
        }
        __syncthreads();
    }
    if (trans == 0) {
        for (int b = 0; b < nblk; b++) {
            int j0 = b * SV_B, nb = min(SV_B, n - j0);
            if (warp == 0) {
                double xi = (lane < nb) ? x[j0 + lane] : 0.0;
                for (int j = 0; j < nb; j++) {
                    double xj = __shfl_sync(0xffffffffu, xi, j) / T[(j0 + j) + (size_t)(j0 + j) * t.ld];
                    if (lane == j) xi = xj;
                    if (lane > j && lane < nb) xi -= T[(j0 + lane) + (size_t)(j0 + j) * t.ld] * xj;
                }
                if (lane < nb) {
                    x[j0 + lane] = xi;
                    xb[lane] = xi;
                }
            }
            __syncthreads();
            for (int i = j0 + nb + tid; i < n; i += SV_T) {
                double s = 0.0;
                for (int j = 0; j < nb; j++) s += T[i + (size_t)(j0 + j) * t.ld] * xb[j];
                x[i] -= s;
            }
            __syncthreads();
        }
    } else if (trans == 1) {
        for (int b = nblk - 1; b >= 0; b--) {
            int j0 = b * SV_B, nb = min(SV_B, n - j0);
            // x_b -= T[j0+nb:, block]^T x[j0+nb:]  (one warp per column)
            for (int j = warp; j < nb; j += SV_T / 32) {
                double s = 0.0;
                for (int i = j0 + nb + lane; i < n; i += 32) s += T[i + (size_t)(j0 + j) * t.ld] * x[i];
                s = warp_sum(s);
                if (lane == 0) xb[j] = x[j0 + j] - s;
            }
            __syncthreads();
            if (warp == 0) {
                double xi = (lane < nb) ? xb[lane] : 0.0;
                for (int j = nb - 1; j >= 0; j--) {
                    double xj = __shfl_sync(0xffffffffu, xi, j) / T[(j0 + j) + (size_t)(j0 + j) * t.ld];
                    if (lane == j) xi = xj;
                    if (lane < j) xi -= T[(j0 + j) + (size_t)(j0 + lane) * t.ld] * xj;
                }
                if (lane < nb) x[j0 + lane] = xi;
            }
            __syncthreads();
        }
    } else {
        for (int b = nblk - 1; b >= 0; b--) {
            int j0 = b * SV_B, nb = min(SV_B, n - j0);
            if (warp == 0) {
                double xi = (lane < nb) ? x[j0 + lane] : 0.0;
                for (int j = nb - 1; j >= 0; j--) {
                    double dj = t.diag ? t.diag[j0 + j] : T[(j0 + j) + (size_t)(j0 + j) * t.ld];
                    double xj = __shfl_sync(0xffffffffu, xi, j) / dj;
                    if (lane == j) xi = xj;
                    if (lane < j) xi -= T[(j0 + lane) + (size_t)(j0 + j) * t.ld] * xj;
                }
                if (lane < nb) {
                    x[j0 + lane] = xi;
                    xb[lane] = xi;
                }
            }
            __syncthreads();
            for (int i = tid; i < j0; i += SV_T) {
                double s = 0.0;
                for (int j = 0; j < nb; j++) s += T[i + (size_t)(j0 + j) * t.ld] * xb[j];
                x[i] -= s;
            }
            __syncthreads();
        }
    }
}

// y -= sum_c A_c x_c (trans == 0, A_c is m x k)  or  y -= sum_c A_c^T x_c (trans == 1, A_c is k x m)
__global__ void __launch_bounds__(SV_T) gemv_kernel(const GemvTask* __restrict__ tasks,
                                                    const GemvContrib* __restrict__ contribs, int trans) {
    GemvTask t = tasks[blockIdx.x];
    int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (trans == 0) {
        for (int i = tid; i < t.m; i += SV_T) {
            double s = 0.0;
            for (int ci = 0; ci < t.nc; ci++) {
                GemvContrib c = contribs[t.c0 + ci];
                for (int p = 0; p < c.k; p++) s += c.A[i + (size_t)p * c.lda] * c.x[p];
            }
            t.y[i] -= s;
        }
    } else {
        for (int j = warp; j < t.m; j += SV_T / 32) {
            double s = 0.0;
            for (int ci = 0; ci < t.nc; ci++) {
                GemvContrib c = contribs[t.c0 + ci];
                for (int i = lane; i < c.k; i += 32) s += c.A[i + (size_t)j * c.lda] * c.x[i];
            }
            s = warp_sum(s);
            if (lane == 0) t.y[j] -= s;
        }
    }
}

// trans == 1: x <- Q^T x = H_{r-1} .. H_0 x ; trans == 0: x <- Q x = H_0 .. H_{r-1} x   (dormqr, one vector)
__global__ void __launch_bounds__(SV_T) house_kernel(const HouseTask* __restrict__ tasks, int trans) {
    HouseTask t = tasks[blockIdx.x];
    int tid = threadIdx.x;
    __shared__ double red[SV_T / 32];
    for (int s = 0; s < t.rank; s++) {
        int j = trans ? s : (t.rank - 1 - s);
        const double* v = t.V + (size_t)j * t.rows;
        double w = 0.0;
        for (int i = j + 1 + tid; i < t.rows; i += SV_T) w += v[i] * t.x[i];
        w = block_sum<SV_T>(w, red);
        w = (w + t.x[j]) * t.tau[j];
        __syncthreads();
        for (int i = j + 1 + tid; i < t.rows; i += SV_T) t.x[i] -= w * v[i];
        if (tid == 0) t.x[j] -= w;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(128) xcopy_kernel(const XCopyTask* __restrict__ tasks, int nt) {
    int ti = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (ti >= nt) return;
    int lane = threadIdx.x & 31;
    XCopyTask t = tasks[ti];
    for (int i = lane; i < t.n; i += 32) t.dst[i] = t.src[i];
}

// ------------------------------------------------------------------------------------------------
// PCG building blocks
// ------------------------------------------------------------------------------------------------
__global__ void spmv_kernel(int n, const int* __restrict__ rowptr, const int* __restrict__ colind,
                            const double* __restrict__ val, const double* __restrict__ x, double* __restrict__ y) {
    int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;  // 4 lanes per row (7-point stencils)
    int sub = threadIdx.x & 3;
    double s = 0.0;
    if (row < n)
        for (int k = rowptr[row] + sub; k < rowptr[row + 1]; k += 4) s += val[k] * x[colind[k]];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (row < n && sub == 0) y[row] = s;
}

__global__ void dot_kernel(int n, const double* __restrict__ a, const double* __restrict__ b, double* out) {
    __shared__ double red[8];
    double s = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += a[i] * b[i];
    s = block_sum<256>(s, red);
    if (threadIdx.x == 0) atomicAdd(out, s);
}

__global__ void axpy_kernel(int n, double alpha, const double* __restrict__ x, double* __restrict__ y) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] += alpha * x[i];
}
__global__ void xpay_kernel(int n, const double* __restrict__ x, double beta, double* __restrict__ y) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] = x[i] + beta * y[i];
}
__global__ void gather_kernel(int n, const int* __restrict__ idx, const double* __restrict__ src, double* dst) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[idx[i]];
}
__global__ void scatter_kernel(int n, const int* __restrict__ idx, const double* __restrict__ src, double* dst) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[idx[i]] = src[i];
}

inline int grid_for(size_t n, int bs, int maxb = 148 * 16) {
    size_t g = (n + bs - 1) / bs;
    if (g < 1) g = 1;
    if (g > (size_t)maxb) g = maxb;
    return (int)g;
}

}  // namespace

void launch_potrf_step(const PotrfTask* t, int nt, int j0, int* err, cudaStream_t st) {
    if (nt > 0) potrf_step_kernel<<<nt, NB, 0, st>>>(t, j0, err);
}

void launch_trsm_step(int mode, const TrsmTask* t, int nt, int j0, int max_m, cudaStream_t st) {
    if (nt <= 0 || max_m <= 0) return;
    dim3 grid(nt, (max_m + NB - 1) / NB);
    constexpr int smem = 2 * NB * LDS * sizeof(double);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(trsm_step_kernel<TRSM_RLT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(trsm_step_kernel<TRSM_LLN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(trsm_step_kernel<TRSM_RUN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(trsm_step_kernel<TRSM_LLU>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
    }
    if (mode == TRSM_RLT) trsm_step_kernel<TRSM_RLT><<<grid, NB, smem, st>>>(t, j0);
    else if (mode == TRSM_LLN) trsm_step_kernel<TRSM_LLN><<<grid, NB, smem, st>>>(t, j0);
    else if (mode == TRSM_LLU) trsm_step_kernel<TRSM_LLU><<<grid, NB, smem, st>>>(t, j0);
    else trsm_step_kernel<TRSM_RUN><<<grid, NB, smem, st>>>(t, j0);
}

void launch_getrf_small(const GetrfTask* t, int nt, int* err, cudaStream_t st) {
    if (nt > 0) getrf_small_kernel<<<nt, NB, 0, st>>>(t, err);
}
void launch_getrf_panel(const GetrfTask* t, int nt, int j0, int* err, cudaStream_t st) {
    if (nt > 0) getrf_panel_kernel<<<nt, GP_T, 0, st>>>(t, j0, err);
}
void launch_getrf_laswp(const GetrfTask* t, int nt, int j0, int max_n, cudaStream_t st) {
    if (nt <= 0 || max_n <= NB) return;
    dim3 grid(nt, (max_n + 127) / 128);
    getrf_laswp_kernel<<<grid, 128, 0, st>>>(t, j0);
}
void launch_getrf_finish(const GetrfTask* t, int nt, cudaStream_t st) {
    if (nt <= 0) return;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(getrf_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
        configured = true;
    }
    getrf_finish_kernel<<<nt, 256, 128 * 1024, st>>>(t);
}
void launch_rowperm(const RowPermTask* t, int nt, cudaStream_t st) {
    if (nt > 0) rowperm_kernel<<<nt, 128, 6144 * sizeof(double) + 64, st>>>(t);
}

void launch_gemm_tiled(const GemmTask* t, int nt, const GemmContrib* c, const int* tile_prefix, int total_tiles,
                       cudaStream_t st) {
    if (nt > 0 && total_tiles > 0) gemm_tiled_kernel<<<total_tiles, 256, 0, st>>>(t, nt, c, tile_prefix);
}

void launch_gemm_small(const GemmTask* t, int nt, const GemmContrib* c, cudaStream_t st) {
    if (nt > 0) gemm_small_kernel<<<(nt + 3) / 4, 128, 0, st>>>(t, nt, c);
}

void launch_copy(const CopyTask* t, int nt, cudaStream_t st) {
    if (nt > 0) copy_kernel<<<(nt + 3) / 4, 128, 0, st>>>(t, nt);
}

void launch_trsv(const TrsvTask* t, int nt, int trans, cudaStream_t st) {
    if (nt > 0) trsv_kernel<<<nt, SV_T, 0, st>>>(t, trans);
}
void launch_gemv(const GemvTask* t, int nt, const GemvContrib* c, int trans, cudaStream_t st) {
    if (nt > 0) gemv_kernel<<<nt, SV_T, 0, st>>>(t, c, trans);
}
void launch_house(const HouseTask* t, int nt, int trans, cudaStream_t st) {
    if (nt > 0) house_kernel<<<nt, SV_T, 0, st>>>(t, trans);
}
void launch_xcopy(const XCopyTask* t, int nt, cudaStream_t st) {
    if (nt > 0) xcopy_kernel<<<(nt + 3) / 4, 128, 0, st>>>(t, nt);
}
void launch_fill(double* p, size_t n, double v, cudaStream_t st) {
    if (n > 0) fill_kernel<<<grid_for(n, 256), 256, 0, st>>>(p, n, v);
}

void launch_spmv(int n, const int* rowptr, const int* colind, const double* val, const double* x, double* y,
                 cudaStream_t st) {
    if (n > 0) spmv_kernel<<<(int)(((size_t)n * 4 + 255) / 256), 256, 0, st>>>(n, rowptr, colind, val, x, y);
}
void launch_dot(int n, const double* a, const double* b, double* out, cudaStream_t st) {
    if (n > 0) dot_kernel<<<grid_for(n, 256, 148 * 4), 256, 0, st>>>(n, a, b, out);
}
void launch_axpy(int n, double alpha, const double* x, double* y, cudaStream_t st) {
    if (n > 0) axpy_kernel<<<grid_for(n, 256), 256, 0, st>>>(n, alpha, x, y);
}
void launch_xpay(int n, const double* x, double beta, double* y, cudaStream_t st) {
    if (n > 0) xpay_kernel<<<grid_for(n, 256), 256, 0, st>>>(n, x, beta, y);
}
void launch_gather(int n, const int* idx, const double* src, double* dst, cudaStream_t st) {
    if (n > 0) gather_kernel<<<grid_for(n, 256), 256, 0, st>>>(n, idx, src, dst);
}
void launch_scatter(int n, const int* idx, const double* src, double* dst, cudaStream_t st) {
    if (n > 0) scatter_kernel<<<grid_for(n, 256), 256, 0, st>>>(n, idx, src, dst);
}

}  // namespace spand
