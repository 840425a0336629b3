// sm_100a kernels for the spaND factorization hot path. See kernels.cuh for the mapping to the
// reference's BLAS/LAPACK call sites. All matrices are FP64, column-major.
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cstdio>
#include <cstdlib>

#include "kernels.cuh"

namespace spand {

namespace {

constexpr int LDS = NB + 1;  // padded leading dimension of 64x64 shared tiles

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// a / d given rinv = RN(1 / d): product, exact remainder (one FMA), one correction (Markstein's division step). Three
// dependent operations instead of the ~25-instruction IEEE division sequence on the substitution chain; the result is
// the correctly rounded quotient except in rare cases that are off by one ulp.
__device__ __forceinline__ double div_by(double a, double d, double rinv) {
    const double q = a * rinv;
    const double r = fma(-d, q, a);
    return fma(r, rinv, q);
}

// ------------------------------------------------------------------------------------------------
// POTRF, one 64x64 diagonal block per CTA (left-looking, one thread per row).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void potrf_tile64(double* A, int ld, int nb, int* err) {
    __shared__ double S[NB * LDS];
    int i = threadIdx.x;
    for (int j = 0; j < nb; j++)
        if (i < nb && i >= j) S[j * LDS + i] = A[i + (size_t)j * ld];
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
        if (i < nb && i >= j) A[i + (size_t)j * ld] = S[j * LDS + i];
    __syncthreads();
}

__global__ void __launch_bounds__(NB) potrf_step_kernel(const PotrfTask* __restrict__ tasks, int j0, int* err) {
    PotrfTask t = tasks[blockIdx.x];
    int nb = min(NB, t.n - j0);
    if (nb <= 0) return;
    potrf_tile64(t.A + j0 + (size_t)j0 * t.ld, t.ld, nb, err);
}

// ------------------------------------------------------------------------------------------------
// TRSM against one 64x64 diagonal block of the triangle; CTA = (task, 64-wide strip of the free dim).
// ------------------------------------------------------------------------------------------------
// One (64-wide step of the triangle) x (64-wide strip of the free dimension). T points at the diagonal block, B at
// the strip (LLN/LLU: nb x fw, rows = unknowns; RLT/RUN: fw x nb, columns = unknowns). 64 threads.
template <int MODE>
__device__ __forceinline__ void trsm_tile64(const double* T, int ldt, const double* diag, double* B, int ldb, int fw,
                                            int nb, double* trsm_smem) {
    double* Ts = trsm_smem;
    double* Xs = trsm_smem + NB * LDS;
    int tid = threadIdx.x;
    // Ts[p*LDS + j] = coefficient multiplying unknown p in equation j (p <= j)
    if (MODE == TRSM_RUN) {
        // U[p][j], p <= j : column j contiguous in p
        for (int j = 0; j < nb; j++)
            if (tid <= j) Ts[tid * LDS + j] = T[tid + (size_t)j * ldt];
    } else {
        // L[j][p], p <= j : column p contiguous in j
        for (int p = 0; p < nb; p++)
            if (tid >= p && tid < nb) Ts[p * LDS + tid] = T[tid + (size_t)p * ldt];
    }
    if (MODE == TRSM_LLU) {
        if (tid < nb) Ts[tid * LDS + tid] = 1.0;
    } else if (diag != nullptr) {
        if (tid < nb) Ts[tid * LDS + tid] = diag[tid];
    }
    if (MODE == TRSM_LLN || MODE == TRSM_LLU) {
        // X = B[0:nb, 0:fw]; Xs[c*LDS + i]
        for (int c = 0; c < fw; c++)
            if (tid < nb) Xs[c * LDS + tid] = B[tid + (size_t)c * ldb];
        __syncthreads();
        if (tid < fw) {
            double* x = Xs + tid * LDS;
            for (int i = 0; i < nb; i++) {
                const double d = Ts[i * LDS + i], rd = 1.0 / d;  // reciprocal off the dependent chain (div_by)
                double v = x[i];
                for (int p = 0; p < i; p++) v -= Ts[p * LDS + i] * x[p];
                x[i] = div_by(v, d, rd);
            }
        }
        __syncthreads();
        for (int c = 0; c < fw; c++)
            if (tid < nb) B[tid + (size_t)c * ldb] = Xs[c * LDS + tid];
    } else {
        // X = B[0:fw, 0:nb]; Xs[j*LDS + r]
        for (int j = 0; j < nb; j++)
            if (tid < fw) Xs[j * LDS + tid] = B[tid + (size_t)j * ldb];
        __syncthreads();
        if (tid < fw) {
            for (int j = 0; j < nb; j++) {
                const double d = Ts[j * LDS + j], rd = 1.0 / d;
                double v = Xs[j * LDS + tid];
                for (int p = 0; p < j; p++) v -= Xs[p * LDS + tid] * Ts[p * LDS + j];
                Xs[j * LDS + tid] = div_by(v, d, rd);
            }
            for (int j = 0; j < nb; j++) B[tid + (size_t)j * ldb] = Xs[j * LDS + tid];
        }
    }
    __syncthreads();
}

template <int MODE>
__global__ void __launch_bounds__(NB) trsm_step_kernel(const TrsmTask* __restrict__ tasks, int j0) {
    TrsmTask t = tasks[blockIdx.x];
    int nb = min(NB, t.n - j0);
    int f0 = blockIdx.y * NB;  // offset in the free dimension
    if (nb <= 0 || f0 >= t.m) return;
    int fw = min(NB, t.m - f0);
    extern __shared__ double trsm_smem[];
    const double* T = t.T + j0 + (size_t)j0 * t.ldt;
    double* B = (MODE == TRSM_LLN || MODE == TRSM_LLU) ? t.B + j0 + (size_t)f0 * t.ldb : t.B + f0 + (size_t)j0 * t.ldb;
    trsm_tile64<MODE>(T, t.ldt, t.diag ? t.diag + j0 : nullptr, B, t.ldb, fw, nb, trsm_smem);
}


// ------------------------------------------------------------------------------------------------
// GETRF with partial pivoting (dgetrf semantics: first maximal |a| wins) + split_LU (src/util.cpp:213-227).
// ------------------------------------------------------------------------------------------------

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
    // split_LU (src/util.cpp:213-227); the square roots and reciprocals once per row, not once per entry
    double di = 1.0, sqi = 1.0;
    if (i < n) {
        di = S[i * LDS + i];
        sqi = sqrt(fabs(di));
        dd[i] = sqi;
        t.perm[i] = prm[i];
    }
    __syncthreads();
    if (i < n) {
        const double si = (di > 0 ? 1.0 : (di < 0 ? -1.0 : 0.0)) * sqi;
        const double ri = 1.0 / di;
        t.ud[i] = si;
        for (int j = 0; j < n; j++) {
            const double a = S[j * LDS + i];
            double o;
            if (i > j) o = a * dd[j];
            else if (i < j) o = si * (ri * a);
            else o = sqi;
            t.A[i + (size_t)j * t.ld] = o;
        }
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
        // a thread owns rows: multiplier of the row, then its part of the rank-1 update of the panel (coalesced over
        // the threads for every column, independent updates along the row, no index arithmetic in the loop)
        const int ncc = nbp - c - 1;
        const double* rk = rowk + c + 1;
        for (int r = col + 1 + tid; r < n; r += GP_T) {
            double* a = A + r + col * ld;
            const double l = *a / akk;
            *a = l;
            a += ld;
            int cc = 0;
            for (; cc + 3 < ncc; cc += 4) {
                const double x0 = a[0] - l * rk[cc], x1 = a[ld] - l * rk[cc + 1];
                const double x2 = a[2 * ld] - l * rk[cc + 2], x3 = a[3 * ld] - l * rk[cc + 3];
                a[0] = x0;
                a[ld] = x1;
                a[2 * ld] = x2;
                a[3 * ld] = x3;
                a += 4 * ld;
            }
            for (; cc < ncc; cc++, a += ld) *a -= l * rk[cc];
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
    // n: the LAPACK diagonal | n: |d|^1/2 | n: 1 / d | 2 n ints: permutation, swap sequence (32 n bytes = 128 KB at n = 4096)
    extern __shared__ double dfin[];
    double* sq = dfin + n;
    double* rinv = sq + n;
    int* prm = reinterpret_cast<int*>(rinv + n);
    int* swp = prm + n;  // the LAPACK swap sequence (32 n bytes in all)
    const int tid = threadIdx.x;
    for (int i = tid; i < n; i += 256) {
        const double di = t.A[i + (size_t)i * t.ld];
        dfin[i] = di;
        sq[i] = sqrt(fabs(di));
        rinv[i] = 1.0 / di;
        prm[i] = i;
        swp[i] = t.ipiv[i];
    }
    __syncthreads();
    if (tid == 0) {  // swap2perm (src/util.cpp:76-88), a serial chain: on shared memory, not on global
        for (int i = 0; i < n; i++) {
            const int p = swp[i];
            const int tmp = prm[p];
            prm[p] = prm[i];
            prm[i] = tmp;
        }
    }
    __syncthreads();
    for (int i = tid; i < n; i += 256) {
        const double di = dfin[i];
        t.ud[i] = (di > 0 ? 1.0 : (di < 0 ? -1.0 : 0.0)) * sq[i];
        t.perm[i] = prm[i];
    }
    // split_LU (src/util.cpp:213-227): L(i,j) = l_ij |d_j|^1/2 ; U(i,j) = s_i ((1 / d_i) u_ij), s_i = sign(d_i) |d_i|^1/2
    for (size_t e = tid; e < (size_t)n * n; e += 256) {
        const int i = (int)(e % n), j = (int)(e / n);
        double* p = t.A + i + (size_t)j * t.ld;
        const double a = *p;
        if (i > j) *p = a * sq[j];
        else if (i < j) {
            const double di = dfin[i];
            const double si = (di > 0 ? 1.0 : (di < 0 ? -1.0 : 0.0)) * sq[i];
            *p = si * (rinv[i] * a);
        } else *p = sq[i];
    }
}

__global__ void __launch_bounds__(128) rowperm_kernel(const RowPermTask* __restrict__ tasks) {
    RowPermTask t = tasks[blockIdx.x];
    extern __shared__ double pbuf[];  // n x w strip
    const int n = t.n;
    if (n == 0 || t.m == 0) return;
    const int w = max(1, min(t.m, 6144 / n));
    for (int c0 = blockIdx.y * w; c0 < t.m; c0 += gridDim.y * w) {  // column strips of a block over blockIdx.y
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

// One 64x64 tile of one target; contributions come through fetch(ci) in a fixed order (deterministic sums).
// Software pipeline over the k chunks of all contributions: the global loads of chunk i + 1 are in flight (registers)
// while the tensor cores work on chunk i out of one of two shared-memory buffers; one block barrier per chunk.
// NBUF = 1 (callers that need the shared memory for themselves): one buffer, a second barrier per chunk.
template <int NBUF = 2, class Fetch>
__device__ __forceinline__ void gemm_tile64(double* C, int ldc, int m, int n, int flags, int row0, int col0, int nc,
                                            Fetch fetch) {
    const bool nn = (flags & GEMM_NN) != 0;
    __shared__ double As[NBUF][GK * GLD];
    __shared__ double Bs[NBUF][GK * GLD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2;
    double acc[2][4][2];
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[a][c][0] = acc[a][c][1] = 0.0;

    const int arow = tid & 63, ak = tid >> 6;  // A loader: rows contiguous
    const int bk_nn = tid & 15, bj_nn = tid >> 4;

    int ci = 0, k0 = 0;
    GemmContrib c{};
    bool valid = false;
    auto seek = [&]() {  // first chunk at or after (ci, k0) that exists
        valid = false;
        while (ci < nc) {
            if (k0 == 0) c = fetch(ci);
            if (k0 < c.k) {
                valid = true;
                return;
            }
            ci++;
            k0 = 0;
        }
    };
    double ra[4], rb[4];
    auto load = [&]() {
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const int kk = k0 + ak + 4 * s;
            const int gi = row0 + arow;
            ra[s] = (gi < m && kk < c.k) ? c.A[gi + (size_t)kk * c.lda] : 0.0;
            if (!nn) {
                const int gj = col0 + arow;
                rb[s] = (gj < n && kk < c.k) ? c.B[gj + (size_t)kk * c.ldb] : 0.0;
            } else {
                const int kk2 = k0 + bk_nn, gj = col0 + bj_nn + 16 * s;
                rb[s] = (gj < n && kk2 < c.k) ? c.B[kk2 + (size_t)gj * c.ldb] : 0.0;
            }
        }
    };
    seek();
    if (valid) load();
    int buf = 0;
    while (valid) {
        double* as = As[buf];
        double* bs = Bs[buf];
#pragma unroll
        for (int s = 0; s < 4; s++) {
            as[(ak + 4 * s) * GLD + arow] = ra[s];
            if (!nn) bs[(ak + 4 * s) * GLD + arow] = rb[s];
            else bs[bk_nn * GLD + bj_nn + 16 * s] = rb[s];
        }
        __syncthreads();
        k0 += GK;
        if (k0 >= c.k) {
            ci++;
            k0 = 0;
        }
        seek();
        if (valid) load();  // in flight during the tensor-core work below
#pragma unroll
        for (int kk = 0; kk < GK; kk += 4) {
            double fa[2], fb[4];
            const int kr = (kk + (lane & 3)) * GLD;
#pragma unroll
            for (int mi = 0; mi < 2; mi++) fa[mi] = as[kr + wm * 16 + mi * 8 + (lane >> 2)];
#pragma unroll
            for (int ni = 0; ni < 4; ni++) fb[ni] = bs[kr + wn * 32 + ni * 8 + (lane >> 2)];
#pragma unroll
            for (int mi = 0; mi < 2; mi++)
#pragma unroll
                for (int ni = 0; ni < 4; ni++) dmma(acc[mi][ni][0], acc[mi][ni][1], fa[mi], fb[ni]);
        }
        if (NBUF == 2) buf ^= 1;
        else __syncthreads();
    }
    bool zero = (flags & GEMM_ZERO_INIT) != 0, lower = (flags & GEMM_LOWER) != 0, pos = (flags & GEMM_POS) != 0;
#pragma unroll
    for (int mi = 0; mi < 2; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                int gi = row0 + wm * 16 + mi * 8 + (lane >> 2);
                int gj = col0 + wn * 32 + ni * 8 + 2 * (lane & 3) + e;
                if (gi < m && gj < n && (!lower || gi >= gj)) {
                    double* p = C + gi + (size_t)gj * ldc;
                    const double v = zero ? 0.0 : *p;
                    *p = pos ? v + acc[mi][ni][e] : v - acc[mi][ni][e];
                }
            }
    __syncthreads();
}

__global__ void __launch_bounds__(256) gemm_tiled_kernel(const GemmTask* __restrict__ tasks, int nt,
                                                         const GemmContrib* __restrict__ contribs,
                                                         const int* __restrict__ tile_prefix,
                                                         const int* __restrict__ tile_task) {
    int b = blockIdx.x;
    int lo = 0, hi = nt - 1;
    if (tile_task != nullptr) lo = tile_task[b];
    else
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
    const GemmContrib* cc = contribs + t.c0;
    // triangular operand: the inner index of this tile stops where the triangle ends
    const int klim = (t.flags & GEMM_TRIB) ? (tile_c + 1) * GT : ((t.flags & GEMM_TRIA) ? (tile_r + 1) * GT : INT_MAX);
    gemm_tile64(t.C, t.ldc, t.m, t.n, t.flags, tile_r * GT, tile_c * GT, t.nc, [cc, klim](int ci) {
        GemmContrib c = cc[ci];
        c.k = min(c.k, klim);
        return c;
    });
}

__global__ void __launch_bounds__(256) ut_kernel(const UtTask* __restrict__ tasks) {
    const UtTask t = tasks[blockIdx.x];
    const size_t total = (size_t)t.n * t.n;
    for (size_t x = (size_t)blockIdx.y * blockDim.x + threadIdx.x; x < total; x += (size_t)gridDim.y * blockDim.x) {
        const int p = (int)(x % t.n), i = (int)(x / t.n);  // reads run down a column of U
        double v = 0.0;
        if (p < i) v = t.U[p + (size_t)i * t.ldu];
        else if (p == i) v = t.ud[i];
        t.Lt[i + (size_t)p * t.n] = v;
    }
}

__global__ void __launch_bounds__(256) eye_kernel(const EyeTask* __restrict__ tasks) {
    const EyeTask t = tasks[blockIdx.x];
    const size_t total = (size_t)t.n * t.n;
    for (size_t x = (size_t)blockIdx.y * blockDim.x + threadIdx.x; x < total; x += (size_t)gridDim.y * blockDim.x) {
        const int i = (int)(x % t.n), j = (int)(x / t.n);
        t.W[i + (size_t)j * t.ld] = (i == j) ? 1.0 : 0.0;
    }
}


// Inverse of every 64 x 64 diagonal block of a lower triangle (thread c: column c by forward substitution); the
// blocks are stored 64 x 64 column-major, zero outside the triangle.
__global__ void __launch_bounds__(NB) trtri_kernel(const TrtriTask* __restrict__ tasks) {
    const TrtriTask t = tasks[blockIdx.x];
    const int j0 = blockIdx.y * NB;
    if (j0 >= t.n) return;
    const int nb = min(NB, t.n - j0);
    __shared__ double S[NB * LDS];  // S[p * LDS + i] = L(i, p)
    const int c = threadIdx.x;
    const double* T = t.T + j0 + (size_t)j0 * t.ldt;
    for (int p = 0; p < nb; p++)
        if (c >= p && c < nb) S[p * LDS + c] = T[c + (size_t)p * t.ldt];
    __syncthreads();
    const bool compact = t.ldw > 0;  // one n x n block (n <= 64) with leading dimension ldw
    double* out = compact ? t.inv + (size_t)c * t.ldw : t.inv + (size_t)blockIdx.y * NB * NB + (size_t)c * NB;
    double x[NB];
#pragma unroll
    for (int i = 0; i < NB; i++) x[i] = 0.0;
    if (c < nb) {
#pragma unroll
        for (int i = 0; i < NB; i++) {
            if (i >= c && i < nb) {
                double v = (i == c) ? 1.0 : 0.0;
#pragma unroll
                for (int p = 0; p < NB; p++)
                    if (p >= c && p < i) v -= S[p * LDS + i] * x[p];
                const double d = S[i * LDS + i];
                x[i] = div_by(v, d, 1.0 / d);  // the reciprocal does not depend on the chain
            }
        }
    }
    if (compact) {
        if (c < nb) {
#pragma unroll
            for (int i = 0; i < NB; i++)
                if (i < nb) out[i] = x[i];
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < NB; i++) out[i] = x[i];
}

template <int MODE>
__global__ void __launch_bounds__(256) trsm_strip_kernel(const TrsmTask* __restrict__ tasks, int nt,
                                                         const int* __restrict__ strip_prefix) {
    int b = blockIdx.x;
    int lo = 0, hi = nt - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (strip_prefix[mid] <= b) lo = mid;
        else hi = mid - 1;
    }
    const TrsmTask t = tasks[lo];
    const int f0 = (b - strip_prefix[lo]) * NB;
    const int fw = min(NB, t.m - f0);
    const int nblk = (t.n + NB - 1) / NB;
    // block lower triangular B (LLN): rows above the strip's own block are zero and stay zero
    const int jfirst = (MODE == TRSM_LLN && t.tri) ? f0 / NB : 0;
    const int r0 = jfirst * NB;
    for (int j = jfirst; j < nblk; j++) {
        const int j0 = j * NB, nb = min(NB, t.n - j0);
        GemmContrib c;
        const double* inv = t.inv + (size_t)j * NB * NB;
        if (MODE == TRSM_RLT) {
            // X_j = (B_j - X_{0:j} L_{j,0:j}^T) inv(L_jj)^T on the row strip [f0, f0 + fw)
            double* Bs = t.B + f0;
            double* C = Bs + (size_t)j0 * t.ldb;
            if (j > 0) {
                c = GemmContrib{Bs, t.T + j0, t.ldb, t.ldt, j0};
                gemm_tile64<1>(C, t.ldb, fw, nb, 0, 0, 0, 1, [c](int) { return c; });
            }
            c = GemmContrib{C, inv, t.ldb, NB, nb};
            gemm_tile64<1>(C, t.ldb, fw, nb, GEMM_ZERO_INIT | GEMM_POS, 0, 0, 1, [c](int) { return c; });
        } else {
            // X_j = inv(L_jj) (B_j - L_{j,0:j} X_{0:j}) on the column strip [f0, f0 + fw)
            double* Bs = t.B + (size_t)f0 * t.ldb;
            double* C = Bs + j0;
            if (j > jfirst) {
                c = GemmContrib{t.T + j0 + (size_t)r0 * t.ldt, Bs + r0, t.ldt, t.ldb, j0 - r0};
                gemm_tile64<1>(C, t.ldb, nb, fw, GEMM_NN, 0, 0, 1, [c](int) { return c; });
            }
            c = GemmContrib{inv, C, NB, t.ldb, nb};
            gemm_tile64<1>(C, t.ldb, nb, fw, GEMM_NN | GEMM_ZERO_INIT | GEMM_POS, 0, 0, 1, [c](int) { return c; });
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
__global__ void __launch_bounds__(SV_T) trsv_kernel(const TrsvTask* __restrict__ tasks, int trans, int skip_small) {
    TrsvTask t = tasks[blockIdx.x];
    int n = t.n;
    if (n == 0 || (skip_small && n <= 32)) return;  // n <= 32: trsv_small_kernel
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
                // the 32 x 32 diagonal block goes to registers first (lane i holds row i): no load on the chain
                double tr[SV_B], dg = 1.0;
#pragma unroll
                for (int j = 0; j < SV_B; j++) {
                    tr[j] = (lane < nb && j < nb) ? T[(j0 + lane) + (size_t)(j0 + j) * t.ld] : 0.0;
                    if (lane == j) dg = (j < nb) ? tr[j] : 1.0;
                }
                double xi = (lane < nb) ? x[j0 + lane] : 0.0;
                const double rg = 1.0 / dg;
#pragma unroll
                for (int j = 0; j < SV_B; j++) {
                    if (j < nb) {
                        double xj = div_by(__shfl_sync(0xffffffffu, xi, j), __shfl_sync(0xffffffffu, dg, j), __shfl_sync(0xffffffffu, rg, j));
                        if (lane == j) xi = xj;
                        if (lane > j && lane < nb) xi -= tr[j] * xj;
                    }
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
                double tr[SV_B], dg = 1.0;  // lane i holds column i of the diagonal block
#pragma unroll
                for (int j = 0; j < SV_B; j++) {
                    tr[j] = (lane < nb && j < nb) ? T[(j0 + j) + (size_t)(j0 + lane) * t.ld] : 0.0;
                    if (lane == j) dg = (j < nb) ? tr[j] : 1.0;
                }
                double xi = (lane < nb) ? xb[lane] : 0.0;
                const double rg = 1.0 / dg;
#pragma unroll
                for (int j = SV_B - 1; j >= 0; j--) {
                    if (j < nb) {
                        double xj = div_by(__shfl_sync(0xffffffffu, xi, j), __shfl_sync(0xffffffffu, dg, j), __shfl_sync(0xffffffffu, rg, j));
                        if (lane == j) xi = xj;
                        if (lane < j) xi -= tr[j] * xj;
                    }
                }
                if (lane < nb) x[j0 + lane] = xi;
            }
            __syncthreads();
        }
    } else {
        for (int b = nblk - 1; b >= 0; b--) {
            int j0 = b * SV_B, nb = min(SV_B, n - j0);
            if (warp == 0) {
                double tr[SV_B], dg = 1.0;  // lane i holds row i of the diagonal block
#pragma unroll
                for (int j = 0; j < SV_B; j++) {
                    tr[j] = (lane < nb && j < nb) ? T[(j0 + lane) + (size_t)(j0 + j) * t.ld] : 0.0;
                    if (lane == j) dg = (j < nb) ? tr[j] : 1.0;
                }
                if (t.diag && lane < nb) dg = t.diag[j0 + lane];
                double xi = (lane < nb) ? x[j0 + lane] : 0.0;
                const double rg = 1.0 / dg;
#pragma unroll
                for (int j = SV_B - 1; j >= 0; j--) {
                    if (j < nb) {
                        double xj = div_by(__shfl_sync(0xffffffffu, xi, j), __shfl_sync(0xffffffffu, dg, j), __shfl_sync(0xffffffffu, rg, j));
                        if (lane == j) xi = xj;
                        if (lane < j) xi -= tr[j] * xj;
                    }
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

// Triangles of at most 32 rows (the bulk of the low levels): one warp per task, the triangle preloaded into
// registers (lane i holds row i, or column i for the transposed solve) so that the substitution chain contains no
// memory access. Same operation order as trsv_kernel's single-block case.
__global__ void __launch_bounds__(128) trsv_small_kernel(const TrsvTask* __restrict__ tasks, int nt, int trans) {
    const int ti = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (ti >= nt) return;
    const TrsvTask t = tasks[ti];
    const int n = t.n, lane = threadIdx.x & 31;
    if (n == 0 || n > 32) return;
    const double* T = t.T;
    double tr[32];
#pragma unroll
    for (int j = 0; j < 32; j++) {
        double v = 0.0;
        if (lane < n && j < n) v = (trans == 1) ? T[j + (size_t)lane * t.ld] : T[lane + (size_t)j * t.ld];
        tr[j] = v;
    }
    double xi = 0.0;
    if (lane < n) xi = (trans == 0 && t.perm != nullptr) ? t.x[t.perm[lane]] : t.x[lane];
    double dg = 1.0;  // diagonal entry of the own row
#pragma unroll
    for (int j = 0; j < 32; j++)
        if (lane == j) dg = tr[j];
    if (trans == 2 && t.diag != nullptr && lane < n) dg = t.diag[lane];
    const double rg = 1.0 / dg;  // reciprocal of the own diagonal entry, off the substitution chain (div_by)
    __syncwarp();
    if (trans == 0) {
#pragma unroll
        for (int j = 0; j < 32; j++) {
            if (j < n) {
                const double xj = div_by(__shfl_sync(0xffffffffu, xi, j), __shfl_sync(0xffffffffu, dg, j), __shfl_sync(0xffffffffu, rg, j));
                if (lane == j) xi = xj;
                if (lane > j) xi -= tr[j] * xj;
            }
        }
    } else {
#pragma unroll
        for (int j = 31; j >= 0; j--) {
            if (j < n) {
                const double xj = div_by(__shfl_sync(0xffffffffu, xi, j), __shfl_sync(0xffffffffu, dg, j), __shfl_sync(0xffffffffu, rg, j));
                if (lane == j) xi = xj;
                if (lane < j) xi -= tr[j] * xj;
            }
        }
    }
    if (lane < n) t.x[lane] = xi;
}

// Large operands (top levels: few tasks, hundreds of rows and columns each): the rows (forward) / columns
// (transposed) of a task are spread over blockIdx.y, and inside a CTA the inner dimension is split over the 8 warps
// with several independent loads in flight per lane; partial sums are combined in a fixed order (deterministic).
constexpr int GV_T = 256;
__global__ void __launch_bounds__(GV_T) gemv_n_big_kernel(const GemvTask* __restrict__ tasks,
                                                         const GemvContrib* __restrict__ contribs) {
    const GemvTask t = tasks[blockIdx.x];
    const int r0 = blockIdx.y * 64;
    if (r0 >= t.m) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ double part[GV_T / 32][64];
    const int i0 = r0 + lane, i1 = r0 + 32 + lane;
    const bool v0 = i0 < t.m, v1 = i1 < t.m;
    const int j0 = v0 ? i0 : r0, j1 = v1 ? i1 : r0;  // clamped row indices: loads stay in range, results are masked
    double s0 = 0.0, s1 = 0.0;
    for (int ci = 0; ci < t.nc; ci++) {
        const GemvContrib c = contribs[t.c0 + ci];
        int p = warp;
        for (; p + 24 < c.k; p += 32) {
            double a0[4], a1[4], xv[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const double* col = c.A + (size_t)(p + 8 * q) * c.lda;
                a0[q] = col[j0];
                a1[q] = col[j1];
                xv[q] = c.x[p + 8 * q];
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                s0 = fma(a0[q], xv[q], s0);
                s1 = fma(a1[q], xv[q], s1);
            }
        }
        for (; p < c.k; p += 8) {
            const double* col = c.A + (size_t)p * c.lda;
            const double xv = c.x[p];
            s0 = fma(col[j0], xv, s0);
            s1 = fma(col[j1], xv, s1);
        }
    }
    part[warp][lane] = s0;
    part[warp][lane + 32] = s1;
    __syncthreads();
    if (tid < 64 && r0 + tid < t.m) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < GV_T / 32; w++) s += part[w][tid];
        t.y[r0 + tid] -= s;
    }
}

__global__ void __launch_bounds__(GV_T) gemv_t_big_kernel(const GemvTask* __restrict__ tasks,
                                                         const GemvContrib* __restrict__ contribs) {
    const GemvTask t = tasks[blockIdx.x];
    const int c0 = blockIdx.y * 32;
    if (c0 >= t.m) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int jb = c0 + warp * 4;  // this warp's four output columns
    if (jb >= t.m) return;
    int jj[4];
#pragma unroll
    for (int q = 0; q < 4; q++) jj[q] = min(jb + q, t.m - 1);
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int ci = 0; ci < t.nc; ci++) {
        const GemvContrib c = contribs[t.c0 + ci];
        const double* a[4];
#pragma unroll
        for (int q = 0; q < 4; q++) a[q] = c.A + (size_t)jj[q] * c.lda;
        int i = lane;
        for (; i + 32 < c.k; i += 64) {
            const double x0 = c.x[i], x1 = c.x[i + 32];
            double b0[4], b1[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                b0[q] = a[q][i];
                b1[q] = a[q][i + 32];
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                s[q] = fma(b0[q], x0, s[q]);
                s[q] = fma(b1[q], x1, s[q]);
            }
        }
        if (i < c.k) {
            const double x0 = c.x[i];
#pragma unroll
            for (int q = 0; q < 4; q++) s[q] = fma(a[q][i], x0, s[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < 4; q++) s[q] = warp_sum(s[q]);
    if (lane < 4 && jb + lane < t.m) {
        double v = s[0];
        if (lane == 1) v = s[1];
        if (lane == 2) v = s[2];
        if (lane == 3) v = s[3];
        t.y[jb + lane] -= v;
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
    if (t.rank >= t.rows) return;  // not sparsified: no Orthogonal op was recorded (src/tree.cpp:1317-1319)
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

// Same operation with the vector held in registers (row i on thread i mod NT): a reflector costs one pass over its
// column of V and ONE barrier (the exchange of the per-warp partial dot products, double buffered), instead of four.
// WARP == true: tasks of at most 64 rows, one warp each, no barrier at all.
template <int NT, int R, bool WARP>
__global__ void __launch_bounds__(WARP ? 128 : NT) house_reg_kernel(const HouseTask* __restrict__ tasks, int nt, int trans) {
    const int ti = WARP ? (int)(blockIdx.x * 4 + (threadIdx.x >> 5)) : (int)blockIdx.x;
    if (ti >= nt) return;
    const HouseTask t = tasks[ti];
    if (t.rank >= t.rows) return;
    const int tid = WARP ? (int)(threadIdx.x & 31) : (int)threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    __shared__ double part[2][NT / 32];
    double xr[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int i = tid + r * NT;
        xr[r] = (i < t.rows) ? t.x[i] : 0.0;
    }
    for (int s = 0; s < t.rank; s++) {
        const int j = trans ? s : (t.rank - 1 - s);
        const double* v = t.V + (size_t)j * t.rows;
        const double tau = t.tau[j];
        double vr[R], w = 0.0;
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int i = tid + r * NT;
            vr[r] = (i > j && i < t.rows) ? v[i] : (i == j ? 1.0 : 0.0);
            w = fma(vr[r], xr[r], w);
        }
        w = warp_sum(w);
        if (!WARP) {
            if (lane == 0) part[s & 1][warp] = w;
            __syncthreads();
            w = 0.0;
#pragma unroll
            for (int q = 0; q < NT / 32; q++) w += part[s & 1][q];
        }
        w *= tau;
#pragma unroll
        for (int r = 0; r < R; r++) xr[r] = fma(-w, vr[r], xr[r]);
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int i = tid + r * NT;
        if (i < t.rows) t.x[i] = xr[r];
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


// ------------------------------------------------------------------------------------------------
// Plan-driven batches (see kernels.cuh): ids resolved through DevTables, all dimensions <= SMALL_DIM.
// ------------------------------------------------------------------------------------------------
constexpr int WLD = 33;  // row stride of the per-warp 32 x 32 tiles
constexpr unsigned FULLM = 0xffffffffu;

__device__ __forceinline__ void push_mid(int* mid, int* cnt, int ti) { mid[atomicAdd(cnt, 1)] = ti; }

__global__ void __launch_bounds__(128) potrf_sym_kernel(DevTables T, const int* __restrict__ clusters,
                                                        const int* __restrict__ piv, int nt, int* mid, int* cnt, int* err) {
    __shared__ double Sw[4][32 * WLD];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ti = blockIdx.x * 4 + w;
    if (ti >= nt) return;
    if (T.rank >= 0 && T.owner[clusters[ti]] != T.rank) return;
    const int n = T.csize[clusters[ti]];
    if (n <= 0 || n > SMALL_DIM) return;
    if (n > 32) {
        if (lane == 0) push_mid(mid, cnt, ti);
        return;
    }
    const int e = piv[ti];
    double* A = T.eptr[e];
    const int ld = T.eld[e];
    double* S = Sw[w];
    for (int x = lane; x < n * n; x += 32) {
        int i = x % n, j = x / n;
        if (i >= j) S[j * WLD + i] = A[i + (size_t)j * ld];
    }
    __syncwarp();
    for (int k = 0; k < n; k++) {
        double v = 0.0;
        if (lane >= k && lane < n) {
            v = S[k * WLD + lane];
            for (int p = 0; p < k; p++) v -= S[p * WLD + lane] * S[p * WLD + k];
        }
        const double d = __shfl_sync(FULLM, v, k);
        const double r = sqrt(d);
        if (lane == k) {
            if (!(d > 0.0)) atomicOr(err, 1);
            S[k * WLD + k] = r;
        } else if (lane > k && lane < n) S[k * WLD + lane] = v / r;
        __syncwarp();
    }
    for (int x = lane; x < n * n; x += 32) {
        int i = x % n, j = x / n;
        if (i >= j) A[i + (size_t)j * ld] = S[j * WLD + i];
    }
}

__global__ void __launch_bounds__(NB) potrf_mid_kernel(DevTables T, const int* __restrict__ clusters,
                                                       const int* __restrict__ piv, const int* __restrict__ mid,
                                                       const int* __restrict__ cnt, int* err) {
    const int nmid = *cnt;
    for (int q = blockIdx.x; q < nmid; q += gridDim.x) {
        const int ti = mid[q];
        const int e = piv[ti];
        potrf_tile64(T.eptr[e], T.eld[e], T.csize[clusters[ti]], err);
    }
}

// In-warp triangular solves on a tile Bs[col * WLD + row] (rows x cols); the triangle is read from global memory
// (uniform addresses: one broadcast transaction per load, L1-resident).
template <bool FASTDIV>
__device__ __forceinline__ void warp_solve_right_lt(double* Bs, int rows, int cols, const double* L, int ldl, int lane) {
    // B <- B L^-T : lane = row of B
    if (lane < rows)
        for (int j = 0; j < cols; j++) {
            const double d = L[j + (size_t)j * ldl];
            const double rinv = FASTDIV ? 1.0 / d : 0.0;  // off the dependent chain
            double v = Bs[j * WLD + lane];
            for (int p = 0; p < j; p++) v -= Bs[p * WLD + lane] * L[j + (size_t)p * ldl];
            Bs[j * WLD + lane] = FASTDIV ? div_by(v, d, rinv) : v / d;
        }
}
template <bool FASTDIV>
__device__ __forceinline__ void warp_solve_left_ln(double* Bs, int rows, int cols, const double* L, int ldl, int lane) {
    // B <- L^-1 B : lane = column of B
    if (lane < cols) {
        double* x = Bs + lane * WLD;
        for (int i = 0; i < rows; i++) {
            const double d = L[i + (size_t)i * ldl];
            const double rinv = FASTDIV ? 1.0 / d : 0.0;
            double v = x[i];
            for (int p = 0; p < i; p++) v -= L[i + (size_t)p * ldl] * x[p];
            x[i] = FASTDIV ? div_by(v, d, rinv) : v / d;
        }
    }
}
__device__ __forceinline__ void warp_tile_load(double* Bs, const double* B, int ldb, int rows, int cols, int lane) {
    for (int x = lane; x < rows * cols; x += 32) {
        int i = x % rows, j = x / rows;
        Bs[j * WLD + i] = B[i + (size_t)j * ldb];
    }
}
__device__ __forceinline__ void warp_tile_store(const double* Bs, double* B, int ldb, int rows, int cols, int lane) {
    for (int x = lane; x < rows * cols; x += 32) {
        int i = x % rows, j = x / rows;
        B[i + (size_t)j * ldb] = Bs[j * WLD + i];
    }
}

// MODE: TRSM_RLT (B is cm x cn) or TRSM_LLN (B is cn x cm)
template <int MODE>
__global__ void __launch_bounds__(128) trsm_sym_kernel(DevTables T, const SymTrsm* __restrict__ tasks, int nt, int* mid,
                                                       int* cnt, int fastdiv) {
    __shared__ double Sw[4][32 * WLD];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ti = blockIdx.x * 4 + w;
    if (ti >= nt) return;
    const SymTrsm t = tasks[ti];
    if (T.rank >= 0 && T.owner[T.en1[t.eB]] != T.rank) return;
    const int m = T.csize[t.cm], n = T.csize[t.cn];
    if (m <= 0 || n <= 0 || m > SMALL_DIM || n > SMALL_DIM) return;
    if (m > 32 || n > 32) {
        if (lane == 0) push_mid(mid, cnt, ti);
        return;
    }
    double* B = T.eptr[t.eB];
    const int ldb = T.eld[t.eB];
    const double* L = T.eptr[t.eT];
    const int ldl = T.eld[t.eT];
    double* S = Sw[w];
    if (MODE == TRSM_RLT) {
        warp_tile_load(S, B, ldb, m, n, lane);
        __syncwarp();
        if (fastdiv) warp_solve_right_lt<true>(S, m, n, L, ldl, lane);
        else warp_solve_right_lt<false>(S, m, n, L, ldl, lane);
        __syncwarp();
        warp_tile_store(S, B, ldb, m, n, lane);
    } else {
        warp_tile_load(S, B, ldb, n, m, lane);
        __syncwarp();
        if (fastdiv) warp_solve_left_ln<true>(S, n, m, L, ldl, lane);
        else warp_solve_left_ln<false>(S, n, m, L, ldl, lane);
        __syncwarp();
        warp_tile_store(S, B, ldb, n, m, lane);
    }
}

template <int MODE>
__global__ void __launch_bounds__(NB) trsm_mid_kernel(DevTables T, const SymTrsm* __restrict__ tasks,
                                                      const int* __restrict__ mid, const int* __restrict__ cnt) {
    extern __shared__ double trsm_smem[];
    const int nmid = *cnt;
    for (int q = blockIdx.x; q < nmid; q += gridDim.x) {
        const SymTrsm t = tasks[mid[q]];
        trsm_tile64<MODE>(T.eptr[t.eT], T.eld[t.eT], nullptr, T.eptr[t.eB], T.eld[t.eB], T.csize[t.cm], T.csize[t.cn],
                          trsm_smem);
    }
}

// Right-looking substitution on a tile Bs (64 x 64, Bs[col * LDS + row]) by the 8 warps of a 256-thread CTA against the
// lower triangle Ls[p * LDS + q] = L(q, p) (zero-filled outside the triangle and beyond nt). A "unit" is a row of the
// tile for B <- B L^-T (ROWS) or a column for B <- L^-1 B: units are independent, unit u = 8 * warp + (lane >> 2)
// lives on four lanes of one warp, lane c of the four holds the entries q = 4k + c in registers. Step p: the holder of
// entry p divides by L(p, p) and shuffles x_p to its three neighbours, everybody subtracts x_p L(q, p) from its
// entries q > p. Every entry sees v <- fma(-x_p, L(q, p), v) for p = 0..q-1 in this order and one division: the
// arithmetic (and its rounding) of the row-by-row substitution it replaces (trsm_tile64), with 16 independent updates
// per thread and step instead of one dependent chain per thread.
template <bool ROWS, int NTM, bool FASTDIV>  // NTM: compile-time bound of the triangle dimension (multiple of 4)
__device__ __forceinline__ void tile_solve_rl_n(double* Bs, const double* Ls, int nt, int nu) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (8 * warp >= nu) return;  // no unit on this warp
    const int u = 8 * warp + (lane >> 2), c = lane & 3;
    constexpr int NK = NTM / 4;
    double v[NK];
#pragma unroll
    for (int k = 0; k < NK; k++) {
        const int q = 4 * k + c;
        v[k] = (q < nt && u < nu) ? (ROWS ? Bs[q * LDS + u] : Bs[u * LDS + q]) : 0.0;
    }
#pragma unroll
    for (int p = 0; p < NTM; p++) {
        if (p < nt) {  // uniform
            // FASTDIV: 1 / L(p, p) sits in the padding row of the tile (tile_load_l)
            double x = FASTDIV ? div_by(v[p >> 2], Ls[p * LDS + p], Ls[p * LDS + NB]) : v[p >> 2] / Ls[p * LDS + p];
            x = __shfl_sync(FULLM, x, (lane & ~3) | (p & 3));
            if (c == (p & 3)) v[p >> 2] = x;
            const double* lp = Ls + p * LDS + c;
#pragma unroll
            for (int k = p >> 2; k < NK; k++)
                if (4 * k + c > p) v[k] = fma(-x, lp[4 * k], v[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < NK; k++) {
        const int q = 4 * k + c;
        if (q < nt && u < nu) {
            if (ROWS) Bs[q * LDS + u] = v[k];
            else Bs[u * LDS + q] = v[k];
        }
    }
}
template <bool ROWS, bool FASTDIV>
__device__ __forceinline__ void tile_solve_rl(double* Bs, const double* Ls, int nt, int nu) {
    if (nt <= 16) tile_solve_rl_n<ROWS, 16, FASTDIV>(Bs, Ls, nt, nu);
    else if (nt <= 32) tile_solve_rl_n<ROWS, 32, FASTDIV>(Bs, Ls, nt, nu);
    else if (nt <= 48) tile_solve_rl_n<ROWS, 48, FASTDIV>(Bs, Ls, nt, nu);
    else tile_solve_rl_n<ROWS, 64, FASTDIV>(Bs, Ls, nt, nu);
}

// B tile (rows x cols) and lower triangle (n x n) into shared memory, zero-filled up to 64 x 64
__device__ __forceinline__ void tile_load_b(double* Bs, const double* B, int ldb, int rows, int cols) {
    for (int x = threadIdx.x; x < NB * NB; x += blockDim.x) {
        const int i = x & (NB - 1), j = x >> 6;
        Bs[j * LDS + i] = (i < rows && j < cols) ? B[i + (size_t)j * ldb] : 0.0;
    }
}
__device__ __forceinline__ void tile_load_l(double* Ls, const double* L, int ldl, int n) {
    for (int x = threadIdx.x; x < NB * NB; x += blockDim.x) {
        const int i = x & (NB - 1), j = x >> 6;
        const double v = (i < n && j <= i) ? L[i + (size_t)j * ldl] : 0.0;
        Ls[j * LDS + i] = v;
        if (i == j) Ls[j * LDS + NB] = (i < n) ? 1.0 / v : 1.0;  // reciprocal of the diagonal in the padding row
    }
}
__device__ __forceinline__ void tile_store_b(const double* Bs, double* B, int ldb, int rows, int cols) {
    for (int x = threadIdx.x; x < NB * NB; x += blockDim.x) {
        const int i = x & (NB - 1), j = x >> 6;
        if (i < rows && j < cols) B[i + (size_t)j * ldb] = Bs[j * LDS + i];
    }
}

template <int MODE, bool FASTDIV>
__global__ void __launch_bounds__(256) trsm_mid256_kernel(DevTables T, const SymTrsm* __restrict__ tasks,
                                                       const int* __restrict__ mid, const int* __restrict__ cnt) {
    extern __shared__ double trsm_smem[];
    double* Bs = trsm_smem;
    double* Ls = trsm_smem + NB * LDS;
    const int nmid = *cnt;
    for (int q = blockIdx.x; q < nmid; q += gridDim.x) {
        const SymTrsm t = tasks[mid[q]];
        const int m = T.csize[t.cm], n = T.csize[t.cn];  // free dimension, triangle
        double* B = T.eptr[t.eB];
        const int ldb = T.eld[t.eB];
        const int rows = MODE == TRSM_RLT ? m : n, cols = MODE == TRSM_RLT ? n : m;
        tile_load_b(Bs, B, ldb, rows, cols);
        tile_load_l(Ls, T.eptr[t.eT], T.eld[t.eT], n);
        __syncthreads();
        tile_solve_rl<MODE == TRSM_RLT, FASTDIV>(Bs, Ls, n, m);
        __syncthreads();
        tile_store_b(Bs, B, ldb, rows, cols);
        __syncthreads();
    }
}

// Two-sided scaling of one off-diagonal block: B (|c2| x |c1|) <- L_c2^-1 (B L_c1^-T)   (src/tree.cpp:796-856)
__global__ void __launch_bounds__(128) scale_sym_kernel(DevTables T, const SymTrsm* __restrict__ right,
                                                        const SymTrsm* __restrict__ left, int nt, int* mid, int* cnt,
                                                        int warp_only) {
    __shared__ double Sw[4][32 * WLD];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ti = blockIdx.x * 4 + w;
    if (ti >= nt) return;
    const SymTrsm r = right[ti];
    if (T.rank >= 0 && T.owner[T.en1[r.eB]] != T.rank) return;
    const int rows = T.csize[r.cm], cols = T.csize[r.cn];
    if (rows <= 0 || cols <= 0 || rows > SMALL_DIM || cols > SMALL_DIM) return;
    if (rows > 32 || cols > 32) {
        if (lane == 0 && !(warp_only & 1)) push_mid(mid, cnt, ti);
        return;
    }
    const int eL = left[ti].eT;
    double* B = T.eptr[r.eB];
    const int ldb = T.eld[r.eB];
    double* S = Sw[w];
    warp_tile_load(S, B, ldb, rows, cols, lane);
    __syncwarp();
    if (warp_only & 2) {  // bit 1: fast division
        warp_solve_right_lt<true>(S, rows, cols, T.eptr[r.eT], T.eld[r.eT], lane);
        __syncwarp();
        warp_solve_left_ln<true>(S, rows, cols, T.eptr[eL], T.eld[eL], lane);
    } else {
        warp_solve_right_lt<false>(S, rows, cols, T.eptr[r.eT], T.eld[r.eT], lane);
        __syncwarp();
        warp_solve_left_ln<false>(S, rows, cols, T.eptr[eL], T.eld[eL], lane);
    }
    __syncwarp();
    warp_tile_store(S, B, ldb, rows, cols, lane);
}

__global__ void __launch_bounds__(NB) scale_mid_kernel(DevTables T, const SymTrsm* __restrict__ right,
                                                       const SymTrsm* __restrict__ left, const int* __restrict__ mid,
                                                       const int* __restrict__ cnt) {
    extern __shared__ double trsm_smem[];
    const int nmid = *cnt;
    for (int q = blockIdx.x; q < nmid; q += gridDim.x) {
        const SymTrsm r = right[mid[q]];
        const int eL = left[mid[q]].eT;
        const int rows = T.csize[r.cm], cols = T.csize[r.cn];
        double* B = T.eptr[r.eB];
        const int ldb = T.eld[r.eB];
        trsm_tile64<TRSM_RLT>(T.eptr[r.eT], T.eld[r.eT], nullptr, B, ldb, rows, cols, trsm_smem);
        trsm_tile64<TRSM_LLN>(T.eptr[eL], T.eld[eL], nullptr, B, ldb, cols, rows, trsm_smem);
    }
}

template <bool FASTDIV>
__global__ void __launch_bounds__(256, 3) scale_mid256_kernel(DevTables T, const SymTrsm* __restrict__ right,
                                                        const SymTrsm* __restrict__ left, const int* __restrict__ mid,
                                                        const int* __restrict__ cnt) {
    extern __shared__ double trsm_smem[];
    double* Bs = trsm_smem;
    double* Ls = trsm_smem + NB * LDS;
    const int nmid = *cnt;
    for (int q = blockIdx.x; q < nmid; q += gridDim.x) {
        const SymTrsm r = right[mid[q]];
        const int eL = left[mid[q]].eT;
        const int rows = T.csize[r.cm], cols = T.csize[r.cn];
        double* B = T.eptr[r.eB];
        const int ldb = T.eld[r.eB];
        tile_load_b(Bs, B, ldb, rows, cols);
        tile_load_l(Ls, T.eptr[r.eT], T.eld[r.eT], cols);
        __syncthreads();
        tile_solve_rl<true, FASTDIV>(Bs, Ls, cols, rows);  // B <- B L_c1^-T
        __syncthreads();
        tile_load_l(Ls, T.eptr[eL], T.eld[eL], rows);
        __syncthreads();
        tile_solve_rl<false, FASTDIV>(Bs, Ls, rows, cols);  // B <- L_c2^-1 B
        __syncthreads();
        tile_store_b(Bs, B, ldb, rows, cols);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// GEN / PLU blocks up to 64 x 64, plan-driven (PluMode in kernels.cuh). The row permutation commutes with the right
// solve (it acts on the other side), so it is applied while the block is loaded: tile row i <- block row perm[i].
// ------------------------------------------------------------------------------------------------
struct PluBlock {
    double* B;
    const double *U, *ud, *L;
    const int* perm;
    int ldb, ldu, ldl, rows, cols;
    bool valid;
};
template <int MODE>
__device__ __forceinline__ PluBlock plu_resolve(const DevTables& T, const SymTrsm* __restrict__ right,
                                                const SymTrsm* __restrict__ left, int ti,
                                                const double* const* __restrict__ ud,
                                                const int* const* __restrict__ pperm) {
    PluBlock b{};
    b.valid = false;
    int eB;
    if (MODE == PLU_LEFT) {
        const SymTrsm l = left[ti];  // B is |cn| x |cm|
        eB = l.eB;
        b.rows = T.csize[l.cn];
        b.cols = T.csize[l.cm];
        b.L = T.eptr[l.eT];
        b.ldl = T.eld[l.eT];
        b.perm = pperm[l.cn];
    } else {
        const SymTrsm r = right[ti];  // B is |cm| x |cn|
        eB = r.eB;
        b.rows = T.csize[r.cm];
        b.cols = T.csize[r.cn];
        b.U = T.eptr[r.eT];
        b.ldu = T.eld[r.eT];
        b.ud = ud[r.cn];
        if (MODE == PLU_BOTH) {
            const SymTrsm l = left[ti];
            b.L = T.eptr[l.eT];
            b.ldl = T.eld[l.eT];
            b.perm = pperm[l.cn];
        }
    }
    if (T.rank >= 0 && T.owner[T.en1[eB]] != T.rank) return b;
    if (b.rows <= 0 || b.cols <= 0 || b.rows > SMALL_DIM || b.cols > SMALL_DIM) return b;
    b.B = T.eptr[eB];
    b.ldb = T.eld[eB];
    b.valid = true;
    return b;
}

template <int MODE>
__global__ void __launch_bounds__(128) plu_sym_kernel(DevTables T, const SymTrsm* __restrict__ right,
                                                      const SymTrsm* __restrict__ left, int nt,
                                                      const double* const* __restrict__ ud,
                                                      const int* const* __restrict__ pperm, int* mid, int* cnt) {
    __shared__ double Sw[4][32 * WLD];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ti = blockIdx.x * 4 + w;
    if (ti >= nt) return;
    const PluBlock b = plu_resolve<MODE>(T, right, left, ti, ud, pperm);
    if (!b.valid) return;
    if (b.rows > 32 || b.cols > 32) {
        if (lane == 0) push_mid(mid, cnt, ti);
        return;
    }
    double* S = Sw[w];
    const int rows = b.rows, cols = b.cols;
    for (int x = lane; x < rows * cols; x += 32) {
        const int i = x % rows, j = x / rows;
        S[j * WLD + i] = b.B[(MODE == PLU_RIGHT ? i : b.perm[i]) + (size_t)j * b.ldb];
    }
    __syncwarp();
    if (MODE != PLU_LEFT) {
        // X U = B : lane = row; U(p, j) above the diagonal of the pivot block, diag(U) beside it
        if (lane < rows)
            for (int j = 0; j < cols; j++) {
                const double d = b.ud[j], rd = 1.0 / d;
                double v = S[j * WLD + lane];
                for (int p = 0; p < j; p++) v -= S[p * WLD + lane] * b.U[p + (size_t)j * b.ldu];
                S[j * WLD + lane] = div_by(v, d, rd);
            }
        __syncwarp();
    }
    if (MODE != PLU_RIGHT) {
        warp_solve_left_ln<true>(S, rows, cols, b.L, b.ldl, lane);
        __syncwarp();
    }
    warp_tile_store(S, b.B, b.ldb, rows, cols, lane);
}

template <int MODE>
__global__ void __launch_bounds__(256, 3) plu_mid256_kernel(DevTables T, const SymTrsm* __restrict__ right,
                                                            const SymTrsm* __restrict__ left,
                                                            const double* const* __restrict__ ud,
                                                            const int* const* __restrict__ pperm,
                                                            const int* __restrict__ mid, const int* __restrict__ cnt) {
    extern __shared__ double trsm_smem[];
    double* Bs = trsm_smem;
    double* Ls = trsm_smem + NB * LDS;
    const int nmid = *cnt;
    for (int q = blockIdx.x; q < nmid; q += gridDim.x) {
        const PluBlock b = plu_resolve<MODE>(T, right, left, mid[q], ud, pperm);
        const int rows = b.rows, cols = b.cols;
        for (int x = threadIdx.x; x < NB * NB; x += blockDim.x) {
            const int i = x & (NB - 1), j = x >> 6;
            Bs[j * LDS + i] =
                (i < rows && j < cols) ? b.B[(MODE == PLU_RIGHT ? i : b.perm[i]) + (size_t)j * b.ldb] : 0.0;
        }
        if (MODE != PLU_LEFT) {
            // U^T as the lower triangle of the right-looking solve: Ls[p * LDS + q] = U(p, q), diagonal from ud
            for (int x = threadIdx.x; x < NB * NB; x += blockDim.x) {
                const int i = x & (NB - 1), j = x >> 6;
                double v = 0.0;
                if (i < cols && j < i) v = b.U[j + (size_t)i * b.ldu];
                else if (i == j && i < cols) v = b.ud[i];
                Ls[j * LDS + i] = v;
                if (i == j) Ls[j * LDS + NB] = (i < cols) ? 1.0 / v : 1.0;
            }
            __syncthreads();
            tile_solve_rl<true, true>(Bs, Ls, cols, rows);  // B <- B U^-1
            __syncthreads();
        }
        if (MODE != PLU_RIGHT) {
            tile_load_l(Ls, b.L, b.ldl, rows);
            __syncthreads();
            tile_solve_rl<false, true>(Bs, Ls, rows, cols);  // B <- L^-1 (P^T B)
            __syncthreads();
        }
        tile_store_b(Bs, b.B, b.ldb, rows, cols);
        __syncthreads();
    }
}

__device__ __forceinline__ GemmContrib resolve_contrib(const DevTables& T, const SymCon c) {
    GemmContrib g;
    g.A = T.eptr[c.e1];
    g.lda = T.eld[c.e1];
    g.B = T.eptr[c.e2];
    g.ldb = T.eld[c.e2];
    g.k = T.csize[T.en1[c.e1]];
    return g;
}

__global__ void __launch_bounds__(128) gemm_sym_small_kernel(DevTables T, const SymGemm* __restrict__ tasks, int nt,
                                                             const SymCon* __restrict__ con, int* mid, int* cnt) {
    const int ti = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (ti >= nt) return;
    const int lane = threadIdx.x & 31;
    const SymGemm t = tasks[ti];
    if (T.rank >= 0 && T.owner[T.en1[t.target]] != T.rank) return;
    const int m = T.csize[T.en2[t.target]], n = T.csize[T.en1[t.target]];
    if (m <= 0 || n <= 0 || m > SMALL_DIM || n > SMALL_DIM) return;
    if (m * n > 1024) {
        if (lane == 0) push_mid(mid, cnt, ti);
        return;
    }
    const bool nn = (t.flags & GEMM_NN) != 0, zero = (t.flags & GEMM_ZERO_INIT) != 0, lower = (t.flags & GEMM_LOWER) != 0;
    double* C = T.eptr[t.target];
    const int ldc = T.eld[t.target];
    const int total = m * n;
    for (int e = lane; e < total; e += 32) {
        const int i = e % m, j = e / m;
        if (lower && j > i) continue;
        double acc = 0.0;
        for (int ci = 0; ci < t.nc; ci++) {
            const GemmContrib c = resolve_contrib(T, con[t.c0 + ci]);
            const double* a = c.A + i;
            if (!nn) {
                const double* bb = c.B + j;
                for (int p = 0; p < c.k; p++) acc += a[(size_t)p * c.lda] * bb[(size_t)p * c.ldb];
            } else {
                const double* bb = c.B + (size_t)j * c.ldb;
                for (int p = 0; p < c.k; p++) acc += a[(size_t)p * c.lda] * bb[p];
            }
        }
        double* p = C + i + (size_t)j * ldc;
        *p = (zero ? 0.0 : *p) - acc;
    }
}

__global__ void __launch_bounds__(256) gemm_sym_mid_kernel(DevTables T, const SymGemm* __restrict__ tasks,
                                                           const SymCon* __restrict__ con, const int* __restrict__ mid,
                                                           const int* __restrict__ cnt) {
    const int nmid = *cnt;
    for (int q = blockIdx.x; q < nmid; q += gridDim.x) {
        const SymGemm t = tasks[mid[q]];
        const int m = T.csize[T.en2[t.target]], n = T.csize[T.en1[t.target]];
        const SymCon* cc = con + t.c0;
        gemm_tile64(T.eptr[t.target], T.eld[t.target], m, n, t.flags, 0, 0, t.nc,
                    [&T, cc](int ci) { return resolve_contrib(T, cc[ci]); });
    }
}

// Merge: child block -> parent block (pre-zeroed). One warp per block of at most COPY_SMALL elements; larger blocks
// are chunked by the host. Identity pivots (after scaling, src/tree.cpp:811) are written as a diagonal of ones.
__global__ void __launch_bounds__(128) copy_sym_kernel(DevTables T, const SymCopy* __restrict__ tasks, int nt,
                                                       int pivots_identity) {
    const int ti = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (ti >= nt) return;
    const int lane = threadIdx.x & 31;
    const SymCopy t = tasks[ti];
    if (T.rank >= 0 && T.owner[t.c1] != T.rank) return;
    const int rows = T.csize[t.c2], cols = T.csize[t.c1];
    if (rows <= 0 || cols <= 0) return;
    const int ldd = T.eld[t.enew];
    double* dst = T.eptr[t.enew] + T.pos[t.c2] + (size_t)T.pos[t.c1] * ldd;
    if (pivots_identity && t.c1 == t.c2) {
        for (int i = lane; i < rows; i += 32) dst[i + (size_t)i * ldd] = 1.0;
        return;
    }
    const int tot = rows * cols;
    if (tot > COPY_SMALL) return;
    const double* src = T.eptr[t.eold];
    const int lds = T.eld[t.eold];
    if (rows >= 32) {
        for (int j = 0; j < cols; j++)
            for (int i = lane; i < rows; i += 32) dst[i + (size_t)j * ldd] = src[i + (size_t)j * lds];
    } else {
        for (int e = lane; e < tot; e += 32) {
            int i = e % rows, j = e / rows;
            dst[i + (size_t)j * ldd] = src[i + (size_t)j * lds];
        }
    }
}

__global__ void expand_qsrc_kernel(DevTables T, const SymQrSrc* __restrict__ s, int n, QrSrc* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    SymQrSrc q = s[i];
    QrSrc o;
    o.blk = T.eptr[q.edge];
    o.ld = T.eld[q.edge];
    o.nbr = q.nbr;
    o.transposed = q.transposed;
    out[i] = o;
}

__global__ void expand_trsv_kernel(DevTables T, const int* __restrict__ clusters, const int* __restrict__ piv, int nt,
                                   TrsvTask* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    const int c = clusters[i], e = piv[i];
    TrsvTask o;
    o.T = T.eptr[e];
    o.x = T.xptr[c];
    o.ld = T.eld[e];
    o.n = (T.rank >= 0 && T.owner[c] != T.rank) ? 0 : T.csize[c];
    o.diag = nullptr;
    o.perm = nullptr;
    out[i] = o;
}

__global__ void expand_gemv_kernel(DevTables T, const SymGemv* __restrict__ t, int nt, const SymGemvCon* __restrict__ c,
                                   int ncon, GemvTask* out, GemvContrib* outc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nt) {
        SymGemv g = t[i];
        GemvTask o;
        o.y = T.xptr[g.cluster];
        o.m = (T.rank >= 0 && T.owner[g.cluster] != T.rank) ? 0 : T.csize[g.cluster];
        o.c0 = g.c0;
        o.nc = g.nc;
        out[i] = o;
    }
    if (i < ncon) {
        SymGemvCon g = c[i];
        GemvContrib o;
        o.A = T.eptr[g.edge];
        o.x = T.xptr[g.xcluster];
        o.lda = T.eld[g.edge];
        o.k = T.csize[g.xcluster];
        outc[i] = o;
    }
}

__global__ void expand_xcopy_kernel(DevTables T, const int* __restrict__ children, int n, XCopyTask* fwd, XCopyTask* bwd) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = children[i];
    double* xc = T.xptr[c];
    double* xp = T.xptr[T.parent[c]] + T.pos[c];
    const int sz = (T.rank >= 0 && T.owner[c] != T.rank) ? 0 : T.csize[c];
    fwd[i] = XCopyTask{xc, xp, sz};
    bwd[i] = XCopyTask{xp, xc, sz};
}

__global__ void expand_house_kernel(DevTables T, const QrTask* __restrict__ q, int n, HouseTask* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    QrTask t = q[i];
    HouseTask o;
    o.V = t.V;
    o.tau = t.tau;
    o.x = T.xptr[t.cluster];
    o.rows = t.rows;
    o.rank = min(t.rows, T.csize[t.cluster]);
    out[i] = o;
}

__global__ void scatter_values_kernel(const double* __restrict__ val, const unsigned* __restrict__ map, size_t n,
                                      double* dst) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        unsigned m = map[i];
        if (m != 0xffffffffu) dst[m] = val[i];
    }
}


__global__ void peer_barrier_kernel(PeerPtrs flags, int rank, int nranks, unsigned epoch) {
    const int r = threadIdx.x;
    if (r < nranks) {
        __threadfence_system();  // everything this GPU wrote before is visible to the peers first
        volatile unsigned* theirs = reinterpret_cast<volatile unsigned*>(flags.p[r]) + rank;
        *theirs = epoch;
        volatile unsigned* mine = reinterpret_cast<volatile unsigned*>(flags.p[rank]) + r;
        // a peer that died (or never arrives) must not hang this GPU: give up after 60 s and fault the context,
        // which the host sees as a CUDA error on the next synchronisation
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        unsigned spins = 0;
        while ((int)(*mine - epoch) < 0) {
            if ((++spins & 0xfffu) == 0) {
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 60000000000ull) __trap();
            }
        }
        __threadfence_system();
    }
}

__global__ void csize_min_kernel(PeerPtrs csize, int rank, int nranks, int first, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int* mine = reinterpret_cast<int*>(csize.p[rank]) + first;
    int v = mine[i];
    for (int r = 0; r < nranks; r++)
        if (r != rank) v = min(v, reinterpret_cast<const volatile int*>(csize.p[r])[first + i]);
    mine[i] = v;
}

__global__ void scatter_owned_kernel(int n, const int* __restrict__ idx, PeerPtrs leaf,
                                     const signed char* __restrict__ dof_owner, double* dst) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        dst[idx[i]] = reinterpret_cast<const double*>(leaf.p[dof_owner[i]])[i];
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

void launch_trtri(const TrtriTask* t, int nt, int max_n, cudaStream_t st) {
    if (nt <= 0 || max_n <= 0) return;
    dim3 grid(nt, (max_n + NB - 1) / NB);
    trtri_kernel<<<grid, NB, 0, st>>>(t);
}
void launch_ut(const UtTask* t, int nt, cudaStream_t st) {
    if (nt > 0) ut_kernel<<<dim3(nt, 8), 256, 0, st>>>(t);
}
void launch_eye(const EyeTask* t, int nt, cudaStream_t st) {
    if (nt > 0) eye_kernel<<<dim3(nt, 8), 256, 0, st>>>(t);
}
void launch_trsm_strip(int mode, const TrsmTask* t, int nt, const int* strip_prefix, int total_strips, cudaStream_t st) {
    if (nt <= 0 || total_strips <= 0) return;
    if (mode == TRSM_RLT) trsm_strip_kernel<TRSM_RLT><<<total_strips, 256, 0, st>>>(t, nt, strip_prefix);
    else trsm_strip_kernel<TRSM_LLN><<<total_strips, 256, 0, st>>>(t, nt, strip_prefix);
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
    if (nt <= 0) return;
    constexpr int smem = 6144 * sizeof(double) + 64;  // above the 48 KB default: needs the opt-in attribute
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(rowperm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
    }
    rowperm_kernel<<<dim3(nt, nt >= 2048 ? 1 : (nt >= 256 ? 4 : 16)), 128, smem, st>>>(t);
}

void launch_gemm_tiled(const GemmTask* t, int nt, const GemmContrib* c, const int* tile_prefix, int total_tiles,
                       cudaStream_t st, const int* tile_task) {
    if (nt > 0 && total_tiles > 0) gemm_tiled_kernel<<<total_tiles, 256, 0, st>>>(t, nt, c, tile_prefix, tile_task);
}

void launch_gemm_small(const GemmTask* t, int nt, const GemmContrib* c, cudaStream_t st) {
    if (nt > 0) gemm_small_kernel<<<(nt + 3) / 4, 128, 0, st>>>(t, nt, c);
}

void launch_copy(const CopyTask* t, int nt, cudaStream_t st) {
    if (nt > 0) copy_kernel<<<(nt + 3) / 4, 128, 0, st>>>(t, nt);
}

void launch_trsv(const TrsvTask* t, int nt, int trans, int max_n, cudaStream_t st) {
    if (nt <= 0) return;
    // triangles of at most 32 rows: one warp each; the CTA kernel takes the rest (and is skipped when there is none)
    trsv_small_kernel<<<(nt + 3) / 4, 128, 0, st>>>(t, nt, trans);
    if (max_n > 32) trsv_kernel<<<nt, SV_T, 0, st>>>(t, trans, 1);
}
void launch_gemv(const GemvTask* t, int nt, const GemvContrib* c, int trans, int max_m, cudaStream_t st) {
    if (nt <= 0) return;
    if (max_m <= 64) {
        gemv_kernel<<<nt, SV_T, 0, st>>>(t, c, trans);
    } else if (trans == 0) {
        gemv_n_big_kernel<<<dim3(nt, (max_m + 63) / 64), GV_T, 0, st>>>(t, c);
    } else {
        gemv_t_big_kernel<<<dim3(nt, (max_m + 31) / 32), GV_T, 0, st>>>(t, c);
    }
}
void launch_house(const HouseTask* t, int nt, int trans, int max_rows, cudaStream_t st) {
    if (nt <= 0) return;
    if (max_rows <= 64) house_reg_kernel<32, 2, true><<<(nt + 3) / 4, 128, 0, st>>>(t, nt, trans);
    else if (max_rows <= 1024) house_reg_kernel<256, 4, false><<<nt, 256, 0, st>>>(t, nt, trans);
    else house_kernel<<<nt, SV_T, 0, st>>>(t, trans);
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

// ---- plan-driven launchers ----
namespace {
constexpr int kMidGrid = 148 * 8;
constexpr int kTrsmSmem = 2 * NB * LDS * sizeof(double);
// SPAND_MID256 (default 1): right-looking 256-thread kernels for the 33..64 class; 0: one thread per row / column
bool mid256() {
    static int v = -1;
    if (v < 0) {
        const char* e = std::getenv("SPAND_MID256");
        v = e ? std::atoi(e) : 1;
    }
    return v != 0;
}
// SPAND_FASTDIV (default 1): divisions of the substitutions (blocks up to 64) as product + exact remainder + correction
// with the reciprocal of the diagonal entry taken off the dependent chain (div_by)
bool fastdiv() {
    static int v = -1;
    if (v < 0) {
        const char* e = std::getenv("SPAND_FASTDIV");
        v = e ? std::atoi(e) : 1;
    }
    return v != 0;
}
void configure_mid_smem() {
    static bool configured = false;
    if (configured) return;
    cudaFuncSetAttribute(trsm_mid_kernel<TRSM_RLT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTrsmSmem);
    cudaFuncSetAttribute(trsm_mid_kernel<TRSM_LLN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTrsmSmem);
    cudaFuncSetAttribute(scale_mid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTrsmSmem);
    cudaFuncSetAttribute(trsm_mid256_kernel<TRSM_RLT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTrsmSmem);
    cudaFuncSetAttribute(trsm_mid256_kernel<TRSM_LLN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTrsmSmem);
    cudaFuncSetAttribute(scale_mid256_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTrsmSmem);
    cudaFuncSetAttribute(trsm_mid256_kernel<TRSM_RLT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTrsmSmem);
    cudaFuncSetAttribute(trsm_mid256_kernel<TRSM_LLN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTrsmSmem);
    cudaFuncSetAttribute(scale_mid256_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTrsmSmem);
    configured = true;
}
}  // namespace

void launch_potrf_sym(const DevTables& T, const int* clusters, const int* piv, int nt, int* mid, int* cnt, int* err,
                      cudaStream_t st) {
    if (nt <= 0) return;
    potrf_sym_kernel<<<(nt + 3) / 4, 128, 0, st>>>(T, clusters, piv, nt, mid, cnt, err);
    potrf_mid_kernel<<<std::min(nt, kMidGrid), NB, 0, st>>>(T, clusters, piv, mid, cnt, err);
}
void launch_trsm_sym(int mode, const DevTables& T, const SymTrsm* tasks, int nt, int* mid, int* cnt, cudaStream_t st) {
    if (nt <= 0) return;
    configure_mid_smem();
    if (mode == TRSM_RLT) {
        trsm_sym_kernel<TRSM_RLT><<<(nt + 3) / 4, 128, 0, st>>>(T, tasks, nt, mid, cnt, fastdiv() ? 1 : 0);
        if (mid256() && fastdiv())
            trsm_mid256_kernel<TRSM_RLT, true><<<std::min(nt, kMidGrid), 256, kTrsmSmem, st>>>(T, tasks, mid, cnt);
        else if (mid256())
            trsm_mid256_kernel<TRSM_RLT, false><<<std::min(nt, kMidGrid), 256, kTrsmSmem, st>>>(T, tasks, mid, cnt);
        else trsm_mid_kernel<TRSM_RLT><<<std::min(nt, kMidGrid), NB, kTrsmSmem, st>>>(T, tasks, mid, cnt);
    } else {
        trsm_sym_kernel<TRSM_LLN><<<(nt + 3) / 4, 128, 0, st>>>(T, tasks, nt, mid, cnt, fastdiv() ? 1 : 0);
        if (mid256() && fastdiv())
            trsm_mid256_kernel<TRSM_LLN, true><<<std::min(nt, kMidGrid), 256, kTrsmSmem, st>>>(T, tasks, mid, cnt);
        else if (mid256())
            trsm_mid256_kernel<TRSM_LLN, false><<<std::min(nt, kMidGrid), 256, kTrsmSmem, st>>>(T, tasks, mid, cnt);
        else trsm_mid_kernel<TRSM_LLN><<<std::min(nt, kMidGrid), NB, kTrsmSmem, st>>>(T, tasks, mid, cnt);
    }
}
void launch_scale_sym(const DevTables& T, const SymTrsm* right, const SymTrsm* left, int nt, int* mid, int* cnt,
                      cudaStream_t st, bool warp_only) {
    if (nt <= 0) return;
    configure_mid_smem();
    scale_sym_kernel<<<(nt + 3) / 4, 128, 0, st>>>(T, right, left, nt, mid, cnt, (warp_only ? 1 : 0) | (fastdiv() ? 2 : 0));
    if (warp_only) return;
    if (mid256() && fastdiv())
        scale_mid256_kernel<true><<<std::min(nt, kMidGrid), 256, kTrsmSmem, st>>>(T, right, left, mid, cnt);
    else if (mid256())
        scale_mid256_kernel<false><<<std::min(nt, kMidGrid), 256, kTrsmSmem, st>>>(T, right, left, mid, cnt);
    else scale_mid_kernel<<<std::min(nt, kMidGrid), NB, kTrsmSmem, st>>>(T, right, left, mid, cnt);
}
namespace {
template <int MODE>
void launch_plu_sym_mode(const DevTables& T, const SymTrsm* right, const SymTrsm* left, int nt, const double* const* ud,
                         const int* const* pperm, int* mid, int* cnt, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(plu_mid256_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTrsmSmem);
        configured = true;
    }
    plu_sym_kernel<MODE><<<(nt + 3) / 4, 128, 0, st>>>(T, right, left, nt, ud, pperm, mid, cnt);
    plu_mid256_kernel<MODE><<<std::min(nt, kMidGrid), 256, kTrsmSmem, st>>>(T, right, left, ud, pperm, mid, cnt);
}
}  // namespace
void launch_plu_sym(int mode, const DevTables& T, const SymTrsm* right, const SymTrsm* left, int nt,
                    const double* const* ud, const int* const* pperm, int* mid, int* cnt, cudaStream_t st) {
    if (nt <= 0) return;
    if (mode == PLU_BOTH) launch_plu_sym_mode<PLU_BOTH>(T, right, left, nt, ud, pperm, mid, cnt, st);
    else if (mode == PLU_RIGHT) launch_plu_sym_mode<PLU_RIGHT>(T, right, left, nt, ud, pperm, mid, cnt, st);
    else launch_plu_sym_mode<PLU_LEFT>(T, right, left, nt, ud, pperm, mid, cnt, st);
}
void launch_gemm_sym(const DevTables& T, const SymGemm* tasks, int nt, const SymCon* con, int* mid, int* cnt,
                     cudaStream_t st) {
    if (nt <= 0) return;
    gemm_sym_small_kernel<<<(nt + 3) / 4, 128, 0, st>>>(T, tasks, nt, con, mid, cnt);
    gemm_sym_mid_kernel<<<std::min(nt, kMidGrid), 256, 0, st>>>(T, tasks, con, mid, cnt);
}
void launch_copy_sym(const DevTables& T, const SymCopy* tasks, int nt, int pivots_identity, cudaStream_t st) {
    if (nt > 0) copy_sym_kernel<<<(nt + 3) / 4, 128, 0, st>>>(T, tasks, nt, pivots_identity);
}
void launch_expand_qsrc(const DevTables& T, const SymQrSrc* s, int n, QrSrc* out, cudaStream_t st) {
    if (n > 0) expand_qsrc_kernel<<<(n + 255) / 256, 256, 0, st>>>(T, s, n, out);
}
void launch_expand_trsv(const DevTables& T, const int* clusters, const int* piv, int nt, TrsvTask* out, cudaStream_t st) {
    if (nt > 0) expand_trsv_kernel<<<(nt + 255) / 256, 256, 0, st>>>(T, clusters, piv, nt, out);
}
void launch_expand_gemv(const DevTables& T, const SymGemv* t, int nt, const SymGemvCon* c, int ncon, GemvTask* out,
                        GemvContrib* outc, cudaStream_t st) {
    int n = std::max(nt, ncon);
    if (n > 0) expand_gemv_kernel<<<(n + 255) / 256, 256, 0, st>>>(T, t, nt, c, ncon, out, outc);
}
void launch_expand_xcopy(const DevTables& T, const int* children, int n, XCopyTask* fwd, XCopyTask* bwd, cudaStream_t st) {
    if (n > 0) expand_xcopy_kernel<<<(n + 255) / 256, 256, 0, st>>>(T, children, n, fwd, bwd);
}
void launch_expand_house(const DevTables& T, const QrTask* q, int n, HouseTask* out, cudaStream_t st) {
    if (n > 0) expand_house_kernel<<<(n + 255) / 256, 256, 0, st>>>(T, q, n, out);
}
void launch_scatter_values(const double* val, const unsigned* map, size_t n, double* dst, cudaStream_t st) {
    if (n > 0) scatter_values_kernel<<<grid_for(n, 256), 256, 0, st>>>(val, map, n, dst);
}

__global__ void err_publish_kernel(const int* err, int* shared_word) { *shared_word = *err; }
__global__ void err_or_kernel(PeerPtrs words, int nranks, int* err) {
    int e = 0;
    for (int r = 0; r < nranks; r++) e |= *reinterpret_cast<volatile int*>(words.p[r]);
    *err |= e;
}
void launch_err_publish(const int* err, int* shared_word, cudaStream_t st) {
    err_publish_kernel<<<1, 1, 0, st>>>(err, shared_word);
}
void launch_err_or(const PeerPtrs& words, int nranks, int* err, cudaStream_t st) {
    err_or_kernel<<<1, 1, 0, st>>>(words, nranks, err);
}

void launch_peer_barrier(const PeerPtrs& flags, int rank, int nranks, unsigned epoch, cudaStream_t st) {
    peer_barrier_kernel<<<1, 32, 0, st>>>(flags, rank, nranks, epoch);
}
void launch_csize_min(const PeerPtrs& csize, int rank, int nranks, int first, int n, cudaStream_t st) {
    if (n > 0) csize_min_kernel<<<(n + 255) / 256, 256, 0, st>>>(csize, rank, nranks, first, n);
}
void launch_scatter_owned(int n, const int* idx, const PeerPtrs& leaf, const signed char* dof_owner, double* dst,
                          cudaStream_t st) {
    if (n > 0) scatter_owned_kernel<<<grid_for(n, 256), 256, 0, st>>>(n, idx, leaf, dof_owner, dst);
}

}  // namespace spand
