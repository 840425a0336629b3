// Device kernels of the Householder GMRES driver (replaces the Eigen vector algebra of /root/reference
// src/is.cpp:123-300: makeHouseholder, applyHouseholderOnTheLeft, Unit(), x += x_new). All of it is HBM-bound
// level-1 work on vectors of N doubles: every reflection reads its Householder vector once per pass.
//
// A reflection needs one global reduction (essential^T tail) before the update, so it runs as two launches on a
// fixed grid of KR_GRID CTAs: the first writes one partial sum per CTA (+ the head element, which the second
// launch overwrites), the second lets every CTA add the partials in the same fixed order — deterministic, no
// atomics, no host round trip — and applies the update to its slice. tau lives in device memory.
#include <cfloat>

#include "kernels.cuh"

namespace spand {

namespace {

constexpr int KR_T = 256;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ double kr_block_sum(double v, double* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < KR_T / 32; w++) s += red[w];
    return s;
}

// every thread returns sum(partial[0 .. KR_GRID)) added in a fixed order
__device__ __forceinline__ double kr_total(const double* __restrict__ partial, double* red) {
    double s = 0.0;
    for (int i = threadIdx.x; i < KR_GRID; i += KR_T) s += partial[i];
    return kr_block_sum(s, red);
}

// partial[b] = sum over the CTA's slice of a[i] * b[i], i in [0, n); partial[KR_GRID] = head[0]
__global__ void __launch_bounds__(KR_T) kr_dot_partial_kernel(int n, const double* __restrict__ a,
                                                             const double* __restrict__ b, const double* head,
                                                             double* __restrict__ partial) {
    __shared__ double red[KR_T / 32];
    double s = 0.0;
    for (int i = blockIdx.x * KR_T + threadIdx.x; i < n; i += KR_GRID * KR_T) s = fma(a[i], b[i], s);
    s = kr_block_sum(s, red);
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = s;
        if (blockIdx.x == 0) partial[KR_GRID] = head[0];
    }
}

// applyHouseholderOnTheLeft on v (length n) with essential part ess (length n - 1) and *tau:
// tmp = ess^T v[1:] + v[0];  v[0] -= tau tmp;  v[1:] -= tau tmp ess     (n == 1: v[0] *= 1 - tau)
__global__ void __launch_bounds__(KR_T) kr_house_apply_kernel(int n, double* __restrict__ v,
                                                             const double* __restrict__ ess,
                                                             const double* __restrict__ tau_p,
                                                             const double* __restrict__ partial) {
    __shared__ double red[KR_T / 32];
    const double tau = *tau_p;
    if (n == 1) {
        if (blockIdx.x == 0 && threadIdx.x == 0) v[0] = partial[KR_GRID] * (1.0 - tau);
        return;
    }
    if (tau == 0.0) return;
    const double tmp = kr_total(partial, red) + partial[KR_GRID];
    const double tt = tau * tmp;
    if (blockIdx.x == 0 && threadIdx.x == 0) v[0] = partial[KR_GRID] - tt;
    for (int i = blockIdx.x * KR_T + threadIdx.x; i < n - 1; i += KR_GRID * KR_T) v[i + 1] -= tt * ess[i];
}

// makeHouseholder of x (length n): partial holds the partial sums of |x[1:]|^2 and x[0].
// ess = x[1:] / (c0 - beta), *tau_out, *beta_out; tail norm (squared) <= DBL_MIN: tau = 0, beta = c0, ess = 0.
__global__ void __launch_bounds__(KR_T) kr_house_make_kernel(int n, const double* __restrict__ x,
                                                            double* __restrict__ ess, double* tau_out,
                                                            double* beta_out, const double* __restrict__ partial) {
    __shared__ double red[KR_T / 32];
    const double tail2 = (n > 1) ? kr_total(partial, red) : 0.0;
    const double c0 = partial[KR_GRID];
    double beta, tau, scal;
    if (n == 1 || tail2 <= DBL_MIN) {
        tau = 0.0;
        beta = c0;
        scal = 0.0;
    } else {
        beta = sqrt(c0 * c0 + tail2);
        if (c0 >= 0.0) beta = -beta;
        scal = c0 - beta;
        tau = (beta - c0) / beta;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *tau_out = tau;
        *beta_out = beta;
    }
    for (int i = blockIdx.x * KR_T + threadIdx.x; i < n - 1; i += KR_GRID * KR_T)
        ess[i] = (scal == 0.0) ? 0.0 : x[i + 1] / scal;
}

// out[0] = sum(partial[0 .. KR_GRID)), one CTA, fixed order
__global__ void __launch_bounds__(KR_T) kr_dot_final_kernel(const double* __restrict__ partial, double* out) {
    __shared__ double red[KR_T / 32];
    const double s = kr_total(partial, red);
    if (threadIdx.x == 0) out[0] = s;
}

// v = e_k
__global__ void kr_unit_kernel(int n, int k, double* __restrict__ v) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] = (i == k) ? 1.0 : 0.0;
}

// y = b - y
__global__ void kr_residual_kernel(int n, const double* __restrict__ b, double* __restrict__ y) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] = b[i] - y[i];
}

}  // namespace

void launch_kr_house_apply(int n, double* v, const double* ess, const double* tau, double* partial, cudaStream_t st) {
    if (n <= 0) return;
    kr_dot_partial_kernel<<<KR_GRID, KR_T, 0, st>>>(n - 1, ess, v + 1, v, partial);
    kr_house_apply_kernel<<<KR_GRID, KR_T, 0, st>>>(n, v, ess, tau, partial);
}
void launch_kr_house_make(int n, const double* x, double* ess, double* tau, double* beta, double* partial,
                          cudaStream_t st) {
    if (n <= 0) return;
    kr_dot_partial_kernel<<<KR_GRID, KR_T, 0, st>>>(n - 1, x + 1, x + 1, x, partial);
    kr_house_make_kernel<<<KR_GRID, KR_T, 0, st>>>(n, x, ess, tau, beta, partial);
}
void launch_dot_det(int n, const double* a, const double* b, double* partial, double* out, cudaStream_t st) {
    if (n <= 0) {
        cudaMemsetAsync(out, 0, sizeof(double), st);
        return;
    }
    kr_dot_partial_kernel<<<KR_GRID, KR_T, 0, st>>>(n, a, b, a, partial);
    kr_dot_final_kernel<<<1, KR_T, 0, st>>>(partial, out);
}
void launch_kr_unit(int n, int k, double* v, cudaStream_t st) {
    if (n > 0) kr_unit_kernel<<<KR_GRID, KR_T, 0, st>>>(n, k, v);
}
void launch_kr_residual(int n, const double* b, double* y, cudaStream_t st) {
    if (n > 0) kr_residual_kernel<<<KR_GRID, KR_T, 0, st>>>(n, b, y);
}

}  // namespace spand
