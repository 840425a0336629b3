// Batched truncated column-pivoted QR (the sparsification kernel) for sm_100a.
//
// Replaces, for every cluster of one wavefront at once, the reference sequence (citations relative to
// /root/reference): assemble_Asn src/tree.cpp:1189-1224 -> LAPACKE_dgeqp3 src/util.cpp:383-394 ->
// choose_rank src/util.cpp:434-452 -> triu(R[:rank,:]) P^T src/tree.cpp:1334-1335 -> scatter
// src/tree.cpp:1004-1046, plus the (V, tau) of the Orthogonal op src/tree.cpp:1322-1331.
//
// Mapping to the machine: one *thread-block cluster* of G CTAs (G = 1,2,4,8,16) per matrix. The gathered
// panel W = [A_s,n ... (A_n,s)^T ...] (rows x cols, cols >> rows) is distributed by column slabs over the
// CTAs and stays resident in shared memory (up to ~3.4 MB per cluster over distributed shared memory); only
// when a slab does not fit is it kept in the L2-resident scratch. Per Householder step the CTAs exchange
// one pivot candidate each and the pivot owner broadcasts the reflector through DSMEM; two cluster barriers
// per step. Columns are never physically swapped: LAPACK's swap sequence is tracked as virtual positions so
// that idamax's first-index tie-breaking (dlaqp2) is reproduced exactly. The factorization stops at the
// first pivot with |R_kk| / |R_00| < tol, which is exactly geqp3 followed by choose_rank.
#include <cooperative_groups.h>

#include <cfloat>
#include <climits>

#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace spand {

namespace {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int NT>
__device__ __forceinline__ double block_sum_d(double v, double* red) {
    v = warp_sum_d(v);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < NT / 32; w++) s += red[w];
    return s;
}

struct Cand {
    double val;
    int pos, col;
};

__device__ __forceinline__ bool better(double v, int p, double bv, int bp) { return v > bv || (v == bv && p < bp); }

template <int G, int NT>
__global__ void __launch_bounds__(NT) rrqr_cluster_kernel(const QrTask* __restrict__ tasks,
                                                          const QrSrc* __restrict__ srcs, int* csize, double tol,
                                                          int smem_bytes) {
    constexpr int NW = NT / 32;
    const int task_id = blockIdx.x / G;
    const int crank = blockIdx.x % G;
    QrTask t = tasks[task_id];
    const QrSrc* src = srcs + t.src0;
    const int rows = t.rows;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    __shared__ double red[NW];
    __shared__ Cand redc[NW];
    __shared__ Cand cand[16];
    __shared__ double ctrl[4];
    extern __shared__ double dsm[];

    int cols = 0;
    for (int s = 0; s < t.nsrc; s++) cols += csize[src[s].nbr];
    // uniform early exits (no cluster barrier has been used yet)
    if (rows == 0) return;
    if (tol >= 1.0 || cols == 0) {  // choose_rank: tol >= 1 -> 0 ; no neighbours -> 0 columns -> rank 0
        if (crank == 0 && tid == 0) csize[t.cluster] = 0;
        return;
    }
    const int mn = min(rows, cols);
    const int cpc = (cols + G - 1) / G;
    const int c_lo = min(cols, crank * cpc), c_hi = min(cols, c_lo + cpc);
    const bool in_smem = ((size_t)rows * cpc + rows) * sizeof(double) <= (size_t)smem_bytes;
    double* Wl = in_smem ? dsm + rows : t.W + (size_t)c_lo * rows;  // local slab, column c at Wl[(c - c_lo) * rows]
    double* vbuf = dsm;                                             // rows doubles, always in shared memory
    double* vn1 = t.W + (size_t)rows * t.maxcols;
    double* vn2 = vn1 + t.maxcols;
    int* pos = t.ipiv;                 // current (virtual) position of original column c
    int* colAt = t.ipiv + t.maxcols;   // original column at position p  (== LAPACK's jpvt)
    cg::cluster_group cluster = cg::this_cluster();
    auto csync = [&]() {
        if (G > 1) cluster.sync();
        else __syncthreads();
    };

    // ---- gather the own column range [c_lo, c_hi) ----
    {
        int c0 = 0;
        for (int s = 0; s < t.nsrc; s++) {
            QrSrc q = src[s];
            int w = csize[q.nbr];
            int a = max(c0, c_lo), b = min(c0 + w, c_hi);
            if (a < b) {
                int nc = b - a, off = a - c0;
                int tot = rows * nc;
                double* dst = Wl + (size_t)(a - c_lo) * rows;
                if (!q.transposed) {
                    const double* sp = q.blk + (size_t)off * q.ld;
                    for (int e = tid; e < tot; e += NT) {
                        int i = e % rows, j = e / rows;
                        dst[i + (size_t)j * rows] = sp[i + (size_t)j * q.ld];
                    }
                } else {
                    const double* sp = q.blk + off;  // block is w x rows
                    for (int e = tid; e < tot; e += NT) {
                        int j = e % nc, i = e / nc;
                        dst[i + (size_t)j * rows] = sp[j + (size_t)i * q.ld];
                    }
                }
            }
            c0 += w;
        }
    }
    __syncthreads();
    const bool thread_cols = rows <= 16;
    // ---- initial column norms, identity positions ----
    if (thread_cols) {
        for (int c = c_lo + tid; c < c_hi; c += NT) {
            const double* cj = Wl + (size_t)(c - c_lo) * rows;
            double s = 0.0;
            for (int i = 0; i < rows; i++) s += cj[i] * cj[i];
            s = sqrt(s);
            vn1[c] = s;
            vn2[c] = s;
            pos[c] = c;
            colAt[c] = c;
        }
    } else {
        for (int c = c_lo + warp; c < c_hi; c += NW) {
            const double* cj = Wl + (size_t)(c - c_lo) * rows;
            double s = 0.0;
            for (int i = lane; i < rows; i += 32) s += cj[i] * cj[i];
            s = sqrt(warp_sum_d(s));
            if (lane == 0) {
                vn1[c] = s;
                vn2[c] = s;
                pos[c] = c;
                colAt[c] = c;
            }
        }
    }
    __threadfence();
    csync();

    const double tol3z = sqrt(DBL_EPSILON);
    int rank = mn;
    double r00 = 0.0;
    for (int k = 0; k < mn; k++) {
        // ---- 1. local pivot candidate: max partial norm, ties -> smallest current position (idamax) ----
        Cand best{-1.0, INT_MAX, -1};
        for (int c = c_lo + tid; c < c_hi; c += NT) {
            int p = pos[c];
            if (p >= k) {
                double v = vn1[c];
                if (better(v, p, best.val, best.pos)) best = Cand{v, p, c};
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, best.val, o);
            int op = __shfl_xor_sync(0xffffffffu, best.pos, o);
            int oc = __shfl_xor_sync(0xffffffffu, best.col, o);
            if (better(ov, op, best.val, best.pos)) best = Cand{ov, op, oc};
        }
        if (lane == 0) redc[warp] = best;
        __syncthreads();
        if (tid == 0) {
            Cand b = redc[0];
            for (int w = 1; w < NW; w++)
                if (better(redc[w].val, redc[w].pos, b.val, b.pos)) b = redc[w];
            if (G > 1) {
                for (int r = 0; r < G; r++) {
                    Cand* rc = cluster.map_shared_rank(cand, r);
                    rc[crank] = b;
                }
            } else {
                cand[0] = b;
            }
        }
        csync();
        // ---- 2. global pivot (every CTA reduces the same G candidates) ----
        Cand piv = cand[0];
#pragma unroll
        for (int r = 1; r < G; r++)
            if (better(cand[r].val, cand[r].pos, piv.val, piv.pos)) piv = cand[r];
        const int pcol = piv.col, ppos = piv.pos;
        const int owner = pcol / cpc;
        if (crank == 0 && tid == 0 && ppos != k) {  // virtual swap of positions k and ppos
            int ck = colAt[k];
            colAt[ppos] = ck;
            pos[ck] = ppos;
            colAt[k] = pcol;
            pos[pcol] = k;
            __threadfence();
        }
        // ---- 3. owner: Householder reflector (dlarfg), broadcast v / beta / tau / stop ----
        if (crank == owner) {
            double* col = Wl + (size_t)(pcol - c_lo) * rows;
            double ss = 0.0;
            for (int i = k + 1 + tid; i < rows; i += NT) ss += col[i] * col[i];
            ss = block_sum_d<NT>(ss, red);
            double alpha = col[k];
            double xnorm = sqrt(ss);
            double beta, tau, scal;
            if (xnorm == 0.0) {
                beta = alpha;
                tau = 0.0;
                scal = 0.0;
            } else {
                beta = -copysign(hypot(alpha, xnorm), alpha);
                tau = (beta - alpha) / beta;
                scal = 1.0 / (alpha - beta);
            }
            bool stop = (k > 0 && tol != 0.0 && !(fabs(beta) / r00 >= tol));
            __syncthreads();
            if (!stop) {
                for (int i = k + 1 + tid; i < rows; i += NT) col[i] *= scal;
                if (tid == 0) {
                    col[k] = beta;
                    t.tau[k] = tau;
                }
            }
            __syncthreads();
            for (int r = 0; r < G; r++) {
                double* rv = (G > 1) ? cluster.map_shared_rank(vbuf, r) : vbuf;
                if (!stop)
                    for (int i = k + 1 + tid; i < rows; i += NT) rv[i] = col[i];
                if (tid == 0) {
                    double* rc = (G > 1) ? cluster.map_shared_rank(ctrl, r) : ctrl;
                    rc[0] = beta;
                    rc[1] = tau;
                    rc[2] = stop ? 1.0 : 0.0;
                }
            }
        }
        csync();
        const double beta = ctrl[0], tau = ctrl[1];
        const bool stop = ctrl[2] != 0.0;
        if (k == 0) r00 = fabs(beta);
        if (stop) {
            rank = k;
            break;
        }
        // ---- 4. apply H to the own active columns, downdate their partial norms (dlaqp2) ----
        const double* v = vbuf;
        if (thread_cols) {
            for (int c = c_lo + tid; c < c_hi; c += NT) {
                if (pos[c] <= k) continue;
                double* cj = Wl + (size_t)(c - c_lo) * rows;
                double w = cj[k];
                for (int i = k + 1; i < rows; i++) w += v[i] * cj[i];
                w *= tau;
                cj[k] -= w;
                for (int i = k + 1; i < rows; i++) cj[i] -= w * v[i];
                double n1 = vn1[c];
                if (n1 != 0.0) {
                    double tmp = fabs(cj[k]) / n1;
                    tmp = fmax(0.0, 1.0 - tmp * tmp);
                    double r = n1 / vn2[c];
                    double tmp2 = tmp * r * r;
                    if (tmp2 <= tol3z) {
                        double s = 0.0;
                        for (int i = k + 1; i < rows; i++) s += cj[i] * cj[i];
                        s = sqrt(s);
                        vn1[c] = s;
                        vn2[c] = s;
                    } else {
                        vn1[c] = n1 * sqrt(tmp);
                    }
                }
            }
        } else {
            for (int c = c_lo + warp; c < c_hi; c += NW) {
                if (pos[c] <= k) continue;
                double* cj = Wl + (size_t)(c - c_lo) * rows;
                double w = 0.0;
                for (int i = k + 1 + lane; i < rows; i += 32) w += v[i] * cj[i];
                w = warp_sum_d(w);
                double ckj = cj[k];
                w = (w + ckj) * tau;
                for (int i = k + 1 + lane; i < rows; i += 32) cj[i] -= w * v[i];
                double newk = ckj - w;
                double n1 = vn1[c];
                bool recompute = false;
                double newn = 0.0;
                if (n1 != 0.0) {
                    double tmp = fabs(newk) / n1;
                    tmp = fmax(0.0, 1.0 - tmp * tmp);
                    double r = n1 / vn2[c];
                    double tmp2 = tmp * r * r;
                    if (tmp2 <= tol3z) recompute = true;
                    else newn = n1 * sqrt(tmp);
                }
                if (recompute) {
                    __syncwarp();
                    double s = 0.0;
                    for (int i = k + 1 + lane; i < rows; i += 32) s += cj[i] * cj[i];
                    s = sqrt(warp_sum_d(s));
                    if (lane == 0) {
                        vn1[c] = s;
                        vn2[c] = s;
                    }
                } else if (lane == 0 && n1 != 0.0) {
                    vn1[c] = newn;
                }
                if (lane == 0) cj[k] = newk;
            }
        }
        __syncthreads();
    }
    if (rank >= rows) return;  // nothing to do (tree.cpp:1317-1319); csize unchanged

    __syncthreads();
    // ---- V: the own pivot columns, in pivot order ----
    for (int c = c_lo + warp; c < c_hi; c += NW) {
        int p = pos[c];
        if (p < rank) {
            const double* cj = Wl + (size_t)(c - c_lo) * rows;
            double* vd = t.V + (size_t)p * rows;
            for (int i = lane; i < rows; i += 32) vd[i] = cj[i];
        }
    }
    // ---- scatter triu(R[:rank,:]) P^T back into the own columns of the edge blocks, in place ----
    {
        int c0 = 0;
        for (int s = 0; s < t.nsrc; s++) {
            QrSrc q = src[s];
            int w = csize[q.nbr];
            int a = max(c0, c_lo), b = min(c0 + w, c_hi);
            if (a < b) {
                int nc = b - a, off = a - c0;
                int tot = rank * nc;
                if (!q.transposed) {
                    double* dp = q.blk + (size_t)off * q.ld;
                    for (int e = tid; e < tot; e += NT) {
                        int i = e % rank, j = e / rank;
                        int p = pos[a + j];
                        dp[i + (size_t)j * q.ld] = (p >= rank || i <= p) ? Wl[i + (size_t)(a + j - c_lo) * rows] : 0.0;
                    }
                } else {
                    double* dp = q.blk + off;
                    for (int e = tid; e < tot; e += NT) {
                        int j = e % nc, i = e / nc;
                        int p = pos[a + j];
                        dp[j + (size_t)i * q.ld] = (p >= rank || i <= p) ? Wl[i + (size_t)(a + j - c_lo) * rows] : 0.0;
                    }
                }
            }
            c0 += w;
        }
    }
    if (crank == 0 && tid == 0) csize[t.cluster] = rank;
}

template <int G, int NT>
void launch_one(const QrTask* t, int nt, const QrSrc* s, int* csize, double tol, int smem, cudaStream_t st) {
    auto kern = rrqr_cluster_kernel<G, NT>;
    static int configured_smem = -1;
    if (smem > configured_smem) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (G > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        configured_smem = smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(nt * G);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = G;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, t, s, csize, tol, smem);
}

}  // namespace

int rrqr_max_smem() { return 216 * 1024; }

void launch_rrqr(const QrTask* t, int nt, const QrSrc* s, int* csize, double tol, int G, int nthreads, int smem,
                 cudaStream_t st) {
    if (nt <= 0) return;
    if (nthreads <= 128) {
        launch_one<1, 128>(t, nt, s, csize, tol, smem, st);
        return;
    }
    switch (G) {
        case 1: launch_one<1, 512>(t, nt, s, csize, tol, smem, st); break;
        case 2: launch_one<2, 512>(t, nt, s, csize, tol, smem, st); break;
        case 4: launch_one<4, 512>(t, nt, s, csize, tol, smem, st); break;
        case 8: launch_one<8, 512>(t, nt, s, csize, tol, smem, st); break;
        default: launch_one<16, 512>(t, nt, s, csize, tol, smem, st); break;
    }
}

}  // namespace spand
