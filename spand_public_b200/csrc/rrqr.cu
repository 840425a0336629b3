// Batched truncated column-pivoted QR (the sparsification kernel) for sm_100a.
//
// Replaces, for every cluster of one wavefront at once, the reference sequence (citations relative to
// /root/reference): assemble_Asn src/tree.cpp:1189-1224 -> LAPACKE_dgeqp3 src/util.cpp:383-394 ->
// choose_rank src/util.cpp:434-452 -> triu(R[:rank,:]) P^T src/tree.cpp:1334-1335 -> scatter
// src/tree.cpp:1004-1046, plus the (V, tau) of the Orthogonal op src/tree.cpp:1322-1331.
//
// Mapping to the machine: one team per matrix, a team being a thread-block cluster of G CTAs (G = 1..16). The
// gathered panel W = [A_s,n ... (A_n,s)^T ...] (rows x cols, cols >> rows) is distributed by column slabs over
// the CTAs and stays resident in (distributed) shared memory; panels that do not fit stay in the L2-resident
// scratch arena. The factorization is *blocked* like LAPACK's dlaqps: inside a block of nb <= 16 steps the
// panel is not updated; each step reads the slab once (F(:,j) = tau W^T v), fixes F with the small V^T v
// correction, updates only row k of the panel (that is what the partial-norm downdate needs) and the trailing
// rows receive one rank-nb update W -= V F^T per block.
//
// One cluster barrier per Householder step: every CTA *speculatively* builds the reflector of its own best
// pivot candidate and writes it, with the candidate record, into a slot of every CTA of the cluster (DSMEM);
// after the barrier all CTAs select the same winner and find its reflector already in local shared memory.
//
// Columns are never physically swapped: LAPACK's swap sequence is tracked as virtual positions (pos[c]) so that
// idamax's first-index tie-breaking (dlaqp2/dlaqps) is reproduced exactly. Partial column norms are carried
// squared: the dlaqps downdate vn1 <- vn1 sqrt(max(0, 1 - (|a|/vn1)^2)) becomes n1 <- max(0, n1 - a^2) and its
// safeguard tmp2 <= sqrt(eps) becomes n1_new <= sqrt(eps) n2; a flagged column gets its exact norm at once
// (column minus its pending block update) instead of closing the block. The factorization stops at the first
// pivot with |R_kk| / |R_00| < tol, which is exactly geqp3 followed by choose_rank.
#include <cooperative_groups.h>

#include <cfloat>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>
#include <algorithm>

#include "kernels.cuh"

namespace cg = cooperative_groups;

// CTAs per SM the streaming shape (256 threads, panel in global memory) is compiled for: 4 -> 64 registers per thread
#ifndef SPAND_STREAM_MINB
#define SPAND_STREAM_MINB 4
#endif

namespace spand {

// Phase timers of the RRQR kernel (debug builds with -DSPAND_RRQR_TIMING only: scripts/rrqr_phases.py). Thread 0 of
// every CTA accumulates clock64() deltas per phase and adds them to this table when the CTA retires.
__device__ unsigned long long g_qr_phase[48];  // [shape class: smem panel / streaming 256 / global 512][16]
#ifdef SPAND_RRQR_TIMING
#define QT_DECL unsigned long long qt_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long qt_t = clock64();
#define QT(i)                                \
    if (tid == 0) {                          \
        const long long qt_n = clock64();    \
        qt_acc[i] += (unsigned long long)(qt_n - qt_t); \
        qt_t = qt_n;                         \
    }
#define QT_FLUSH                                                         \
    if (tid == 0) {                                                      \
        for (int qi = 0; qi < 10; qi++) atomicAdd(&g_qr_phase[QT_CLASS * 16 + qi], qt_acc[qi]); \
        atomicAdd(&g_qr_phase[QT_CLASS * 16 + 10], 1ull);                                \
    }
// hot / cold statistics: [11] steps, [12] blocks closed full, [13] blocks closed early, [14] hot columns at the
// classifications, [15] unpivoted columns at the classifications
#define QC(i, v) \
    if (tid == 0) atomicAdd(&g_qr_phase[QT_CLASS * 16 + (i)], (unsigned long long)(v));
#else
#define QT_DECL
#define QT(i)
#define QT_FLUSH
#define QC(i, v)
#endif

namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ double group_sum(double v, int L) {
    for (int o = L >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

template <int NT>
__device__ __forceinline__ double block_sum_d(double v, double* red) {
    v = group_sum(v, 32);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < NT / 32; w++) s += red[w];
    return s;
}

struct Cand {  // local pivot candidate
    double val;
    int pos, col;
};

struct CandRec {  // what a CTA tells the cluster about its candidate for the next step
    double val;   // squared partial norm, < 0: none
    double beta, tau;
    int pos, col;
    int stop, pad;
};

// D = C + A B on one 8 x 8 x 4 FP64 tensor-core tile: a = A[lane / 4][lane % 4], b = B[lane % 4][lane / 4],
// (c0, c1) = C[lane / 4][2 (lane % 4) + {0, 1}]
__device__ __forceinline__ void dmma_f64(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ bool better(double v, int p, double bv, int bp) { return v > bv || (v == bv && p < bp); }

__device__ __forceinline__ Cand warp_best(Cand b, int width) {
    for (int o = width >> 1; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(FULL, b.val, o);
        int op = __shfl_xor_sync(FULL, b.pos, o);
        int oc = __shfl_xor_sync(FULL, b.col, o);
        if (better(ov, op, b.val, b.pos)) b = Cand{ov, op, oc};
    }
    return b;
}

__host__ __device__ constexpr int ceil_pow2(int x) { int p = 1; while (p < x) p *= 2; return p; }

// QNB: compile-time bound of the block size (t.nb <= QNB); MINB: CTAs per SM the register budget is sized for
template <int G, int NT, bool SMEM_PANEL, int QNB, int MINB>
__global__ void __launch_bounds__(NT, MINB) rrqr_blocked_kernel(const QrTask* __restrict__ tasks,
                                                                const QrSrc* __restrict__ srcs, int* csize, double tol) {
    constexpr int NW = NT / 32;
    constexpr int NC = 2;  // columns a lane group works on at once (shares the v loads)
    const int task_id = blockIdx.x / G;
    const int crank = blockIdx.x % G;
    const QrTask t = tasks[task_id];
    const QrSrc* src = srcs + t.src0;
    const int rows = t.rows;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    QT_DECL
    constexpr int QT_CLASS = SMEM_PANEL ? 0 : (NT == 256 ? 1 : 2);
    (void)QT_CLASS;

    __shared__ double red[2][NW];
    __shared__ double alpha_s[2];
    __shared__ Cand redc[2][NW];
    __shared__ CandRec rec[2][G];
    __shared__ double aux[QNB];
    extern __shared__ __align__(16) double dsm[];

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
    const int cpcm = (t.maxcols + G - 1) / G;  // layout bound used by the host when sizing shared memory
    const int cpce = (cpcm + 3) & ~3;
    const int c_lo = min(cols, crank * cpc), c_hi = min(cols, c_lo + cpc);
    const int ncl = c_hi - c_lo;
    const int ld = t.ld, ldv = (rows + 1) & ~1, L = t.L, FLD = t.nb | 1;
    const int ngroups = NT / L, grp = tid / L, lig = tid & (L - 1);
    const int npair = ldv >> 1;

    double* Vs = dsm;                                        // ldv x nb   reflectors of the current block
    double* Fs = Vs + (size_t)ldv * t.nb;                    // cpcm x FLD
    double* nq1 = Fs + (((size_t)cpcm * FLD + 1) & ~(size_t)1);  // squared partial norms
    double* nq2 = nq1 + cpce;                                // squared norms at the last exact computation
    double* slots = nq2 + cpce;                              // 2 x ldv: own candidate reflectors (double buffered)
    int* pos = (int*)(slots + (size_t)2 * ldv);              // virtual position of every local column
    double* slab = (double*)(pos + cpce);
    double* P;                                               // local slab, column cl at P[cl * ld]
    if constexpr (SMEM_PANEL) P = slab;
    else P = t.W + (size_t)c_lo * ld;

    cg::cluster_group cluster = cg::this_cluster();
    auto csync = [&]() {
        if constexpr (G > 1) cluster.sync();
        else __syncthreads();
    };
    const double tol3z = sqrt(DBL_EPSILON);
    double r00 = 0.0;

    // Block-wide best of the per-thread candidates (one block barrier); every thread returns the same result.
    auto block_best = [&](Cand b, int par) {
        b = warp_best(b, 32);
        if (lane == 0) redc[par][warp] = b;
        __syncthreads();
        Cand bb = (lane < NW) ? redc[par][lane] : Cand{-1.0, INT_MAX, -1};
        bb = warp_best(bb, ceil_pow2(NW));
        bb.val = __shfl_sync(FULL, bb.val, 0);
        bb.pos = __shfl_sync(FULL, bb.pos, 0);
        bb.col = __shfl_sync(FULL, bb.col, 0);
        return bb;
    };
    // Speculative reflector of the local candidate `lb` for step kn (index jn inside the current block): the vector
    // stays in the own slot [kn & 1] (siblings pull it through DSMEM if it wins), its record goes to every CTA.
    auto propose = [&](Cand lb, int kn, int jn) {
        const int buf = kn & 1;
        double* mine = slots + (size_t)buf * ldv;
        CandRec r;
        r.val = -1.0;
        r.beta = r.tau = 0.0;
        r.pos = INT_MAX;
        r.col = -1;
        r.stop = 0;
        r.pad = 0;
        if (lb.col >= 0) {
            const int pl = lb.col - c_lo;
            const double* pc = P + (size_t)pl * ld;
            const double* fr = Fs + (size_t)pl * FLD;
            double ss = 0.0;
            for (int i = kn + tid; i < rows; i += NT) {
                double u = pc[i];
                for (int tt = 0; tt < jn; tt++) u -= Vs[i + (size_t)tt * ldv] * fr[tt];
                mine[i] = u;
                if (i > kn) ss += u * u;
                else alpha_s[buf] = u;
            }
            ss = group_sum(ss, 32);
            if (lane == 0) red[buf][warp] = ss;
            __syncthreads();
            ss = (lane < NW) ? red[buf][lane] : 0.0;
            ss = group_sum(ss, ceil_pow2(NW));
            ss = __shfl_sync(FULL, ss, 0);
            const double alpha = alpha_s[buf];
            double beta, tau, scal;
            if (ss == 0.0) {
                beta = alpha;
                tau = 0.0;
                scal = 0.0;
            } else {
                beta = -copysign(sqrt(fma(alpha, alpha, ss)), alpha);
                tau = (beta - alpha) / beta;
                scal = 1.0 / (alpha - beta);
            }
            const double ref = (kn == 0) ? fabs(beta) : r00;
            r.val = lb.val;
            r.pos = lb.pos;
            r.col = lb.col;
            r.beta = beta;
            r.tau = tau;
            r.stop = (tol != 0.0 && !(fabs(beta) / ref >= tol)) ? 1 : 0;
            for (int i = kn + tid; i < rows; i += NT) mine[i] = (i == kn) ? 1.0 : mine[i] * scal;
        }
        if constexpr (G > 1) {
            if (warp == 0 && lane < G) cluster.map_shared_rank(&rec[0][0], lane)[buf * G + crank] = r;
        } else {
            if (tid == 0) rec[buf][0] = r;
        }
    };

    // ---- gather the own column range [c_lo, c_hi); padding rows [rows, ld) are zeroed ----
    {
        int c0 = 0;
        for (int s = 0; s < t.nsrc; s++) {
            QrSrc q = src[s];
            int w = csize[q.nbr];
            int a = max(c0, c_lo), b = min(c0 + w, c_hi);
            if (a < b) {
                int nc = b - a, off = a - c0;
                int tot = rows * nc;
                double* dst = P + (size_t)(a - c_lo) * ld;
                if (!q.transposed) {
                    const double* sp = q.blk + (size_t)off * q.ld;
                    for (int e = tid; e < tot; e += NT) {
                        int i = e % rows, j = e / rows;
                        dst[i + (size_t)j * ld] = sp[i + (size_t)j * q.ld];
                    }
                } else {
                    const double* sp = q.blk + off;  // block is w x rows
                    for (int e = tid; e < tot; e += NT) {
                        int j = e % nc, i = e / nc;
                        dst[i + (size_t)j * ld] = sp[j + (size_t)i * q.ld];
                    }
                }
            }
            c0 += w;
        }
        const int padr = ld - rows;
        for (int e = tid; e < padr * ncl; e += NT) P[rows + e % padr + (size_t)(e / padr) * ld] = 0.0;
    }
    __syncthreads();
    // ---- initial squared column norms, identity positions, first candidates ----
    {
        Cand best{-1.0, INT_MAX, -1};
        for (int base = 0; base < ncl; base += ngroups) {
            int cl = base + grp;
            bool valid = cl < ncl;
            const double* cj = P + (size_t)(valid ? cl : 0) * ld;
            double s = 0.0;
            if (valid)
                for (int i = lig; i < rows; i += L) s += cj[i] * cj[i];
            s = group_sum(s, L);
            if (valid && lig == 0) {
                nq1[cl] = s;
                nq2[cl] = s;
                pos[cl] = c_lo + cl;
                if (better(s, c_lo + cl, best.val, best.pos)) best = Cand{s, c_lo + cl, c_lo + cl};
            }
        }
        Cand lb = block_best(best, 1);
        propose(lb, 0, 0);
    }
    csync();
    QT(0)  // gather + initial norms + first proposal

    int rank = mn;
    int k0 = 0, nb = min(t.nb, mn);
    for (int k = 0; k < mn; k++) {
        const int j = k - k0, par = k & 1;
        // ---- every CTA selects the same winner among the G proposals ----
        int wi = 0;
        if constexpr (G > 2) {
            Cand c = (lane < G) ? Cand{rec[par][lane].val, rec[par][lane].pos, lane} : Cand{-2.0, INT_MAX, 0};
            wi = warp_best(c, ceil_pow2(G)).col;
            wi = __shfl_sync(FULL, wi, 0);
        } else if constexpr (G == 2) {
            if (better(rec[par][1].val, rec[par][1].pos, rec[par][0].val, rec[par][0].pos)) wi = 1;
        }
        const CandRec win = rec[par][wi];
        const int pcol = win.col, ppos = win.pos;
        if (pcol < 0) {  // no admissible column (NaN norms): stop here
            rank = k;
            break;
        }
        const double beta = win.beta, tau = win.tau;
        if (k == 0) r00 = fabs(beta);
        if (win.stop) {
            rank = k;
            break;
        }
        // ---- pull the winner's reflector into column j of the block (zero outside [k, rows)) ----
        double* vj = Vs + (size_t)j * ldv;
        {
            const double* wv = slots + (size_t)par * ldv;
            if constexpr (G > 1) wv = cluster.map_shared_rank(wv, wi);
            const bool own = (crank == wi);
            for (int i = tid; i < ldv; i += NT) {
                double v = (i >= k && i < rows) ? wv[i] : 0.0;
                vj[i] = v;
                if (own && i > k && i < rows) t.V[i + (size_t)k * rows] = v;
            }
            if (own && tid == 0) {
                P[k + (size_t)(pcol - c_lo) * ld] = beta;
                t.tau[k] = tau;
            }
        }
        __syncthreads();
        QT(1)  // winner selection + reflector pull
        // ---- aux = V(k:, 0:j)^T v ----
        if (j > 0) {
            for (int tt = warp; tt < j; tt += NW) {
                const double* vt = Vs + (size_t)tt * ldv;
                double s = 0.0;
                for (int i = k + lane; i < rows; i += 32) s += vt[i] * vj[i];
                s = group_sum(s, 32);
                if (lane == 0) aux[tt] = s;
            }
            __syncthreads();
        }
        QT(2)  // aux = V^T v
        // ---- F(:, j), row k of the panel, partial norms, local candidate for step k + 1 ----
        Cand best{-1.0, INT_MAX, -1};
        const double2* v2 = reinterpret_cast<const double2*>(vj);
        if constexpr (!SMEM_PANEL) {
            // Panel in global memory (L == 32): a warp takes WB columns at once. All loads of the batch are in flight
            // together, the WB x 32 partial sums are transpose-reduced with shuffles, and the scalar epilogue of the WB
            // columns runs lane-parallel (column c on lanes c * LPC .. c * LPC + LPC - 1).
            // WB consecutive columns per warp pass (one base pointer + WB small offsets, so that the WB loads of a
            // row chunk are issued back to back into distinct registers: the memory-level parallelism of this loop
            // is what bounds the kernel). Columns already pivoted are streamed too (a fraction k / cols).
            constexpr int WB = (MINB == 1) ? 8 : 4;
            constexpr int LPC = 32 / WB;  // lanes per column in the scalar epilogue
            const int ld2 = ld >> 1;
            for (int first = warp * WB; first < ncl; first += WB * NW) {
                const int myc = lane / LPC;
                int my_cl = first + myc, my_p = -1;
                const bool my_valid = my_cl < ncl;
                if (my_valid) {
                    my_p = pos[my_cl];
                    if (c_lo + my_cl == pcol) {
                        my_p = k;
                        if ((lane & (LPC - 1)) == 0) pos[my_cl] = k;
                    } else if (my_p == k) {
                        my_p = ppos;
                        if ((lane & (LPC - 1)) == 0) pos[my_cl] = ppos;
                    }
                }
                const bool my_act = my_valid && my_p > k;
                if (!my_act) my_cl = 0;
                if (__ballot_sync(FULL, my_act) == 0) continue;
                double s[WB];
                const int nhere = min(WB, ncl - first);
                int i2 = (k >> 1) + lane;
                if constexpr (MINB != 4) {
                    // 128- / 80-register shapes: one pointer per column advanced by a constant (no per-load index
                    // arithmetic), two row chunks = 2 WB loads of 16 bytes in flight per lane.
                    // Columns past the slab re-read the last one.
                    const double2* pc[WB];
#pragma unroll
                    for (int c = 0; c < WB; c++) {
                        s[c] = 0.0;
                        pc[c] = reinterpret_cast<const double2*>(P) + ((size_t)(first + min(c, nhere - 1)) * ld2 + i2);
                    }
                    const double2* vp = v2 + i2;
                    int rem = npair - i2;  // row pairs left for this lane (stride 32)
                    for (; rem > 32; rem -= 64) {
                        const double2 v0 = vp[0], v1 = vp[32];
                        double2 a0[WB], a1[WB];
#pragma unroll
                        for (int c = 0; c < WB; c++) {
                            a0[c] = pc[c][0];
                            a1[c] = pc[c][32];
                            pc[c] += 64;
                        }
                        vp += 64;
#pragma unroll
                        for (int c = 0; c < WB; c++) {
                            s[c] = fma(a0[c].x, v0.x, s[c]);
                            s[c] = fma(a0[c].y, v0.y, s[c]);
                            s[c] = fma(a1[c].x, v1.x, s[c]);
                            s[c] = fma(a1[c].y, v1.y, s[c]);
                        }
                    }
                    if (rem > 0) {
                        const double2 vv = vp[0];
                        double2 a0[WB];
#pragma unroll
                        for (int c = 0; c < WB; c++) a0[c] = pc[c][0];
#pragma unroll
                        for (int c = 0; c < WB; c++) {
                            s[c] = fma(a0[c].x, vv.x, s[c]);
                            s[c] = fma(a0[c].y, vv.y, s[c]);
                        }
                    }
                } else {
                    // 64-register shape: 32-bit column offsets from one base keep the four loads of a row chunk
                    // back to back (with per-column pointers the register allocator serialises them)
                    int coff[WB];
#pragma unroll
                    for (int c = 0; c < WB; c++) {
                        s[c] = 0.0;
                        coff[c] = min(c, nhere - 1) * ld2;
                    }
                    const double2* cb = reinterpret_cast<const double2*>(P) + (size_t)first * ld2;
                    for (; i2 < npair; i2 += 32) {
                        const double2 vv = v2[i2];
                        double2 a0[WB];
#pragma unroll
                        for (int c = 0; c < WB; c++) a0[c] = cb[coff[c] + i2];
#pragma unroll
                        for (int c = 0; c < WB; c++) {
                            s[c] = fma(a0[c].x, vv.x, s[c]);
                            s[c] = fma(a0[c].y, vv.y, s[c]);
                        }
                    }
                }
                int bit = 16;
#pragma unroll
                for (int h = WB / 2; h >= 1; h >>= 1, bit >>= 1) {
                    const bool up = (lane & bit) != 0;
#pragma unroll
                    for (int i = 0; i < h; i++) {
                        const double keep = up ? s[i + h] : s[i];
                        const double send = up ? s[i] : s[i + h];
                        s[i] = keep + __shfl_xor_sync(FULL, send, bit);
                    }
                }
                double tot = s[0];
#pragma unroll
                for (int o = LPC / 2; o >= 1; o >>= 1) tot += __shfl_xor_sync(FULL, tot, o);
                double my_f = 0.0, my_ak = 0.0, my_n1 = 0.0, my_newn = 0.0;
                bool my_need = false;
                if (my_act) {
                    const double* cj = P + (size_t)my_cl * ld;
                    const double* fr = Fs + (size_t)my_cl * FLD;
                    double corr = 0.0, rk = cj[k];
                    for (int tt = 0; tt < j; tt++) {
                        const double ft = fr[tt];
                        corr = fma(ft, aux[tt], corr);
                        rk = fma(-Vs[k + (size_t)tt * ldv], ft, rk);
                    }
                    my_f = tau * (tot - corr);
                    my_ak = rk - my_f;
                    my_n1 = nq1[my_cl];
                    if (my_n1 != 0.0) {
                        my_newn = fmax(0.0, my_n1 - my_ak * my_ak);
                        my_need = my_newn <= tol3z * nq2[my_cl];
                    }
                }
                unsigned nm = __ballot_sync(FULL, my_need && (lane & (LPC - 1)) == 0);
                while (nm) {
                    const int srcl = __ffs(nm) - 1;
                    nm &= nm - 1;
                    const int ccl = __shfl_sync(FULL, my_cl, srcl);
                    const double ff = __shfl_sync(FULL, my_f, srcl);
                    const double* cj = P + (size_t)ccl * ld;
                    const double* fr = Fs + (size_t)ccl * FLD;
                    double q = 0.0;
                    for (int i = k + 1 + lane; i < rows; i += 32) {
                        double u = cj[i];
                        for (int tt = 0; tt < j; tt++) u -= Vs[i + (size_t)tt * ldv] * fr[tt];
                        u -= vj[i] * ff;
                        q += u * u;
                    }
                    q = group_sum(q, 32);
                    if ((lane / LPC) == (srcl / LPC)) my_newn = q;
                }
                if (my_act && (lane & (LPC - 1)) == 0) {
                    Fs[(size_t)my_cl * FLD + j] = my_f;
                    P[k + (size_t)my_cl * ld] = my_ak;
                    if (my_n1 != 0.0) {
                        nq1[my_cl] = my_newn;
                        if (my_need) nq2[my_cl] = my_newn;
                    }
                    if (better(my_newn, my_p, best.val, best.pos)) best = Cand{my_newn, my_p, c_lo + my_cl};
                }
            }
        } else {
            for (int base = 0; base < ncl; base += NC * ngroups) {
                int cl[NC], p[NC];
                bool act[NC], need[NC];
                double s[NC], f[NC], a_k[NC], n1[NC], newn[NC];
                const double2* c2[NC];
                bool any_act = false;
    #pragma unroll
                for (int c = 0; c < NC; c++) {
                    cl[c] = base + c * ngroups + grp;
                    const bool valid = cl[c] < ncl;
                    p[c] = -1;
                    if (valid) {
                        p[c] = pos[cl[c]];
                        if (c_lo + cl[c] == pcol) {
                            p[c] = k;
                            if (lig == 0) pos[cl[c]] = k;
                        } else if (p[c] == k) {  // virtual swap: the column at position k takes the pivot's place
                            p[c] = ppos;
                            if (lig == 0) pos[cl[c]] = ppos;
                        }
                    }
                    act[c] = valid && p[c] > k;
                    any_act |= act[c];
                    if (!act[c]) cl[c] = 0;
                    c2[c] = reinterpret_cast<const double2*>(P + (size_t)cl[c] * ld);
                    s[c] = 0.0;
                    need[c] = false;
                    f[c] = a_k[c] = n1[c] = newn[c] = 0.0;
                }
                if (any_act) {
                    for (int i2 = (k >> 1) + lig; i2 < npair; i2 += L) {
                        const double2 vv = v2[i2];
    #pragma unroll
                        for (int c = 0; c < NC; c++)
                            if (act[c]) {
                                const double2 a = c2[c][i2];
                                s[c] = fma(a.x, vv.x, s[c]);
                                s[c] = fma(a.y, vv.y, s[c]);
                            }
                    }
                }
                bool any_need = false;
    #pragma unroll
                for (int c = 0; c < NC; c++) {
                    s[c] = group_sum(s[c], L);
                    if (act[c]) {
                        const double* cj = P + (size_t)cl[c] * ld;
                        const double* fr = Fs + (size_t)cl[c] * FLD;
                        double corr = 0.0, rk = cj[k];
                        for (int tt = 0; tt < j; tt++) {
                            const double ft = fr[tt];
                            corr = fma(ft, aux[tt], corr);
                            rk = fma(-Vs[k + (size_t)tt * ldv], ft, rk);
                        }
                        f[c] = tau * (s[c] - corr);
                        a_k[c] = rk - f[c];
                        n1[c] = nq1[cl[c]];
                        if (n1[c] != 0.0) {
                            newn[c] = fmax(0.0, n1[c] - a_k[c] * a_k[c]);
                            need[c] = newn[c] <= tol3z * nq2[cl[c]];
                        }
                    }
                    any_need |= need[c];
                }
                if (__any_sync(FULL, any_need)) {
    #pragma unroll
                    for (int c = 0; c < NC; c++) {
                        if (!__any_sync(FULL, need[c])) continue;
                        double q = 0.0;
                        if (need[c]) {
                            const double* cj = P + (size_t)cl[c] * ld;
                            const double* fr = Fs + (size_t)cl[c] * FLD;
                            for (int i = k + 1 + lig; i < rows; i += L) {
                                double u = cj[i];
                                for (int tt = 0; tt < j; tt++) u -= Vs[i + (size_t)tt * ldv] * fr[tt];
                                u -= vj[i] * f[c];
                                q += u * u;
                            }
                        }
                        q = group_sum(q, L);
                        if (need[c]) newn[c] = q;
                    }
                }
    #pragma unroll
                for (int c = 0; c < NC; c++)
                    if (act[c] && lig == 0) {
                        Fs[(size_t)cl[c] * FLD + j] = f[c];
                        P[k + (size_t)cl[c] * ld] = a_k[c];
                        if (n1[c] != 0.0) {
                            nq1[cl[c]] = newn[c];
                            if (need[c]) nq2[cl[c]] = newn[c];
                        }
                        if (better(newn[c], p[c], best.val, best.pos)) best = Cand{newn[c], p[c], c_lo + cl[c]};
                    }
            }
        }
        if (k + 1 >= mn) break;  // factorization complete, rank = mn
        QT(3)  // sweep (warp 0's share)
        Cand lb = block_best(best, par);  // contains a block barrier: F and row k are visible below
        QT(4)  // block_best = waiting for the slowest warp of the sweep
        // ---- end of block: trailing rows of the own active columns  W -= V F^T ----
        int jn = j + 1;
        if (j == nb - 1) {
            const int kend = k + 1;
            if constexpr (!SMEM_PANEL) {
                // Panel in global memory: the rank-nb update W(kend:, active) -= V F^T runs on the FP64 tensor cores
                // (DMMA m8n8k4). A warp owns a strip of 8 columns: the B fragments (F of those columns) are loaded
                // once, then the strip is swept in tiles of 8 rows (two tiles in flight): A = -V from shared memory,
                // C read-modify-written in place. Rows above kend (final R entries) and columns that are already
                // pivots are computed but never stored.
                const int g = lane >> 2, q = lane & 3;
                const int r_lo = kend & ~7;
                for (int first = warp * 8; first < ncl; first += 8 * NW) {
                    unsigned am = 0;
#pragma unroll
                    for (int c = 0; c < 8; c++)
                        if (first + c < ncl && pos[first + c] >= kend) am |= 1u << c;
                    if (am == 0) continue;
                    double bf[QNB / 4];
#pragma unroll
                    for (int kk = 0; kk < QNB / 4; kk++) {
                        const int tt = kk * 4 + q;
                        bf[kk] = (tt < nb && first + g < ncl) ? Fs[(size_t)(first + g) * FLD + tt] : 0.0;
                    }
                    const bool actA = (am >> (2 * q)) & 1u, actB = (am >> (2 * q + 1)) & 1u;
                    double* colA = P + (size_t)(first + 2 * q) * ld;
                    double* colB = colA + ld;
                    for (int r = r_lo; r < rows; r += 16) {
                        const int row0 = r + g, row1 = r + 8 + g;
                        const bool ok0 = row0 >= kend && row0 < rows, ok1 = row1 >= kend && row1 < rows;
                        double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
                        if (ok0 && actA) c00 = colA[row0];
                        if (ok0 && actB) c01 = colB[row0];
                        if (ok1 && actA) c10 = colA[row1];
                        if (ok1 && actB) c11 = colB[row1];
#pragma unroll
                        for (int kk = 0; kk < QNB / 4; kk++) {
                            const int tt = kk * 4 + q;
                            if (kk * 4 < nb) {  // uniform
                                const double a0 = (tt < nb && row0 < ldv) ? -Vs[row0 + (size_t)tt * ldv] : 0.0;
                                const double a1 = (tt < nb && row1 < ldv) ? -Vs[row1 + (size_t)tt * ldv] : 0.0;
                                dmma_f64(c00, c01, a0, bf[kk]);
                                dmma_f64(c10, c11, a1, bf[kk]);
                            }
                        }
                        if (ok0 && actA) colA[row0] = c00;
                        if (ok0 && actB) colB[row0] = c01;
                        if (ok1 && actA) colA[row1] = c10;
                        if (ok1 && actB) colB[row1] = c11;
                    }
                }
            } else
            for (int base = 0; base < ncl; base += ngroups) {
                const int cl = base + grp;
                if (cl >= ncl || pos[cl] < kend) continue;
                double* cj = P + (size_t)cl * ld;
                double f[QNB];
#pragma unroll
                for (int tt = 0; tt < QNB; tt++) f[tt] = (tt < nb) ? Fs[(size_t)cl * FLD + tt] : 0.0;
                for (int i = kend + lig; i < rows; i += L) {
                    double a = cj[i];
#pragma unroll
                    for (int tt = 0; tt < QNB; tt++)
                        if (tt < nb) a -= Vs[i + (size_t)tt * ldv] * f[tt];
                    cj[i] = a;
                }
            }
            k0 = kend;
            nb = min(t.nb, mn - k0);
            jn = 0;
            __syncthreads();
        }
        QT(5)  // end-of-block trailing update
        propose(lb, k + 1, jn);
        QT(6)  // speculative reflector + record broadcast
        csync();
        QT(7)  // cluster barrier
    }
    // All CTAs leave the loop at the same step. One more barrier so that no CTA exits (or starts scattering over
    // its panel) while a sibling may still be pulling from its slots.
    csync();
    if (rank >= rows) {  // nothing to do (tree.cpp:1317-1319); csize unchanged
        QT_FLUSH
        return;
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
                        int i = e % rank, jj = e / rank;
                        int p = pos[a + jj - c_lo];
                        dp[i + (size_t)jj * q.ld] = (p >= rank || i <= p) ? P[i + (size_t)(a + jj - c_lo) * ld] : 0.0;
                    }
                } else {
                    double* dp = q.blk + off;
                    for (int e = tid; e < tot; e += NT) {
                        int jj = e % nc, i = e / nc;
                        int p = pos[a + jj - c_lo];
                        dp[jj + (size_t)i * q.ld] = (p >= rank || i <= p) ? P[i + (size_t)(a + jj - c_lo) * ld] : 0.0;
                    }
                }
            }
            c0 += w;
        }
    }
    if (crank == 0 && tid == 0) csize[t.cluster] = rank;
    QT(8)  // scatter
    QT_FLUSH
}


// ------------------------------------------------------------------------------------------------
// Hot / cold truncated QRCP for panels in global memory.
//
// The blocked algorithm above sweeps EVERY unpivoted column once per Householder step (F(:, j) = tau W^T v), which
// is what bounds it (BLAS-2: the panel is re-read from L2 / HBM rank times). But a column only has to be current
// while it can win the pivot search. Partial column norms never grow, so the norm a column had at the start of a
// block is an upper bound during the block. At the start of a block the columns whose squared norm is at least
// theta^2 times the pivot's are "hot": they are swept every step exactly like dlaqps does (same arithmetic, same
// downdated norms). The others are "cold": they are not touched during the block, and the block stays open only
// while the best hot candidate is strictly larger than the largest cold bound (cluster-wide) - otherwise it is
// closed early, so the pivot sequence is the one of the greedy algorithm. When a block of jc reflectors closes, the
// F rows of the cold columns are computed in ONE pass on the FP64 tensor cores, F_cold = W_cold^T V T (the
// compact-WY form of the same recurrence: T(0:j, j) = -tau_j T(0:j, 0:j) V(:, 0:j)^T v_j, T(j, j) = tau_j), then
// every unpivoted column receives W -= V F^T (cold columns from the first row of the block, which also produces
// their R entries) and the cold columns get their exact new norm from the registers of that update. Measured on
// the oracle (SPAND_ORACLE_HOTCOLS hook) with theta = 1/2 the hot columns carry 4-7 % of the bytes the full sweep
// reads; a cold column costs two passes per block instead of nb + 2.
// ------------------------------------------------------------------------------------------------
struct HcRec {      // what a CTA tells the cluster about its candidate for the next step
    double val;     // squared partial norm of the candidate, < 0: none
    double beta, tau;
    double cold;    // largest squared norm bound among the CTA's cold columns, < 0: none
    int pos, col;
    int stop, pad;
};

template <int G, int NT, int QNB, int MINB>
__global__ void __launch_bounds__(NT, MINB) rrqr_hc_kernel(const QrTask* __restrict__ tasks,
                                                           const QrSrc* __restrict__ srcs, int* csize, double tol,
                                                           double theta2) {
    constexpr int NW = NT / 32;
    constexpr int WB = 4;          // hot columns a warp sweeps at once
    constexpr int LPC = 32 / WB;   // lanes per column in the scalar epilogue
    constexpr int MT = QNB / 8;    // 8-row tiles of Y = V^T W
    const int task_id = blockIdx.x / G;
    const int crank = blockIdx.x % G;
    const QrTask t = tasks[task_id];
    const QrSrc* src = srcs + t.src0;
    const int rows = t.rows;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    QT_DECL
    constexpr int QT_CLASS = (NT == 256 ? 1 : 2);
    (void)QT_CLASS;

    __shared__ double red[2][NW];
    __shared__ double alpha_s[2];
    __shared__ Cand redc[2][NW];
    __shared__ HcRec rec[2][G];
    __shared__ double aux[QNB];
    __shared__ double Ts[QNB * QNB];       // compact-WY factor of the open block (upper triangular, row-major)
    __shared__ double ysm[NW][QNB * 8];    // per-warp tile Y[t][col] of the cold pass
    __shared__ double cmax_s[NW];
    __shared__ int nhot_s;
    extern __shared__ __align__(16) double dsm[];

    int cols = 0;
    for (int s = 0; s < t.nsrc; s++) cols += csize[src[s].nbr];
    if (rows == 0) return;
    if (tol >= 1.0 || cols == 0) {
        if (crank == 0 && tid == 0) csize[t.cluster] = 0;
        return;
    }
    const int mn = min(rows, cols);
    const int cpc = (cols + G - 1) / G;
    const int cpcm = (t.maxcols + G - 1) / G;
    const int cpce = (cpcm + 3) & ~3;
    const int c_lo = min(cols, crank * cpc), c_hi = min(cols, c_lo + cpc);
    const int ncl = c_hi - c_lo;
    const int ld = t.ld, ldv = (rows + 1) & ~1, FLD = t.nb | 1;
    const int npair = ldv >> 1, ld2 = ld >> 1;
    const int nbmax = t.nb;

    double* Vs = dsm;                                            // ldv x nb   reflectors of the open block
    double* Fs = Vs + (size_t)ldv * t.nb;                        // cpcm x FLD
    double* nq1 = Fs + (((size_t)cpcm * FLD + 1) & ~(size_t)1);  // squared partial norms (bounds for cold columns)
    double* nq2 = nq1 + cpce;
    double* slots = nq2 + cpce;                                  // 2 x ldv: own candidate reflectors
    int* pos = (int*)(slots + (size_t)2 * ldv);                  // virtual position of every local column
    int* state = pos + cpce;                                     // 1: hot in the open block
    int* hot = state + cpce;                                     // compacted list of the hot local columns
    double* P = t.W + (size_t)c_lo * ld;                         // local slab, column cl at P[cl * ld]

    cg::cluster_group cluster = cg::this_cluster();
    auto csync = [&]() {
        if constexpr (G > 1) cluster.sync();
        else __syncthreads();
    };
    const double tol3z = sqrt(DBL_EPSILON);
    double r00 = 0.0;
    int seq = 0;   // proposals made so far: rec / slots / red / alpha_s are double buffered on its parity
    int bseq = 0;  // block_best calls so far

    auto block_best = [&](Cand b) {
        const int par = (bseq++) & 1;
        b = warp_best(b, 32);
        if (lane == 0) redc[par][warp] = b;
        __syncthreads();
        Cand bb = (lane < NW) ? redc[par][lane] : Cand{-1.0, INT_MAX, -1};
        bb = warp_best(bb, ceil_pow2(NW));
        bb.val = __shfl_sync(FULL, bb.val, 0);
        bb.pos = __shfl_sync(FULL, bb.pos, 0);
        bb.col = __shfl_sync(FULL, bb.col, 0);
        return bb;
    };
    // Speculative reflector of the local candidate for step kn; jn reflectors of the open block are pending on it.
    auto propose = [&](Cand lb, int kn, int jn, double coldmax) {
        const int buf = (seq++) & 1;
        double* mine = slots + (size_t)buf * ldv;
        HcRec r;
        r.val = -1.0;
        r.beta = r.tau = 0.0;
        r.cold = coldmax;
        r.pos = INT_MAX;
        r.col = -1;
        r.stop = 0;
        r.pad = 0;
        if (lb.col >= 0) {
            const int pl = lb.col - c_lo;
            const double* pc = P + (size_t)pl * ld;
            const double* fr = Fs + (size_t)pl * FLD;
            double ss = 0.0;
            for (int i = kn + tid; i < rows; i += NT) {
                double u = pc[i];
                for (int tt = 0; tt < jn; tt++) u -= Vs[i + (size_t)tt * ldv] * fr[tt];
                mine[i] = u;
                if (i > kn) ss += u * u;
                else alpha_s[buf] = u;
            }
            ss = group_sum(ss, 32);
            if (lane == 0) red[buf][warp] = ss;
            __syncthreads();
            ss = (lane < NW) ? red[buf][lane] : 0.0;
            ss = group_sum(ss, ceil_pow2(NW));
            ss = __shfl_sync(FULL, ss, 0);
            const double alpha = alpha_s[buf];
            double beta, tau, scal;
            if (ss == 0.0) {
                beta = alpha;
                tau = 0.0;
                scal = 0.0;
            } else {
                beta = -copysign(sqrt(fma(alpha, alpha, ss)), alpha);
                tau = (beta - alpha) / beta;
                scal = 1.0 / (alpha - beta);
            }
            const double ref = (kn == 0) ? fabs(beta) : r00;
            r.val = lb.val;
            r.pos = lb.pos;
            r.col = lb.col;
            r.beta = beta;
            r.tau = tau;
            r.stop = (tol != 0.0 && !(fabs(beta) / ref >= tol)) ? 1 : 0;
            for (int i = kn + tid; i < rows; i += NT) mine[i] = (i == kn) ? 1.0 : mine[i] * scal;
        }
        if constexpr (G > 1) {
            if (warp == 0 && lane < G) cluster.map_shared_rank(&rec[0][0], lane)[buf * G + crank] = r;
        } else {
            if (tid == 0) rec[buf][0] = r;
        }
    };
    // best unpivoted local column (positions >= kk) by its current squared norm; all columns must be up to date
    auto best_of_all = [&](int kk) {
        Cand best{-1.0, INT_MAX, -1};
        for (int cl = tid; cl < ncl; cl += NT) {
            const int p = pos[cl];
            if (p >= kk) {
                const double v = nq1[cl];
                if (better(v, p, best.val, best.pos)) best = Cand{v, p, c_lo + cl};
            }
        }
        return best;
    };
    // Closes the open block of jc reflectors that started at step kb0: F rows of the cold columns on the tensor cores,
    // W -= V F^T for every unpivoted column, exact norms of the cold columns. Ends with a block barrier.
    auto close_block = [&](int jc, int kb0) {
        const int kend = kb0 + jc;
        const int g = lane >> 2, q = lane & 3;
        double* yw = ysm[warp];
        for (int first = warp * 8; first < ncl; first += 8 * NW) {
            unsigned am = 0, cm = 0;
#pragma unroll
            for (int c = 0; c < 8; c++)
                if (first + c < ncl && pos[first + c] >= kend) {
                    am |= 1u << c;
                    if (!state[first + c]) cm |= 1u << c;
                }
            if (am == 0) continue;
            if (cm) {
                // Y = V^T W for the 8 columns of the strip: lane (g, q) feeds rows r + 2q, r + 2q + 1 of column
                // first + g (B fragments) and of reflector mt * 8 + g (A fragments); two accumulation chains
                double y0[MT][2], y1[MT][2];
#pragma unroll
                for (int mt = 0; mt < MT; mt++) y0[mt][0] = y0[mt][1] = y1[mt][0] = y1[mt][1] = 0.0;
                const bool colok = first + g < ncl;
                const double2* pc2 = reinterpret_cast<const double2*>(P + (size_t)(first + (colok ? g : 0)) * ld);
                // eight 16-byte loads in flight per lane (64 rows of the strip) before the tensor-core work on them
                for (int rb = kb0 & ~7; rb < ldv; rb += 64) {
                    double2 bq[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int rr = rb + 8 * u + 2 * q;
                        bq[u] = make_double2(0.0, 0.0);
                        if (colok && rr < ldv) bq[u] = pc2[rr >> 1];
                    }
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int rr = rb + 8 * u + 2 * q;
                        const bool rok = rr < ldv;
#pragma unroll
                        for (int mt = 0; mt < MT; mt++) {
                            if (mt * 8 < jc) {  // uniform
                                const int tt = mt * 8 + g;
                                double2 a = make_double2(0.0, 0.0);
                                if (tt < jc && rok) a = *reinterpret_cast<const double2*>(Vs + rr + (size_t)tt * ldv);
                                dmma_f64(y0[mt][0], y0[mt][1], a.x, bq[u].x);
                                dmma_f64(y1[mt][0], y1[mt][1], a.y, bq[u].y);
                            }
                        }
                    }
                }
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    yw[(mt * 8 + g) * 8 + 2 * q] = y0[mt][0] + y1[mt][0];
                    yw[(mt * 8 + g) * 8 + 2 * q + 1] = y0[mt][1] + y1[mt][1];
                }
                __syncwarp();
                // F(col, jj) = sum_{t <= jj} Y(t, col) T(t, jj)
                const int col = lane & 7;
                for (int jj = lane >> 3; jj < jc; jj += 4) {
                    double f = 0.0;
                    for (int tt = 0; tt <= jj; tt++) f = fma(yw[tt * 8 + col], Ts[tt * QNB + jj], f);
                    if ((cm >> col) & 1u) Fs[(size_t)(first + col) * FLD + jj] = f;
                }
                __syncwarp();
            }
            // W(lo:, strip) -= V F^T, tiles of 8 rows, two tiles in flight; lo = kb0 for cold columns (their R entries
            // of the block are produced here), kend for hot ones (row k was kept current step by step)
            double bf[QNB / 4];
#pragma unroll
            for (int kk = 0; kk < QNB / 4; kk++) {
                const int tt = kk * 4 + q;
                bf[kk] = (tt < jc && first + g < ncl) ? Fs[(size_t)(first + g) * FLD + tt] : 0.0;
            }
            const bool actA = (am >> (2 * q)) & 1u, actB = (am >> (2 * q + 1)) & 1u;
            const bool cldA = (cm >> (2 * q)) & 1u, cldB = (cm >> (2 * q + 1)) & 1u;
            const int loA = cldA ? kb0 : kend, loB = cldB ? kb0 : kend;
            double* colA = P + (size_t)(first + 2 * q) * ld;
            double* colB = colA + ld;
            double nA = 0.0, nB = 0.0;
            const int r_lo = (cm ? kb0 : kend) & ~7;
            for (int r = r_lo; r < rows; r += 32) {  // four tiles of 8 rows: eight loads in flight per lane
                double cA[4], cB[4];
#pragma unroll
                for (int h = 0; h < 4; h++) {
                    const int row = r + 8 * h + g;
                    cA[h] = (actA && row >= loA && row < rows) ? colA[row] : 0.0;
                    cB[h] = (actB && row >= loB && row < rows) ? colB[row] : 0.0;
                }
#pragma unroll
                for (int kk = 0; kk < QNB / 4; kk++) {
                    const int tt = kk * 4 + q;
                    if (kk * 4 < jc) {  // uniform
#pragma unroll
                        for (int h = 0; h < 4; h++) {
                            const int row = r + 8 * h + g;
                            const double av = (tt < jc && row < ldv) ? -Vs[row + (size_t)tt * ldv] : 0.0;
                            dmma_f64(cA[h], cB[h], av, bf[kk]);
                        }
                    }
                }
#pragma unroll
                for (int h = 0; h < 4; h++) {
                    const int row = r + 8 * h + g;
                    if (actA && row >= loA && row < rows) {
                        colA[row] = cA[h];
                        if (row >= kend) nA = fma(cA[h], cA[h], nA);
                    }
                    if (actB && row >= loB && row < rows) {
                        colB[row] = cB[h];
                        if (row >= kend) nB = fma(cB[h], cB[h], nB);
                    }
                }
            }
            if (cm) {  // exact squared norms of the cold columns (sum over the 8 row groups, fixed order)
#pragma unroll
                for (int o = 4; o <= 16; o <<= 1) {
                    nA += __shfl_xor_sync(FULL, nA, o);
                    nB += __shfl_xor_sync(FULL, nB, o);
                }
                if (g == 0) {
                    if (cldA) nq1[first + 2 * q] = nq2[first + 2 * q] = nA;
                    if (cldB) nq1[first + 2 * q + 1] = nq2[first + 2 * q + 1] = nB;
                }
            }
        }
        __syncthreads();
    };

    // ---- gather the own column range [c_lo, c_hi); padding rows [rows, ld) are zeroed ----
    {
        int c0 = 0;
        for (int s = 0; s < t.nsrc; s++) {
            QrSrc q = src[s];
            int w = csize[q.nbr];
            int a = max(c0, c_lo), b = min(c0 + w, c_hi);
            if (a < b) {
                int nc = b - a, off = a - c0;
                int tot = rows * nc;
                double* dst = P + (size_t)(a - c_lo) * ld;
                if (!q.transposed) {
                    const double* sp = q.blk + (size_t)off * q.ld;
                    for (int e = tid; e < tot; e += NT) {
                        int i = e % rows, j = e / rows;
                        dst[i + (size_t)j * ld] = sp[i + (size_t)j * q.ld];
                    }
                } else {
                    const double* sp = q.blk + off;  // block is w x rows
                    for (int e = tid; e < tot; e += NT) {
                        int j = e % nc, i = e / nc;
                        dst[i + (size_t)j * ld] = sp[j + (size_t)i * q.ld];
                    }
                }
            }
            c0 += w;
        }
        const int padr = ld - rows;
        for (int e = tid; e < padr * ncl; e += NT) P[rows + e % padr + (size_t)(e / padr) * ld] = 0.0;
    }
    __syncthreads();
    // ---- initial squared column norms (one warp per column), identity positions, first candidates ----
    {
        Cand best{-1.0, INT_MAX, -1};
        for (int cl = warp; cl < ncl; cl += NW) {
            const double* cj = P + (size_t)cl * ld;
            double s = 0.0;
            for (int i = lane; i < rows; i += 32) s += cj[i] * cj[i];
            s = group_sum(s, 32);
            if (lane == 0) {
                nq1[cl] = s;
                nq2[cl] = s;
                pos[cl] = c_lo + cl;
                state[cl] = 1;
                if (better(s, c_lo + cl, best.val, best.pos)) best = Cand{s, c_lo + cl, c_lo + cl};
            }
        }
        Cand lb = block_best(best);
        propose(lb, 0, 0, -1.0);
    }
    csync();
    QT(0)  // gather + initial norms + first proposal

    int rank = mn;
    int k0 = 0, j = 0, nhot = 0, k = 0;
    double coldmax_l = -1.0;
    bool flush = false;
    while (k < mn) {
        const int par = (seq - 1) & 1;
        // ---- every CTA selects the same winner among the G proposals, and the largest cold bound ----
        int wi = 0;
        double coldmax_g = rec[par][0].cold;
        if constexpr (G > 2) {
            Cand c = (lane < G) ? Cand{rec[par][lane].val, rec[par][lane].pos, lane} : Cand{-2.0, INT_MAX, 0};
            wi = warp_best(c, ceil_pow2(G)).col;
            wi = __shfl_sync(FULL, wi, 0);
        } else if constexpr (G == 2) {
            if (better(rec[par][1].val, rec[par][1].pos, rec[par][0].val, rec[par][0].pos)) wi = 1;
        }
        if constexpr (G > 1) {
#pragma unroll
            for (int i = 1; i < G; i++) coldmax_g = fmax(coldmax_g, rec[par][i].cold);
        }
        const HcRec win = rec[par][wi];
        const int pcol = win.col, ppos = win.pos;
        if (j > 0 && !(win.val > coldmax_g)) {
            // a column that was not kept current in this block may beat the best hot candidate: close the block,
            // after which every column is current, and choose again among all of them
            close_block(j, k0);
            QC(13, 1)
            k0 = k;
            j = 0;
            coldmax_l = -1.0;
            Cand lb = block_best(best_of_all(k));
            QT(5)
            propose(lb, k, 0, -1.0);
            QT(6)
            csync();
            QT(7)
            continue;
        }
        if (pcol < 0) {  // no admissible column (NaN norms): stop here
            rank = k;
            flush = j > 0;
            break;
        }
        const double beta = win.beta, tau = win.tau;
        if (k == 0) r00 = fabs(beta);
        if (win.stop) {
            rank = k;
            flush = j > 0;
            break;
        }
        // ---- virtual swap of positions k and ppos ----
        for (int cl = tid; cl < ncl; cl += NT) {
            if (c_lo + cl == pcol) pos[cl] = k;
            else if (pos[cl] == k) pos[cl] = ppos;
        }
        // ---- a new block starts: hot = unpivoted columns within theta of the pivot's norm ----
        if (j == 0) {
            const double thr = theta2 * win.val;
            if (tid == 0) nhot_s = 0;
            __syncthreads();
            double cm = -1.0;
            for (int base = 0; base < ncl; base += NT) {
                const int cl = base + tid;
                bool h = false;
                if (cl < ncl) {
                    const bool act = pos[cl] > k;
                    const double n1 = nq1[cl];
                    h = act && n1 >= thr;
                    state[cl] = h ? 1 : 0;
                    if (act && !h) cm = fmax(cm, n1);
                }
                const unsigned m = __ballot_sync(FULL, h);
                if (m) {
                    int b0 = 0;
                    if (lane == 0) b0 = atomicAdd(&nhot_s, __popc(m));
                    b0 = __shfl_sync(FULL, b0, 0);
                    if (h) hot[b0 + __popc(m & ((1u << lane) - 1u))] = cl;
                }
            }
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) cm = fmax(cm, __shfl_xor_sync(FULL, cm, o));
            if (lane == 0) cmax_s[warp] = cm;
            __syncthreads();
            cm = cmax_s[0];
#pragma unroll
            for (int w = 1; w < NW; w++) cm = fmax(cm, cmax_s[w]);
            coldmax_l = cm;
            nhot = nhot_s;
            QC(14, nhot)
            QC(15, min(ncl, cols - k))
        }
        // ---- pull the winner's reflector into column j of the block (zero outside [k, rows)) ----
        double* vj = Vs + (size_t)j * ldv;
        {
            const double* wv = slots + (size_t)par * ldv;
            if constexpr (G > 1) wv = cluster.map_shared_rank(wv, wi);
            const bool own = (crank == wi);
            for (int i = tid; i < ldv; i += NT) {
                double v = (i >= k && i < rows) ? wv[i] : 0.0;
                vj[i] = v;
                if (own && i > k && i < rows) t.V[i + (size_t)k * rows] = v;
            }
            if (own && tid == 0) {
                P[k + (size_t)(pcol - c_lo) * ld] = beta;
                t.tau[k] = tau;
            }
        }
        __syncthreads();
        QT(1)  // winner selection + classification + reflector pull
        // ---- aux = V(k:, 0:j)^T v, then column j of T ----
        if (j > 0) {
            for (int tt = warp; tt < j; tt += NW) {
                const double* vt = Vs + (size_t)tt * ldv;
                double s = 0.0;
                for (int i = k + lane; i < rows; i += 32) s += vt[i] * vj[i];
                s = group_sum(s, 32);
                if (lane == 0) aux[tt] = s;
            }
            __syncthreads();
            if (tid < j) {
                double s = 0.0;
                for (int u = tid; u < j; u++) s = fma(Ts[tid * QNB + u], aux[u], s);
                Ts[tid * QNB + j] = -tau * s;
            }
        }
        if (tid == 0) Ts[j * QNB + j] = tau;
        QT(2)  // aux = V^T v
        // ---- hot columns: F(:, j), row k of the panel, partial norms, local candidate for step k + 1 ----
        Cand best{-1.0, INT_MAX, -1};
        const double2* v2 = reinterpret_cast<const double2*>(vj);
        const double2* cb = reinterpret_cast<const double2*>(P);
        for (int first = warp * WB; first < nhot; first += WB * NW) {
            const int myc = lane / LPC;
            int my_cl = 0, my_p = -1;
            const bool my_valid = first + myc < nhot;
            if (my_valid) {
                my_cl = hot[first + myc];
                my_p = pos[my_cl];
            }
            const bool my_act = my_valid && my_p > k;
            if (!my_act) my_cl = 0;
            if (__ballot_sync(FULL, my_act) == 0) continue;
            double s[WB];
            int coff[WB];
#pragma unroll
            for (int c = 0; c < WB; c++) {
                s[c] = 0.0;
                coff[c] = hot[min(first + c, nhot - 1)] * ld2;
            }
            for (int i2 = (k >> 1) + lane; i2 < npair; i2 += 32) {
                const double2 vv = v2[i2];
                double2 a0[WB];
#pragma unroll
                for (int c = 0; c < WB; c++) a0[c] = cb[coff[c] + i2];
#pragma unroll
                for (int c = 0; c < WB; c++) {
                    s[c] = fma(a0[c].x, vv.x, s[c]);
                    s[c] = fma(a0[c].y, vv.y, s[c]);
                }
            }
            int bit = 16;
#pragma unroll
            for (int h = WB / 2; h >= 1; h >>= 1, bit >>= 1) {
                const bool up = (lane & bit) != 0;
#pragma unroll
                for (int i = 0; i < h; i++) {
                    const double keep = up ? s[i + h] : s[i];
                    const double send = up ? s[i] : s[i + h];
                    s[i] = keep + __shfl_xor_sync(FULL, send, bit);
                }
            }
            double tot = s[0];
#pragma unroll
            for (int o = LPC / 2; o >= 1; o >>= 1) tot += __shfl_xor_sync(FULL, tot, o);
            double my_f = 0.0, my_ak = 0.0, my_n1 = 0.0, my_newn = 0.0;
            bool my_need = false;
            if (my_act) {
                const double* cj = P + (size_t)my_cl * ld;
                const double* fr = Fs + (size_t)my_cl * FLD;
                double corr = 0.0, rk = cj[k];
                for (int tt = 0; tt < j; tt++) {
                    const double ft = fr[tt];
                    corr = fma(ft, aux[tt], corr);
                    rk = fma(-Vs[k + (size_t)tt * ldv], ft, rk);
                }
                my_f = tau * (tot - corr);
                my_ak = rk - my_f;
                my_n1 = nq1[my_cl];
                if (my_n1 != 0.0) {
                    my_newn = fmax(0.0, my_n1 - my_ak * my_ak);
                    my_need = my_newn <= tol3z * nq2[my_cl];
                }
            }
            unsigned nm = __ballot_sync(FULL, my_need && (lane & (LPC - 1)) == 0);
            while (nm) {
                const int srcl = __ffs(nm) - 1;
                nm &= nm - 1;
                const int ccl = __shfl_sync(FULL, my_cl, srcl);
                const double ff = __shfl_sync(FULL, my_f, srcl);
                const double* cj = P + (size_t)ccl * ld;
                const double* fr = Fs + (size_t)ccl * FLD;
                double qq = 0.0;
                for (int i = k + 1 + lane; i < rows; i += 32) {
                    double u = cj[i];
                    for (int tt = 0; tt < j; tt++) u -= Vs[i + (size_t)tt * ldv] * fr[tt];
                    u -= vj[i] * ff;
                    qq += u * u;
                }
                qq = group_sum(qq, 32);
                if ((lane / LPC) == (srcl / LPC)) my_newn = qq;
            }
            if (my_act && (lane & (LPC - 1)) == 0) {
                Fs[(size_t)my_cl * FLD + j] = my_f;
                P[k + (size_t)my_cl * ld] = my_ak;
                if (my_n1 != 0.0) {
                    nq1[my_cl] = my_newn;
                    if (my_need) nq2[my_cl] = my_newn;
                }
                if (better(my_newn, my_p, best.val, best.pos)) best = Cand{my_newn, my_p, c_lo + my_cl};
            }
        }
        if (k + 1 >= mn) break;  // factorization complete, rank = mn (every column is a pivot, or rank == rows)
        QT(3)  // sweep (warp 0's share)
        QC(11, 1)
        j++;
        Cand lb;
        if (j == nbmax) {  // the block is full
            __syncthreads();
            QT(4)
            close_block(j, k0);
            QC(12, 1)
            k0 = k + 1;
            j = 0;
            coldmax_l = -1.0;
            lb = block_best(best_of_all(k + 1));
        } else {
            lb = block_best(best);  // contains a block barrier: F and row k are visible below
            QT(4)
        }
        QT(5)  // end of block
        propose(lb, k + 1, j, coldmax_l);
        QT(6)  // speculative reflector + record broadcast
        csync();
        QT(7)  // cluster barrier
        k++;
    }
    // All CTAs leave the loop at the same step. One more barrier so that no CTA exits (or starts scattering over
    // its panel) while a sibling may still be pulling from its slots.
    csync();
    if (rank >= rows) {  // nothing to do (tree.cpp:1317-1319); csize unchanged
        QT_FLUSH
        return;
    }
    // reflectors of the open block are still pending on the cold columns: their rows [k0, rank) are R entries
    if (flush) close_block(j, k0);

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
                        int i = e % rank, jj = e / rank;
                        int p = pos[a + jj - c_lo];
                        dp[i + (size_t)jj * q.ld] = (p >= rank || i <= p) ? P[i + (size_t)(a + jj - c_lo) * ld] : 0.0;
                    }
                } else {
                    double* dp = q.blk + off;
                    for (int e = tid; e < tot; e += NT) {
                        int jj = e % nc, i = e / nc;
                        int p = pos[a + jj - c_lo];
                        dp[jj + (size_t)i * q.ld] = (p >= rank || i <= p) ? P[i + (size_t)(a + jj - c_lo) * ld] : 0.0;
                    }
                }
            }
            c0 += w;
        }
    }
    if (crank == 0 && tid == 0) csize[t.cluster] = rank;
    QT(8)  // scatter
    QT_FLUSH
}

template <int G, int NT, bool SP, int QNB = QR_NB, int MINB = (NT <= 128 ? 5 : (NT <= 256 ? 2 : 1))>
void launch_one(const QrTask* t, int nt, const QrSrc* s, int* csize, double tol, int smem, cudaStream_t st) {
    auto kern = rrqr_blocked_kernel<G, NT, SP, QNB, MINB>;
    // function attributes are per device and cheap to set: no process-wide cache (several GPUs / threads may launch)
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (G > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
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
    const cudaError_t err = cudaLaunchKernelEx(&cfg, kern, t, s, csize, tol);
    if (err != cudaSuccess)  // a shape the device cannot schedule must not pass silently (the ranks would be garbage)
        throw std::runtime_error(std::string("rrqr launch failed (G=") + std::to_string(G) + ", threads=" +
                                 std::to_string(NT) + ", smem=" + std::to_string(smem) + "): " + cudaGetErrorString(err));
}

template <int G, int NT, int QNB, int MINB>
void launch_one_hc(const QrTask* t, int nt, const QrSrc* s, int* csize, double tol, double theta2, int smem,
                   cudaStream_t st) {
    auto kern = rrqr_hc_kernel<G, NT, QNB, MINB>;
    // attributes are per device and cheap to set: no process-wide cache (several GPUs / threads may launch)
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (G > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
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
    const cudaError_t err = cudaLaunchKernelEx(&cfg, kern, t, s, csize, tol, theta2);
    if (err != cudaSuccess)
        throw std::runtime_error(std::string("rrqr (hot/cold) launch failed (G=") + std::to_string(G) + ", threads=" +
                                 std::to_string(NT) + ", smem=" + std::to_string(smem) + "): " + cudaGetErrorString(err));
}

}  // namespace

int rrqr_max_smem() { return 224 * 1024; }

void rrqr_phase_cycles(unsigned long long* out48, bool reset) {
    cudaMemcpyFromSymbol(out48, g_qr_phase, sizeof(unsigned long long) * 48);
    if (reset) {
        unsigned long long z[48] = {};
        cudaMemcpyToSymbol(g_qr_phase, z, sizeof(z));
    }
}

size_t rrqr_smem_bytes(int rows, int maxcols, int G, int nb, int ld, bool in_smem) {
    size_t cpcm = (size_t)(maxcols + G - 1) / G;
    size_t cpce = (cpcm + 3) & ~(size_t)3;
    size_t ldv = ((size_t)rows + 1) & ~(size_t)1;
    size_t fld = (size_t)(nb | 1);
    size_t doubles = ldv * nb + ((cpcm * fld + 1) & ~(size_t)1) + 2 * cpce + 2 * ldv + (in_smem ? (size_t)ld * cpcm : 0);
    // global panels: pos + hot flag + hot list of the hot / cold kernel
    return doubles * sizeof(double) + (in_smem ? 1 : 3) * cpce * sizeof(int);
}

void launch_rrqr(const QrTask* t, int nt, const QrSrc* s, int* csize, double tol, int G, int nthreads, bool in_smem,
                 int smem, cudaStream_t st, double theta) {
    if (nt <= 0) return;
    if (!in_smem && theta > 0.0) {
        // hot / cold kernel: 256 threads (nb <= QR_NBS, 4 CTAs per SM) or 512 threads (nb <= QR_NB, 1 CTA per SM)
        const double th2 = theta * theta;
        if (nthreads >= 512) {
            switch (G) {
                case 1: launch_one_hc<1, 512, QR_NB, 1>(t, nt, s, csize, tol, th2, smem, st); break;
                case 2: launch_one_hc<2, 512, QR_NB, 1>(t, nt, s, csize, tol, th2, smem, st); break;
                case 4: launch_one_hc<4, 512, QR_NB, 1>(t, nt, s, csize, tol, th2, smem, st); break;
                case 8: launch_one_hc<8, 512, QR_NB, 1>(t, nt, s, csize, tol, th2, smem, st); break;
                default: launch_one_hc<16, 512, QR_NB, 1>(t, nt, s, csize, tol, th2, smem, st); break;
            }
        } else {
            switch (G) {
                case 1: launch_one_hc<1, 256, QR_NBS, SPAND_STREAM_MINB>(t, nt, s, csize, tol, th2, smem, st); break;
                case 2: launch_one_hc<2, 256, QR_NBS, SPAND_STREAM_MINB>(t, nt, s, csize, tol, th2, smem, st); break;
                case 4: launch_one_hc<4, 256, QR_NBS, SPAND_STREAM_MINB>(t, nt, s, csize, tol, th2, smem, st); break;
                case 8: launch_one_hc<8, 256, QR_NBS, SPAND_STREAM_MINB>(t, nt, s, csize, tol, th2, smem, st); break;
                default: launch_one_hc<16, 256, QR_NBS, SPAND_STREAM_MINB>(t, nt, s, csize, tol, th2, smem, st); break;
            }
        }
        return;
    }
    if (!in_smem) {
        if (nthreads >= 512) {
            if (G <= 8) launch_one<8, 512, false>(t, nt, s, csize, tol, smem, st);
            else launch_one<16, 512, false>(t, nt, s, csize, tol, smem, st);
            return;
        }
        // streaming shape: panel in global memory, small CTAs, several per SM
        switch (G) {
            case 1: launch_one<1, 256, false, QR_NBS, SPAND_STREAM_MINB>(t, nt, s, csize, tol, smem, st); break;
            case 2: launch_one<2, 256, false, QR_NBS, SPAND_STREAM_MINB>(t, nt, s, csize, tol, smem, st); break;
            case 4: launch_one<4, 256, false, QR_NBS, SPAND_STREAM_MINB>(t, nt, s, csize, tol, smem, st); break;
            case 8: launch_one<8, 256, false, QR_NBS, SPAND_STREAM_MINB>(t, nt, s, csize, tol, smem, st); break;
            default: launch_one<16, 256, false, QR_NBS, SPAND_STREAM_MINB>(t, nt, s, csize, tol, smem, st); break;
        }
        return;
    }
    if (nthreads <= 128) {
        launch_one<1, 128, true>(t, nt, s, csize, tol, smem, st);
    } else if (nthreads <= 256) {
        switch (G) {
            case 1: launch_one<1, 256, true>(t, nt, s, csize, tol, smem, st); break;
            case 2: launch_one<2, 256, true>(t, nt, s, csize, tol, smem, st); break;
            case 4: launch_one<4, 256, true>(t, nt, s, csize, tol, smem, st); break;
            case 8: launch_one<8, 256, true>(t, nt, s, csize, tol, smem, st); break;
            default: launch_one<16, 256, true>(t, nt, s, csize, tol, smem, st); break;
        }
    } else {
        switch (G) {
            case 1: launch_one<1, 512, true>(t, nt, s, csize, tol, smem, st); break;
            case 2: launch_one<2, 512, true>(t, nt, s, csize, tol, smem, st); break;
            case 4: launch_one<4, 512, true>(t, nt, s, csize, tol, smem, st); break;
            case 8: launch_one<8, 512, true>(t, nt, s, csize, tol, smem, st); break;
            default: launch_one<16, 512, true>(t, nt, s, csize, tol, smem, st); break;
        }
    }
}

// Stand-alone truncated QRCP of one dense matrix through the batch kernels (geqp3 + choose_rank + triu(R) P^T of
// src/util.cpp:383-452 and src/tree.cpp:1334-1335 on a single matrix). The matrix is given as `nsrc` column blocks
// (sources) of equal width, optionally stored transposed, so that the gather path is exercised as well. Used by the
// kernel-level parity tests (tests/test_gpu_rrqr.py) to drive every launch shape on matrices of any size.
int rrqr_single(int rows, int cols, const double* A_host, int nsrc, int transposed, double tol, int G, int nthreads,
                int in_smem, int nb, double theta, int* rank_out, double* R_host, double* V_host, double* tau_host,
                std::string& err) {
    auto fail = [&](const char* what, cudaError_t e) {
        err = std::string(what) + ": " + cudaGetErrorString(e);
        return -1;
    };
    if (rows <= 0 || cols <= 0 || nsrc <= 0 || cols % nsrc != 0) {
        err = "rrqr_single: bad shape";
        return -1;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        err = "no CUDA device available (the spaND B200 path has no CPU fallback)";
        return -1;
    }
    const int w = cols / nsrc;
    QrTask t{};
    t.cluster = 0;
    t.rows = rows;
    t.src0 = 0;
    t.nsrc = nsrc;
    t.maxcols = cols;
    t.nb = std::max(1, std::min(nb, std::min(rows, cols)));
    t.in_smem = in_smem;
    auto pow2floor = [](int x) { int p = 1; while (2 * p <= x) p *= 2; return p; };
    auto pow2ceil = [](int x) { int p = 1; while (p < x) p *= 2; return p; };
    const int cpcm = std::max(1, (cols + G - 1) / G);
    int Ln = in_smem ? std::min(32, std::max(1, pow2floor(nthreads / cpcm))) : 32;
    Ln = std::min(Ln, pow2ceil(std::max(1, (rows + 1) / 2)));
    int ld = (rows + 1) & ~1;
    if (in_smem && Ln < 8)
        while (ld % 16 != (2 * Ln) % 16) ld += 2;
    t.L = Ln;
    t.ld = ld;
    const bool hc2 = in_smem == 2;  // hot-set kernel: panel in global memory, `nb` = capacity of the hot set
    size_t smem = rrqr_smem_bytes(rows, cols, G, t.nb, ld, in_smem != 0);
    if (hc2) {
        if (hc2_row_pairs(rows) == 0) {
            err = "rrqr_single: too many rows for the hot-set kernel";
            return -1;
        }
        t.in_smem = 0;
        t.hcap = std::max(1, std::min(nb, cols));
        t.nb = HC2_NB;
        t.L = 32;
        t.ld = ld = (rows + 1) & ~1;
        smem = hc2_smem_bytes(rows, cols, G, t.hcap, nsrc);
    }
    const bool colk = in_smem == 3 || in_smem == 4;  // column kernel; 4: panel in global scratch instead of shared memory
    const bool colgp = in_smem == 4;
    if (colk) {
        if (rows > rrqr_col_max_rows() || G != 1) {
            err = "rrqr_single: the column kernel takes panels of at most 128 rows on one CTA";
            return -1;
        }
        t.in_smem = colgp ? 0 : 1;
        t.L = 1;
        t.nb = 1;
        t.ld = ld = rrqr_col_ld(rows);
        smem = rrqr_col_smem_bytes(rows, cols, nsrc, colgp);
    }
    if (smem > (size_t)rrqr_max_smem()) {
        err = "rrqr_single: shape does not fit the shared memory of one CTA";
        return -1;
    }
    const int mn = std::min(rows, cols);
    double *dA = nullptr, *dW = nullptr, *dV = nullptr, *dtau = nullptr, *dX = nullptr;
    int* dcs = nullptr;
    QrTask* dt = nullptr;
    QrSrc* ds = nullptr;
    cudaError_t e;
    const size_t abytes = sizeof(double) * (size_t)rows * cols;
    if ((e = cudaMalloc(&dA, abytes)) != cudaSuccess) return fail("cudaMalloc", e);
    cudaMalloc(&dW, sizeof(double) * (size_t)ld * cols);
    cudaMalloc(&dV, sizeof(double) * (size_t)rows * mn);
    cudaMalloc(&dtau, sizeof(double) * mn);
    cudaMalloc(&dcs, sizeof(int) * (nsrc + 1));
    cudaMalloc(&dt, sizeof(QrTask));
    cudaMalloc(&ds, sizeof(QrSrc) * nsrc);
    cudaMemcpy(dA, A_host, abytes, cudaMemcpyHostToDevice);
    cudaMemset(dV, 0, sizeof(double) * (size_t)rows * mn);
    cudaMemset(dtau, 0, sizeof(double) * mn);
    std::vector<int> cs(nsrc + 1, w);
    cs[0] = rows;
    cudaMemcpy(dcs, cs.data(), sizeof(int) * (nsrc + 1), cudaMemcpyHostToDevice);
    std::vector<QrSrc> srcs(nsrc);
    for (int i = 0; i < nsrc; i++) {
        // block i: rows x w at column i w (ld = rows), or its transpose w x rows stored with ld = w
        srcs[i].blk = dA + (size_t)i * w * rows;
        srcs[i].ld = transposed ? w : rows;
        srcs[i].nbr = 1 + i;
        srcs[i].transposed = transposed;
    }
    cudaMemcpy(ds, srcs.data(), sizeof(QrSrc) * nsrc, cudaMemcpyHostToDevice);
    if (hc2 && G >= 8) {
        cudaMalloc(&dX, sizeof(double) * hc2_exchange_doubles(rows, G));
        cudaMemset(dX, 0, sizeof(double) * hc2_exchange_doubles(rows, G));
    }
    t.X = dX;
    t.W = dW;
    t.V = dV;
    t.tau = dtau;
    cudaMemcpy(dt, &t, sizeof(QrTask), cudaMemcpyHostToDevice);
    if (const char* cps = getenv("SPAND_QR_COPIES")) {
        // micro-benchmark hook (scripts/qr_bench.py): `copies` independent replicas of the task in ONE launch, as a
        // wavefront of that many clusters would be; a warm-up launch on a first set of replicas, then the timed one
        const int copies = std::max(1, atoi(cps));
        const size_t wd = (size_t)ld * cols, vd = (size_t)rows * mn, xd = hc2 ? hc2_exchange_doubles(rows, G) : 0;
        double *bA, *bW, *bV, *bT, *bX = nullptr;
        int* bcs;
        QrTask* bt;
        QrSrc* bs;
        const int tot = 2 * copies;
        cudaMalloc(&bA, abytes * tot);
        cudaMalloc(&bW, sizeof(double) * wd * tot);
        cudaMalloc(&bV, sizeof(double) * vd * tot);
        cudaMalloc(&bT, sizeof(double) * mn * tot);
        if (xd) cudaMalloc(&bX, sizeof(double) * xd * tot);
        cudaMalloc(&bcs, sizeof(int) * (nsrc + 1) * tot);
        cudaMalloc(&bt, sizeof(QrTask) * tot);
        cudaMalloc(&bs, sizeof(QrSrc) * nsrc * tot);
        std::vector<QrTask> ht(tot, t);
        std::vector<QrSrc> hs((size_t)nsrc * tot);
        std::vector<int> hcs((size_t)(nsrc + 1) * tot);
        for (int c = 0; c < tot; c++) {
            cudaMemcpy(bA + (size_t)c * rows * cols, dA, abytes, cudaMemcpyDeviceToDevice);
            ht[c].cluster = c * (nsrc + 1);
            ht[c].src0 = c * nsrc;
            ht[c].W = bW + wd * c;
            ht[c].V = bV + vd * c;
            ht[c].tau = bT + (size_t)mn * c;
            ht[c].X = xd ? bX + xd * c : nullptr;
            hcs[(size_t)c * (nsrc + 1)] = rows;
            for (int i = 0; i < nsrc; i++) {
                hs[(size_t)c * nsrc + i] = srcs[i];
                hs[(size_t)c * nsrc + i].blk = bA + (size_t)c * rows * cols + (size_t)i * w * rows;
                hs[(size_t)c * nsrc + i].nbr = c * (nsrc + 1) + 1 + i;
                hcs[(size_t)c * (nsrc + 1) + 1 + i] = w;
            }
        }
        cudaMemcpy(bt, ht.data(), sizeof(QrTask) * tot, cudaMemcpyHostToDevice);
        cudaMemcpy(bs, hs.data(), sizeof(QrSrc) * nsrc * tot, cudaMemcpyHostToDevice);
        cudaMemcpy(bcs, hcs.data(), sizeof(int) * (nsrc + 1) * tot, cudaMemcpyHostToDevice);
        if (xd) cudaMemset(bX, 0, sizeof(double) * xd * tot);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        const int sm = (int)((smem + 1023) & ~(size_t)1023);
        float ms = -1.f;
        try {
            for (int rep = 0; rep < 2; rep++) {
                if (rep == 1) cudaEventRecord(e0, 0);
                if (colk) launch_rrqr_col(bt + rep * copies, copies, bs, bcs, tol, nthreads, sm, colgp, 0);
                else if (hc2) launch_rrqr_hc2(bt + rep * copies, copies, bs, bcs, tol, G, hc2_row_pairs(rows), sm, theta, 0);
                else launch_rrqr(bt + rep * copies, copies, bs, bcs, tol, G, nthreads, in_smem != 0, sm, 0, theta);
            }
            cudaEventRecord(e1, 0);
            const cudaError_t ee = cudaDeviceSynchronize();
            if (ee != cudaSuccess) fprintf(stderr, "QRBENCH kernel error: %s\n", cudaGetErrorString(ee));
            cudaEventElapsedTime(&ms, e0, e1);
        } catch (std::exception& ex) {
            fprintf(stderr, "QRBENCH launch error: %s\n", ex.what());
        }
        int r0 = -1;
        cudaMemcpy(&r0, bcs + (size_t)copies * (nsrc + 1), sizeof(int), cudaMemcpyDeviceToHost);
        fprintf(stderr, "QRBENCH rows %d cols %d nsrc %d G %d copies %d mode %d nb/hot %d theta %.3f smem %d rank %d: %.3f ms\n",
                rows, cols, nsrc, G, copies, in_smem, nb, theta, sm, r0, ms);
        cudaFree(bA); cudaFree(bW); cudaFree(bV); cudaFree(bT); cudaFree(bX); cudaFree(bcs); cudaFree(bt); cudaFree(bs);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    int rc = 0;
    try {
        if (colk)
            launch_rrqr_col(dt, 1, ds, dcs, tol, nthreads, (int)((smem + 1023) & ~(size_t)1023), colgp, 0);
        else if (hc2)
            launch_rrqr_hc2(dt, 1, ds, dcs, tol, G, hc2_row_pairs(rows), (int)((smem + 1023) & ~(size_t)1023), theta, 0);
        else
            launch_rrqr(dt, 1, ds, dcs, tol, G, nthreads, in_smem != 0, (int)((smem + 1023) & ~(size_t)1023), 0, theta);
    } catch (std::exception& ex) {
        err = ex.what();
        rc = -1;
    }
    if (rc == 0 && (e = cudaDeviceSynchronize()) != cudaSuccess) rc = fail("rrqr_single kernel", e);
    if (rc == 0) {
        cudaMemcpy(cs.data(), dcs, sizeof(int) * (nsrc + 1), cudaMemcpyDeviceToHost);
        *rank_out = cs[0];
        cudaMemcpy(R_host, dA, abytes, cudaMemcpyDeviceToHost);
        cudaMemcpy(V_host, dV, sizeof(double) * (size_t)rows * mn, cudaMemcpyDeviceToHost);
        cudaMemcpy(tau_host, dtau, sizeof(double) * mn, cudaMemcpyDeviceToHost);
    }
    cudaFree(dA);
    cudaFree(dW);
    cudaFree(dX);
    cudaFree(dV);
    cudaFree(dtau);
    cudaFree(dcs);
    cudaFree(dt);
    cudaFree(ds);
    return rc;
}

}  // namespace spand
