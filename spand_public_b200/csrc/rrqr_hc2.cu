// Truncated column-pivoted QR with the pivot search confined to a shared-memory "hot set" (sm_100a).
//
// Same contract as rrqr_blocked_kernel (rrqr.cu): for every interface cluster of a wavefront, assemble_Asn
// (src/tree.cpp:1189-1224) -> dgeqp3 (src/util.cpp:383-394) -> choose_rank (src/util.cpp:434-452) ->
// triu(R[:rank, :]) P^T scattered back in place (src/tree.cpp:1334-1335, :1004-1046), (V, tau) kept for the
// Orthogonal op (src/tree.cpp:1322-1331). Citations relative to /root/reference.
//
// Why another kernel. A greedy pivot search only needs the columns that can win it. Partial column norms never grow,
// so the norm a column had when it was last brought up to date bounds it from above. Work proceeds in blocks:
//
//   block start  every unpivoted column of the panel (in global memory, column slabs owned by the G CTAs of a
//                thread-block cluster) is current and carries its exact squared norm. Every CTA proposes its best
//                column; the winner is the next pivot and sets the level. Each CTA then copies the columns of ITS slab
//                that lie within theta of the level - the largest first, as many as fit - into its shared memory:
//                the hot set is distributed over the cluster, every CTA keeps `coldmax`, its largest norm outside it.
//   hot loop     unblocked Householder QRCP on the hot columns (dlaqp2 arithmetic: reflector applied at once, the
//                dlaqps / dlaqp2 norm downdate with its recomputation safeguard, first-index tie-breaking through
//                virtual positions). Per step: every CTA builds the reflector of its own best candidate
//                speculatively (one warp, no block barrier) and sends a 48-byte record to its siblings through
//                distributed shared memory; ONE cluster barrier; all pick the same winner and pull its reflector
//                from the winner's shared memory; one warp per hot column applies it. The block stays open while the
//                winner is strictly larger than every CTA's coldmax (then it is the column the full search would
//                pick), up to NB steps.
//   block end    T of the compact-WY form of the block's reflectors from their Gram matrix; hot columns go back to
//                the panel; every cold column of the CTA's slab gets F = W^T V T and W -= V F^T on the FP64 tensor
//                cores (mma.sync m8n8k4) in one read + one write, and its exact new norm from the registers of that
//                update; positions are replayed from the swap log.
//
// Measured on config C4 (profiles/r2_rrqr.md): the hot set is 2-5 % of the unpivoted columns; a cold column costs one
// read + one write per block instead of one read per step.
#include <cooperative_groups.h>

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace spand {

__device__ unsigned long long g_hc2_stat[16];  // -DSPAND_RRQR_TIMING: [0] steps [1] blocks [2] early closes
                                               // [3] hot columns [4] unpivoted columns [5] threshold retries
                                               // [6..9] cycles: boundary / hot loop / block end / setup+scatter
#ifdef SPAND_RRQR_TIMING
#define HC(i, v) \
    if (threadIdx.x == 0 && crank == 0) atomicAdd(&g_hc2_stat[i], (unsigned long long)(v));
#define HT_DECL long long ht_t = clock64();
#define HT(i)                                                        \
    if (threadIdx.x == 0 && crank == 0) {                            \
        const long long ht_n = clock64();                            \
        atomicAdd(&g_hc2_stat[i], (unsigned long long)(ht_n - ht_t)); \
        ht_t = ht_n;                                                 \
    }
#else
#define HC(i, v)
#define HT_DECL
#define HT(i)
#endif

namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

struct Cand {
    double val;
    int pos, col;  // col: hot slot inside the hot loop, global column index at block boundaries
};

__device__ __forceinline__ void dmma_f64(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ bool better(double v, int p, double bv, int bp) { return v > bv || (v == bv && p < bp); }

__device__ __forceinline__ Cand warp_best(Cand b) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(FULL, b.val, o);
        int op = __shfl_xor_sync(FULL, b.pos, o);
        int oc = __shfl_xor_sync(FULL, b.col, o);
        if (better(ov, op, b.val, b.pos)) b = Cand{ov, op, oc};
    }
    return b;
}

// Exchange area of one task in global memory (G > 1), per CTA: 8 doubles of header + 2 doubles per local column.
//   hdr[0] best squared norm, hdr[1] = (pos, col) of it, hdr[2] = (count, -) , hdr[3] = local coldmax,
//   hdr[8 .. 8 + HC2_BINS / 2) histogram (two bins per double); then entry e: [n1, (col, pos)]
constexpr int HC2_BINS = 64;
constexpr int COL_MAXROWS = 128;  // tallest panel of the column kernel
// Doubles of the hot-column area of a CTA: the hot set itself, or (block end) the scratch of the cold refresh
// followed by at least two TMA stages of one strip (8 columns) each.
__host__ __device__ inline size_t hc2_scratch_doubles(int nw) {
    return ((size_t)nw * (8 * (HC2_NB + 1) + HC2_NB * 8) + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t hc2_hot_doubles(size_t ldv, int hcap, int nw) {
    const size_t a = ldv * (size_t)hcap, b = hc2_scratch_doubles(nw) + 16 * ldv;
    return a > b ? a : b;
}
__host__ __device__ inline size_t hc2_xstride(int cpce) { return 8 + HC2_BINS / 2 + 2 * (size_t)cpce; }

__device__ __forceinline__ double pack2(int a, int b) {
    return __hiloint2double(b, a);
}
__device__ __forceinline__ void unpack2(double d, int& a, int& b) {
    a = __double2loint(d);
    b = __double2hiint(d);
}

// ---- TMA (1-D bulk copy) + mbarrier, raw PTX ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy (bytes: multiple of 16, both addresses 16-byte aligned), completion on `bar`
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_global, unsigned bytes,
                                            unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_global), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void team_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Warp-cooperative copy of an nf x ns tile with lanes along the fast index f: element (f, s) is produced by load(f, s)
// and consumed by store(f, s, value). Eight loads are issued per lane before the first store, so that the copy runs
// at the memory-level parallelism of the loads (a plain load/store loop is serialised by possible aliasing).
template <class LoadF, class StoreF>
__device__ __forceinline__ void warp_tile_copy(int nf, int ns, int lane, LoadF load, StoreF store) {
    const int tot = nf * ns;
    const int dq = 32 / nf, dr = 32 % nf;
    int f = lane % nf, sidx = lane / nf;
    for (int e = lane; e < tot; e += 32 * 8) {
        double tmp[8];
        int ff[8], ssx[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            ff[u] = f;
            ssx[u] = sidx;
            tmp[u] = (e + 32 * u < tot) ? load(f, sidx) : 0.0;
            f += dr;
            sidx += dq;
            if (f >= nf) {
                f -= nf;
                sidx++;
            }
        }
#pragma unroll
        for (int u = 0; u < 8; u++)
            if (e + 32 * u < tot) store(ff[u], ssx[u], tmp[u]);
    }
}

struct HcRec {      // what a CTA tells the cluster about its candidate for the next step
    double val;     // squared partial norm of the candidate, < 0: none
    double beta, tau;
    double cold;    // largest squared norm among the CTA's cold columns (bound during the block), < 0: none
    int pos, col;   // virtual position and global column index
    int stop, pad;
};

// G: CTAs per task (cluster), NT threads, NB: largest block, RP: row pairs per lane (rows <= 64 RP)
template <int G, int NT, int NB, int RP, int MINB>
__global__ void __launch_bounds__(NT, MINB) rrqr_hc2_kernel(const QrTask* __restrict__ tasks,
                                                            const QrSrc* __restrict__ srcs, int* csize, double tol,
                                                            double theta2, int staged_refresh) {
    constexpr int NW = NT / 32;
    constexpr int FLD = NB + 1;
    constexpr int MT = NB / 8;
    const int task_id = blockIdx.x / G;
    const int crank = blockIdx.x % G;
    const QrTask t = tasks[task_id];
    const QrSrc* src = srcs + t.src0;
    const int rows = t.rows;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    HT_DECL

    __shared__ double Ts[NB * NB];      // compact-WY factor of the block (upper triangular, row-major)
    __shared__ double Gm[NB * NB];      // Gram matrix of the block's reflectors (upper triangle)
    __shared__ double tau_s[NB], beta_s[NB];
    __shared__ int lcol[NB], lpos[NB];  // swap log of the block: pivot column and its position before the swap
    __shared__ Cand wbest[2][NW];
    __shared__ double wmax[NW];
    __shared__ HcRec rec[2][G];         // proposals of the cluster, double buffered on the proposal parity
    __shared__ int hist[HC2_BINS];      // own unpivoted columns by log2(level / norm^2), quarter-octave bins
    __shared__ int nh_s;
    __shared__ __align__(8) unsigned long long mbar[16];  // "stage full" barriers of the cold refresh: [team][buffer]
    extern __shared__ __align__(128) double dsm[];
    if (threadIdx.x < 16) mbar_init(&mbar[threadIdx.x], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    unsigned tma_uses = 0;  // strips this warp's team has consumed so far (buffer and phase parity of the pipeline)

    // column offset of every source block inside the panel: widths read in parallel, prefix sum by the first warp
    // (an interface cluster of the upper levels has hundreds of neighbours: no serial chain of global loads)
    int* soff;
    {
        const int cpcm0 = (t.maxcols + G - 1) / G, cpce0 = (cpcm0 + 3) & ~3, hce0 = (t.hcap + 3) & ~3;
        const size_t ldv0 = (size_t)((rows + 1) & ~1);
        const size_t hot0 = hc2_hot_doubles(ldv0, t.hcap, NW);
        soff = (int*)(dsm + ldv0 * NB + hot0 + 2 * (size_t)hce0 + cpce0) + 2 * cpce0 + 2 * hce0;
    }
    for (int s = tid; s < t.nsrc; s += NT) soff[s + 1] = csize[src[s].nbr];
    if (tid == 0) soff[0] = 0;
    __syncthreads();
    if (warp == 0) {
        const int per = (t.nsrc + 31) / 32;
        const int lo = 1 + lane * per, hi = min(t.nsrc + 1, lo + per);
        int sum = 0;
        for (int i = lo; i < hi; i++) sum += soff[i];
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
        }
        int run = incl - sum;
        for (int i = lo; i < hi; i++) {
            run += soff[i];
            soff[i] = run;
        }
    }
    __syncthreads();
    const int cols = soff[t.nsrc];
    if (rows == 0) return;
    if (tol >= 1.0 || cols == 0) {  // choose_rank: tol >= 1 -> 0 ; no neighbours -> rank 0
        if (crank == 0 && tid == 0) csize[t.cluster] = 0;
        return;
    }
    const int mn = min(rows, cols);
    const int cpc = (cols + G - 1) / G;
    const int cpcm = (t.maxcols + G - 1) / G;
    const int cpce = (cpcm + 3) & ~3;
    const int c_lo = min(cols, crank * cpc), c_hi = min(cols, c_lo + cpc);
    const int ncl = c_hi - c_lo;
    const int ld = t.ld, ldv = (rows + 1) & ~1;  // ld == ldv for global panels
    const int npair = ldv >> 1;
    const int HCAP = t.hcap, HCAPE = (HCAP + 3) & ~3;

    // H is dead between the write-back of the hot columns and the next block start: the per-warp scratch of the
    // cold refresh (F rows and Y = V^T W of one strip of 8 columns) lives in the same bytes
    const size_t hbytes_d = hc2_hot_doubles((size_t)ldv, HCAP, NW);
    double* Vs = dsm;                              // ldv x NB  reflectors of the open block (zero above the diagonal)
    double* H = Vs + (size_t)ldv * NB;             // ldv x HCAP  own hot columns, always current
    double* ftmp = H;                              // per warp: F rows of one strip of 8 cold columns
    double* ysm = ftmp + (size_t)NW * 8 * FLD;     // per warp: Y = V^T W of that strip
    double* hn1 = H + hbytes_d;                    // squared partial norms of the hot columns
    double* hn2 = hn1 + HCAPE;                     // ... at their last exact computation
    double* nq1 = hn2 + HCAPE;                     // squared norms of the local columns (exact at block boundaries)
    int* pos = (int*)(nq1 + cpce);                 // virtual position of every local column
    int* state = pos + cpce;                       // 1: local column is in the hot set of the open block
    int* hcol = state + cpce;                      // local column index of every hot slot
    int* hpos = hcol + HCAPE;                      // its virtual position (kept live inside the block)
    double* slots = (double*)(soff + ((t.nsrc + 2) & ~1));  // 2 x ldv: own candidate reflectors (double buffered)
    // Wide clusters: 16 CTAs pulling one reflector out of the winner's shared memory are serialised by its port
    // (about 20 bytes per cycle); they exchange the candidates through L2 instead (t.X: G x 2 x ldv doubles)
    constexpr bool XG = G >= 8;
    double* P = t.W + (size_t)c_lo * ld;           // local slab, column cl at P[cl * ld]

    cg::cluster_group cluster = cg::this_cluster();
    auto csync = [&]() {
        if constexpr (G > 1) cluster.sync();
        else __syncthreads();
    };
    const double tol3z = sqrt(DBL_EPSILON);
    const int bins_theta = max(1, min(HC2_BINS - 1, (int)(-4.0f * __log2f((float)theta2))));
    double r00 = 0.0;
    int seq = 0;    // proposals made so far
    int bseq = 0;   // block-wide candidate reductions so far
    int my_slot = -1;  // hot slot of the own last proposal (-1: proposed from the panel)
    auto block_best = [&](Cand b) {
        const int par = (bseq++) & 1;
        b = warp_best(b);
        if (lane == 0) wbest[par][warp] = b;
        __syncthreads();
        Cand bb = (lane < NW) ? wbest[par][lane] : Cand{-1.0, INT_MAX, -1};
        return warp_best(bb);  // xor butterfly: every lane holds the result
    };
    // best unpivoted local column (positions >= kk) by its squared norm; every column must be current
    auto best_of_slab = [&](int kk) {
        Cand best{-1.0, INT_MAX, -1};
        for (int cl = tid; cl < ncl; cl += NT) {
            const int p = pos[cl];
            if (p >= kk) {
                const double v = nq1[cl];
                if (better(v, p, best.val, best.pos)) best = Cand{v, p, cl};
            }
        }
        return best;
    };
    // Speculative reflector of the local candidate for step kn (first warp only: no block barrier), written to the own
    // slot; its record goes to every CTA of the cluster. lb.col: hot slot (from_hot) or local column of the panel.
    auto propose = [&](Cand lb, int kn, bool from_hot, double coldmax) {
        const int buf = (seq++) & 1;
        my_slot = from_hot ? lb.col : -1;
        if (warp == 0) {
            double* mine = XG ? t.X + ((size_t)crank * 2 + buf) * ldv : slots + (size_t)buf * ldv;
            HcRec r;
            r.val = -1.0;
            r.beta = r.tau = 0.0;
            r.cold = coldmax;
            r.pos = INT_MAX;
            r.col = -1;
            r.stop = 0;
            r.pad = 0;
            if (lb.col >= 0) {
                const double* pc = from_hot ? H + (size_t)lb.col * ldv : P + (size_t)lb.col * ld;
                double ss = 0.0;
                for (int i = kn + 1 + lane; i < rows; i += 32) ss = fma(pc[i], pc[i], ss);
                ss = warp_sum(ss);
                const double alpha = pc[kn];
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
                r.col = c_lo + (from_hot ? hcol[lb.col] : lb.col);
                r.beta = beta;
                r.tau = tau;
                r.stop = (tol != 0.0 && !(fabs(beta) / ref >= tol)) ? 1 : 0;
                for (int i = kn + lane; i < rows; i += 32) mine[i] = (i == kn) ? 1.0 : pc[i] * scal;
            }
            if constexpr (G > 1) {
                if (lane < G) cluster.map_shared_rank(&rec[0][0], lane)[buf * G + crank] = r;
            } else {
                if (lane == 0) rec[buf][0] = r;
            }
        }
    };
    int nh = 0;  // own hot slots of the open block
    // Closes the open block of jc reflectors that started at step k0. Ends with a block barrier.
    auto close_block = [&](int jc, int k0) {
        __syncthreads();
        // ---- Gram matrix of the reflectors, then T: T(0:jj, jj) = -tau_jj T(0:jj, 0:jj) G(0:jj, jj) ----
        for (int pr = warp; pr < jc * jc; pr += NW) {
            const int a = pr / jc, b = pr % jc;
            if (a > b) continue;
            const double* va = Vs + (size_t)a * ldv;
            const double* vb = Vs + (size_t)b * ldv;
            double s = 0.0;
            for (int i = k0 + b + lane; i < rows; i += 32) s = fma(va[i], vb[i], s);
            s = warp_sum(s);
            if (lane == 0) Gm[a * NB + b] = s;
        }
        __syncthreads();
        if (warp == 0) {
            for (int jj = 0; jj < jc; jj++) {
                if (lane < jj) {
                    double s = 0.0;
                    for (int u = lane; u < jj; u++) s = fma(Ts[lane * NB + u], Gm[u * NB + jj], s);
                    Ts[lane * NB + jj] = -tau_s[jj] * s;
                }
                if (lane == jj) Ts[jj * NB + jj] = tau_s[jj];
                __syncwarp();
            }
        }
        // ---- Householder vectors and tau of the block to the factor storage (first CTA) ----
        if (crank == 0) {
            for (int e = tid; e < jc * rows; e += NT) {
                const int jj = e / rows, i = e % rows;
                if (i > k0 + jj) t.V[i + (size_t)(k0 + jj) * rows] = Vs[i + (size_t)jj * ldv];
            }
            if (tid < jc) t.tau[k0 + tid] = tau_s[tid];
        }
        // ---- own hot columns back to the panel (rows >= k0), their norms, R diagonal of the pivots ----
        for (int s = warp; s < nh; s += NW) {
            const int col = c_lo + hcol[s];
            const double2* sp = reinterpret_cast<const double2*>(H + (size_t)s * ldv);
            double2* dp = reinterpret_cast<double2*>(P + (size_t)(col - c_lo) * ld);
            for (int pi = (k0 >> 1) + lane; pi < npair; pi += 32) dp[pi] = sp[pi];
            if (lane == 0) nq1[col - c_lo] = hn1[s];
        }
        // ---- positions: replay the swap log on every local column ----
        for (int cl = tid; cl < ncl; cl += NT) {
            int p = pos[cl];
            const int col = c_lo + cl;
            for (int jj = 0; jj < jc; jj++) {
                if (lcol[jj] == col) p = k0 + jj;
                else if (p == k0 + jj) p = lpos[jj];
            }
            pos[cl] = p;
        }
        __syncthreads();
        for (int jj = tid; jj < jc; jj += NT) {
            const int col = lcol[jj];
            if (col >= c_lo && col < c_hi) P[k0 + jj + (size_t)(col - c_lo) * ld] = beta_s[jj];
        }
        // ---- cold columns of the own slab: F = W^T V T, W -= V F^T, exact norms (FP64 tensor cores) ----
        // Staged path: a strip of 8 columns is ONE contiguous piece of the slab; a TMA bulk copy brings it into a
        // stage of the (now dead) hot-column area while the team works on the previous strip. A team of TW warps
        // splits the rows of a strip: partial Y = V^T W per warp, summed in a fixed order, F = Y^T T, then every warp
        // updates its rows from shared memory and stores them; the strip is read from L2 / HBM once.
        const size_t strip_d = (size_t)8 * ld;         // doubles per stage
        const size_t scr_d = hc2_scratch_doubles(NW);  // scratch in front of the stages
        const int nst = hbytes_d > scr_d ? (int)((hbytes_d - scr_d) / strip_d) : 0;
        if (nst >= 2 && staged_refresh) {
            // teams of warps with a private ring of D stages each: D - 1 strips are in flight ahead of the one being
            // worked on (the bytes in flight are what hides the L2 / HBM latency)
            const int T = nst >= 12 ? 4 : (nst >= 6 ? 2 : 1);
            const int D = min(4, nst / T);
            const int TW = NW / T;
            const int team = warp / TW, tw = warp % TW;
            const int kend = k0 + jc;
            const int g = lane >> 2, q = lane & 3;
            double* stage = H + scr_d + (size_t)team * D * strip_d;
            double* ypart = H + (size_t)team * TW * (NB * 8);          // [TW][NB * 8] partial Y tiles of the team
            double* ysum = H + (size_t)NW * (NB * 8) + team * (NB * 8);  // [NB * 8] their sum
            double* npart = H + (size_t)(NW + 4) * (NB * 8) + team * TW * 8;  // [TW][8] partial squared norms
            unsigned long long* full = &mbar[team * 4];
            const bool leader = tw == 0 && lane == 0;
            fence_proxy_async();  // generic writes (hot columns, panel) before the async-proxy copies
            __syncthreads();
            auto cold_mask = [&](int first) {
                unsigned m = 0;
#pragma unroll
                for (int c = 0; c < 8; c++)
                    if (first + c < ncl && pos[first + c] >= kend && !state[first + c]) m |= 1u << c;
                return m;
            };
            auto next_cold = [&](int first, unsigned& m) {
                m = 0;
                for (; first < ncl; first += 8 * T) {
                    m = cold_mask(first);
                    if (m) return first;
                }
                return -1;
            };
            auto issue = [&](int first, int buf) {
                const unsigned bytes = (unsigned)(min(8, ncl - first) * ld) * 8u;
                mbar_expect_tx(&full[buf], bytes);
                tma_load_1d(stage + (size_t)buf * strip_d, P + (size_t)first * ld, bytes, &full[buf]);
            };
            auto finish_norms = [&](int first, unsigned m) {  // one warp: partial norms of a strip in warp order
                if (lane < 8 && ((m >> lane) & 1u)) {
                    double acc = npart[lane];
                    for (int w = 1; w < TW; w++) acc += npart[w * 8 + lane];
                    nq1[first + lane] = acc;
                }
            };
            unsigned cm = 0, cmp = 0;
            int prev = -1, cur = -1;
            int issued = 0, consumed = 0, scan = team * 8;
            int fq[4];
            unsigned mq[4];
            const unsigned tma_base = tma_uses;
            auto pump = [&]() {  // keep the ring full: strips consumed .. issued - 1 own a stage each
                while (issued - consumed < D && scan < ncl) {
                    unsigned m;
                    const int f = next_cold(scan, m);
                    if (f < 0) {
                        scan = ncl;
                        break;
                    }
                    scan = f + 8 * T;
                    fq[issued & 3] = f;
                    mq[issued & 3] = m;
                    if (leader) issue(f, (int)((tma_base + issued) % D));
                    issued++;
                }
            };
            pump();
            const int rstart = k0 & ~7;
            while (consumed < issued) {
                cur = fq[consumed & 3];
                cm = mq[consumed & 3];
                const unsigned use = tma_base + consumed;
                const int buf = (int)(use % D);
                mbar_wait(&full[buf], (use / D) & 1);
                const double* sg = stage + (size_t)buf * strip_d;
                // ---- partial Y = V^T W over the 8-row groups of this warp ----
                double y0[MT][2], y1[MT][2];
#pragma unroll
                for (int mt = 0; mt < MT; mt++) y0[mt][0] = y0[mt][1] = y1[mt][0] = y1[mt][1] = 0.0;
                for (int r = rstart + 8 * tw; r < ldv; r += 8 * TW) {
                    const int rr = r + 2 * q;
                    const bool rok = rr < ldv;  // loads are predicated per lane, the mma is issued by the whole warp
                    double2 b2 = make_double2(0.0, 0.0);
                    if (rok) b2 = *reinterpret_cast<const double2*>(sg + (size_t)g * ld + rr);
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) {
                        if (mt * 8 < jc) {  // uniform
                            const int tt = mt * 8 + g;
                            double2 a2 = make_double2(0.0, 0.0);
                            if (tt < jc && rok) a2 = *reinterpret_cast<const double2*>(Vs + rr + (size_t)tt * ldv);
                            dmma_f64(y0[mt][0], y0[mt][1], a2.x, b2.x);
                            dmma_f64(y1[mt][0], y1[mt][1], a2.y, b2.y);
                        }
                    }
                }
                {
                    double* yw = ypart + (size_t)tw * NB * 8;
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) {
                        yw[(mt * 8 + g) * 8 + 2 * q] = y0[mt][0] + y1[mt][0];
                        yw[(mt * 8 + g) * 8 + 2 * q + 1] = y0[mt][1] + y1[mt][1];
                    }
                }
                team_barrier(1 + team, TW * 32);  // B1: partial tiles complete; the stage of the previous strip is free
                pump();
                if (tw == TW - 1 && prev >= 0) finish_norms(prev, cmp);
                // the team sums the partial tiles (each warp a slice, partials added in warp order)
                for (int o = tw * 32 + lane; o < NB * 8; o += TW * 32) {
                    double acc = ypart[o];
                    for (int w = 1; w < TW; w++) acc += ypart[(size_t)w * NB * 8 + o];
                    ysum[o] = acc;
                }
                team_barrier(1 + team, TW * 32);  // B2
                // B fragments of the update: F(col g, tt) = sum_{t <= tt} Y(t, g) T(t, tt), tt = 4 kk + q
                double bf[NB / 4];
#pragma unroll
                for (int kk = 0; kk < NB / 4; kk++) bf[kk] = 0.0;
                for (int tr = 0; tr < jc; tr++) {
                    const double yv = ysum[tr * 8 + g];
#pragma unroll
                    for (int kk = 0; kk < NB / 4; kk++) {
                        const int tt = kk * 4 + q;
                        if (tt >= tr && tt < jc) bf[kk] = fma(yv, Ts[tr * NB + tt], bf[kk]);
                    }
                }
                // ---- W -= V F^T on the 8-row groups of this warp, from the stage, stored to the panel ----
                const bool actA = (cm >> (2 * q)) & 1u, actB = (cm >> (2 * q + 1)) & 1u;
                const double* sA = sg + (size_t)(2 * q) * ld;
                const double* sB = sA + ld;
                double* colA = P + (size_t)(cur + 2 * q) * ld;
                double* colB = colA + ld;
                double nA = 0.0, nB = 0.0;
                for (int r = rstart + 8 * tw; r < rows; r += 16 * TW) {  // two groups in flight
                    double cA[2], cB[2];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int row = r + 8 * TW * h + g;
                        const bool in = row >= k0 && row < rows;
                        cA[h] = in ? sA[row] : 0.0;
                        cB[h] = in ? sB[row] : 0.0;
                    }
#pragma unroll
                    for (int kk = 0; kk < NB / 4; kk++) {
                        const int tt = kk * 4 + q;
                        if (kk * 4 < jc) {  // uniform
#pragma unroll
                            for (int h = 0; h < 2; h++) {
                                const int row = r + 8 * TW * h + g;
                                const double av = (tt < jc && row < ldv) ? -Vs[row + (size_t)tt * ldv] : 0.0;
                                dmma_f64(cA[h], cB[h], av, bf[kk]);
                            }
                        }
                    }
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int row = r + 8 * TW * h + g;
                        const bool in = row >= k0 && row < rows;
                        if (actA && in) {
                            colA[row] = cA[h];
                            if (row >= kend) nA = fma(cA[h], cA[h], nA);
                        }
                        if (actB && in) {
                            colB[row] = cB[h];
                            if (row >= kend) nB = fma(cB[h], cB[h], nB);
                        }
                    }
                }
#pragma unroll
                for (int o = 4; o <= 16; o <<= 1) {
                    nA += __shfl_xor_sync(FULL, nA, o);
                    nB += __shfl_xor_sync(FULL, nB, o);
                }
                if (g == 0) {
                    npart[tw * 8 + 2 * q] = nA;
                    npart[tw * 8 + 2 * q + 1] = nB;
                }
                prev = cur;
                cmp = cm;
                consumed++;
            }
            tma_uses += consumed;
            team_barrier(1 + team, TW * 32);
            if (tw == TW - 1 && prev >= 0) finish_norms(prev, cmp);
        } else {
            const int kend = k0 + jc;
            const int g = lane >> 2, q = lane & 3;
            double* yw = ysm + (size_t)warp * NB * 8;
            double* fw = ftmp + (size_t)warp * 8 * FLD;
            for (int first = warp * 8; first < ncl; first += 8 * NW) {
                unsigned cm = 0;
#pragma unroll
                for (int c = 0; c < 8; c++)
                    if (first + c < ncl && pos[first + c] >= kend && !state[first + c]) cm |= 1u << c;
                if (cm == 0) continue;
                double y0[MT][2], y1[MT][2];
#pragma unroll
                for (int mt = 0; mt < MT; mt++) y0[mt][0] = y0[mt][1] = y1[mt][0] = y1[mt][1] = 0.0;
                const bool colok = first + g < ncl;
                const double2* pc2 = reinterpret_cast<const double2*>(P + (size_t)(first + (colok ? g : 0)) * ld);
                // eight 16-byte loads in flight per lane (64 rows of the strip) before the tensor-core work on them:
                // the refresh is a streaming pass, its speed is the memory-level parallelism of this loop
                for (int rb = k0 & ~7; rb < ldv; rb += 64) {
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
                {
                    const int col = lane & 7;
                    for (int jj = lane >> 3; jj < jc; jj += 4) {
                        double f = 0.0;
                        for (int tt = 0; tt <= jj; tt++) f = fma(yw[tt * 8 + col], Ts[tt * NB + jj], f);
                        fw[col * FLD + jj] = f;
                    }
                }
                __syncwarp();
                double bf[NB / 4];
#pragma unroll
                for (int kk = 0; kk < NB / 4; kk++) {
                    const int tt = kk * 4 + q;
                    bf[kk] = (tt < jc) ? fw[g * FLD + tt] : 0.0;
                }
                const bool actA = (cm >> (2 * q)) & 1u, actB = (cm >> (2 * q + 1)) & 1u;
                double* colA = P + (size_t)(first + 2 * q) * ld;
                double* colB = colA + ld;
                double nA = 0.0, nB = 0.0;
                for (int r = k0 & ~7; r < rows; r += 32) {  // four tiles of 8 rows: eight loads in flight per lane
                    double cA[4], cB[4];
#pragma unroll
                    for (int h = 0; h < 4; h++) {
                        const int row = r + 8 * h + g;
                        const bool in = row >= k0 && row < rows;
                        cA[h] = (actA && in) ? colA[row] : 0.0;
                        cB[h] = (actB && in) ? colB[row] : 0.0;
                    }
#pragma unroll
                    for (int kk = 0; kk < NB / 4; kk++) {
                        const int tt = kk * 4 + q;
                        if (kk * 4 < jc) {  // uniform
#pragma unroll
                            for (int h = 0; h < 4; h++) {
                                const int row = r + 8 * h + g;
                                const double a = (tt < jc && row < ldv) ? -Vs[row + (size_t)tt * ldv] : 0.0;
                                dmma_f64(cA[h], cB[h], a, bf[kk]);
                            }
                        }
                    }
#pragma unroll
                    for (int h = 0; h < 4; h++) {
                        const int row = r + 8 * h + g;
                        const bool in = row >= k0 && row < rows;
                        if (actA && in) {
                            colA[row] = cA[h];
                            if (row >= kend) nA = fma(cA[h], cA[h], nA);
                        }
                        if (actB && in) {
                            colB[row] = cB[h];
                            if (row >= kend) nB = fma(cB[h], cB[h], nB);
                        }
                    }
                }
#pragma unroll
                for (int o = 4; o <= 16; o <<= 1) {
                    nA += __shfl_xor_sync(FULL, nA, o);
                    nB += __shfl_xor_sync(FULL, nB, o);
                }
                if (g == 0) {
                    if (actA) nq1[first + 2 * q] = nA;
                    if (actB) nq1[first + 2 * q + 1] = nB;
                }
                __syncwarp();
            }
        }
        __syncthreads();
        __syncthreads();
    };

    // ---- gather the own column range [c_lo, c_hi): one warp per source block; padding rows are zeroed ----
    for (int s = warp; s < t.nsrc; s += NW) {
        const int c0 = soff[s], c1 = soff[s + 1];
        const int a = max(c0, c_lo), b = min(c1, c_hi);
        if (a >= b) continue;
        const QrSrc q = src[s];
        const int nc = b - a, off = a - c0;
        double* dst = P + (size_t)(a - c_lo) * ld;
        if (!q.transposed) {  // rows x w: lanes along the rows
            const double* sp = q.blk + (size_t)off * q.ld;
            const int sld = q.ld;
            warp_tile_copy(rows, nc, lane, [&](int i, int jj) { return sp[i + (size_t)jj * sld]; },
                           [&](int i, int jj, double v) { dst[i + (size_t)jj * ld] = v; });
        } else {  // w x rows: lanes along the contiguous source rows
            const double* sp = q.blk + off;
            const int sld = q.ld;
            warp_tile_copy(nc, rows, lane, [&](int jj, int i) { return sp[jj + (size_t)i * sld]; },
                           [&](int jj, int i, double v) { dst[i + (size_t)jj * ld] = v; });
        }
    }
    {
        const int padr = ld - rows;
        for (int e = tid; e < padr * ncl; e += NT) P[rows + e % padr + (size_t)(e / padr) * ld] = 0.0;
    }
    __syncthreads();
    // ---- exact squared norms, identity positions ----
    for (int cl = warp; cl < ncl; cl += NW) {
        const double* cj = P + (size_t)cl * ld;
        double s = 0.0;
#pragma unroll 8
        for (int i = lane; i < rows; i += 32) s += cj[i] * cj[i];
        s = warp_sum(s);
        if (lane == 0) {
            nq1[cl] = s;
            pos[cl] = c_lo + cl;
            state[cl] = 0;
        }
    }
    __syncthreads();
    HT(9)

    int rank = mn;
    int k = 0, k0 = 0, j = 0;
    double coldmax_l = -1.0;
    bool flush = false;
    propose(block_best(best_of_slab(0)), 0, false, -1.0);
    csync();
    while (true) {
        const int par = (seq - 1) & 1;
        // ---- every CTA selects the same winner among the G proposals, and the largest cold bound ----
        int wi = 0;
        double coldmax_g = rec[par][0].cold;
        if constexpr (G > 2) {
            Cand c = (lane < G) ? Cand{rec[par][lane].val, rec[par][lane].pos, lane} : Cand{-2.0, INT_MAX, 0};
            wi = warp_best(c).col;
        } else if constexpr (G == 2) {
            if (better(rec[par][1].val, rec[par][1].pos, rec[par][0].val, rec[par][0].pos)) wi = 1;
        }
        if constexpr (G > 1) {
#pragma unroll
            for (int i = 1; i < G; i++) coldmax_g = fmax(coldmax_g, rec[par][i].cold);
        }
        const HcRec win = rec[par][wi];
        if (j > 0 && !(win.val > coldmax_g)) {
            // a column that was not kept current in this block may beat the best hot candidate: close the block,
            // after which every column is current, and choose again among all of them
            HT(7)
            close_block(j, k0);
            HC(2, 1)
            HT(8)
            k0 = k;
            j = 0;
            nh = 0;
            coldmax_l = -1.0;
            propose(block_best(best_of_slab(k)), k, false, -1.0);
            csync();
            HT(6)
            continue;
        }
        if (win.col < 0) {  // no admissible column (NaN norms): stop here
            rank = k;
            flush = j > 0;
            break;
        }
        const double beta = win.beta, tau = win.tau;
        if (k == 0) r00 = fabs(beta);
        if (win.stop) {  // geqp3 + choose_rank stop here
            rank = k;
            flush = j > 0;
            break;
        }
        if (j == 0) {
            // ---- a new block starts with pivot win.col: hot = own unpivoted columns within theta of its norm, the
            //      largest ones first as far as the capacity goes (histogram of quarter octaves below the level) ----
            const double level = win.val;
            auto bin_of = [&](double n1) {
                const float ratio = (float)(n1 / level);
                if (!(ratio > 1e-30f)) return HC2_BINS - 1;
                if (ratio >= 1.0f) return 0;
                return min(HC2_BINS - 1, (int)(-4.0f * __log2f(ratio)));
            };
            if (tid < HC2_BINS) hist[tid] = 0;
            if (tid == 0) nh_s = 0;
            __syncthreads();
            for (int cl = tid; cl < ncl; cl += NT)
                if (pos[cl] >= k && c_lo + cl != win.col) atomicAdd(&hist[bin_of(nq1[cl])], 1);
            __syncthreads();
            int blim = -1;
            {
                int cum = 0;
                for (int b = 0; b <= bins_theta; b++) {
                    cum += hist[b];
                    if (cum > HCAP) break;
                    blim = b;
                }
            }
            double cm = -1.0;
            for (int base = 0; base < ncl; base += NT) {
                const int cl = base + tid;
                bool h = false;
                if (cl < ncl) {
                    if (pos[cl] >= k && c_lo + cl != win.col) {
                        const double n1 = nq1[cl];
                        h = bin_of(n1) <= blim;
                        if (!h) cm = fmax(cm, n1);
                    }
                    state[cl] = h ? 1 : 0;
                }
                const unsigned m = __ballot_sync(FULL, h);
                if (m) {
                    int b0 = 0;
                    if (lane == 0) b0 = atomicAdd(&nh_s, __popc(m));
                    b0 = __shfl_sync(FULL, b0, 0);
                    if (h) {
                        const int e = b0 + __popc(m & ((1u << lane) - 1u));
                        const double n1 = nq1[cl];
                        hn1[e] = n1;
                        hn2[e] = n1;
                        hcol[e] = cl;
                        hpos[e] = pos[cl];
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cm = fmax(cm, __shfl_xor_sync(FULL, cm, o));
            if (lane == 0) wmax[warp] = cm;
            __syncthreads();
            cm = wmax[0];
#pragma unroll
            for (int w = 1; w < NW; w++) cm = fmax(cm, wmax[w]);
            coldmax_l = cm;
            nh = nh_s;
            // copy rows [k & ~1, ldv) of the own hot columns; one warp per slot, 16-byte loads
            for (int s = warp; s < nh; s += NW) {
                const double2* sp = reinterpret_cast<const double2*>(P + (size_t)hcol[s] * ld);
                double2* dp = reinterpret_cast<double2*>(H + (size_t)s * ldv);
#pragma unroll 4
                for (int pi = (k >> 1) + lane; pi < npair; pi += 32) dp[pi] = sp[pi];
            }
            HC(1, 1)
            HC(3, nh)
            HC(4, min(ncl, cols - k))
            HT(6)
        }
        // ---- the winner's reflector becomes column j of the block (zero outside [k, rows)) ----
        double* vj = Vs + (size_t)j * ldv;
        {
            if constexpr (XG) {
                const double* wv = t.X + ((size_t)wi * 2 + par) * ldv;
                for (int i = tid; i < ldv; i += NT) vj[i] = (i >= k && i < rows) ? __ldcg(wv + i) : 0.0;
            } else {
                const double* wv = slots + (size_t)par * ldv;
                if constexpr (G > 1) wv = cluster.map_shared_rank(wv, wi);
                for (int i = tid; i < ldv; i += NT) vj[i] = (i >= k && i < rows) ? wv[i] : 0.0;
            }
            if (tid == 0) {
                tau_s[j] = tau;
                beta_s[j] = beta;
                lcol[j] = win.col;
                lpos[j] = win.pos;
            }
        }
        const int ps = (wi == crank) ? my_slot : -1;  // hot slot of the pivot, if it is one of the own hot columns
        __syncthreads();
        // ---- apply it to the own hot columns (one warp per slot), downdate norms, pick the local candidate ----
        Cand best{-1.0, INT_MAX, -1};
        {
            const int kpair = k >> 1;
            const double2* v2 = reinterpret_cast<const double2*>(vj);
            constexpr int NU = RP <= 6 ? 2 : 1;  // hot columns a warp works on at once (independent shuffle chains)
            for (int s0 = warp; s0 < nh; s0 += NU * NW) {
                int sl[NU], p[NU];
                bool act[NU];
                double2* cp[NU];
                double ck[NU], w[NU], q[NU];
                double2 c[NU][RP];
#pragma unroll
                for (int u = 0; u < NU; u++) {
                    sl[u] = s0 + u * NW;
                    act[u] = sl[u] < nh;
                    p[u] = act[u] ? hpos[sl[u]] : -1;
                    if (act[u] && sl[u] == ps) {
                        if (lane == 0) hpos[sl[u]] = k;
                        act[u] = false;
                    }
                    if (act[u] && p[u] == k) {  // virtual swap: the column at position k takes the pivot's old place
                        p[u] = win.pos;
                        if (lane == 0) hpos[sl[u]] = p[u];
                    }
                    if (p[u] < k) act[u] = false;  // pivoted earlier in this block
                    cp[u] = reinterpret_cast<double2*>(H + (size_t)(act[u] ? sl[u] : 0) * ldv);
                    ck[u] = act[u] ? H[(size_t)sl[u] * ldv + k] : 0.0;
                    w[u] = 0.0;
                }
                if (!act[0] && !act[NU - 1]) continue;
#pragma unroll
                for (int m = 0; m < RP; m++) {
                    const int pi = lane + 32 * m;
                    const bool in = pi >= kpair && pi < npair;
                    const double2 vv = in ? v2[pi] : make_double2(0.0, 0.0);
#pragma unroll
                    for (int u = 0; u < NU; u++) {
                        c[u][m] = (in && act[u]) ? cp[u][pi] : make_double2(0.0, 0.0);
                        w[u] = fma(c[u][m].x, vv.x, w[u]);
                        w[u] = fma(c[u][m].y, vv.y, w[u]);
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                    for (int u = 0; u < NU; u++) w[u] += __shfl_xor_sync(FULL, w[u], o);
                }
                double f[NU];
#pragma unroll
                for (int u = 0; u < NU; u++) {
                    f[u] = tau * w[u];
                    q[u] = 0.0;
                }
#pragma unroll
                for (int m = 0; m < RP; m++) {
                    const int pi = lane + 32 * m;
                    const bool in = pi >= kpair && pi < npair;
                    const double2 vv = in ? v2[pi] : make_double2(0.0, 0.0);
                    const int r0 = 2 * pi;
#pragma unroll
                    for (int u = 0; u < NU; u++) {
                        c[u][m].x = fma(-f[u], vv.x, c[u][m].x);
                        c[u][m].y = fma(-f[u], vv.y, c[u][m].y);
                        if (in && act[u]) cp[u][pi] = c[u][m];
                        if (r0 > k) q[u] = fma(c[u][m].x, c[u][m].x, q[u]);
                        if (r0 + 1 > k) q[u] = fma(c[u][m].y, c[u][m].y, q[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < NU; u++) {
                    if (!act[u]) continue;  // warp-uniform
                    const double ak = ck[u] - f[u];  // row k of the updated column (v_k = 1)
                    const double n1 = hn1[sl[u]];
                    double newn = 0.0;
                    if (n1 != 0.0) {
                        newn = fmax(0.0, n1 - ak * ak);
                        if (newn <= tol3z * hn2[sl[u]]) {  // dlaqps / dlaqp2 safeguard: exact norm of the updated column
                            newn = warp_sum(q[u]);
                            if (lane == 0) hn2[sl[u]] = newn;
                        }
                    }
                    if (lane == 0) hn1[sl[u]] = newn;
                    if (better(newn, p[u], best.val, best.pos)) best = Cand{newn, p[u], sl[u]};
                }
            }
        }
        j++;
        k++;
        HC(0, 1)
        if (k >= mn) {  // factorization complete: rank = mn
            rank = mn;
            flush = true;
            break;
        }
        if (j == NB) {  // the block is full
            HT(7)
            close_block(j, k0);
            HT(8)
            k0 = k;
            j = 0;
            nh = 0;
            coldmax_l = -1.0;
            propose(block_best(best_of_slab(k)), k, false, -1.0);
        } else {
            propose(block_best(best), k, true, coldmax_l);
        }
        csync();
    }
    HT(7)
    // All CTAs leave the loop at the same step. One more barrier so that no CTA exits (or reuses its slots) while a
    // sibling may still be pulling from them.
    csync();
    if (rank >= rows) return;  // nothing to do (tree.cpp:1317-1319); csize unchanged
    // reflectors of the open block are still pending on the cold columns, hot columns still live in shared memory
    if (flush) close_block(j, k0);
    HT(8)

    // ---- scatter triu(R[:rank,:]) P^T back into the own columns of the edge blocks, in place (warp per source) ----
    for (int s = warp; s < t.nsrc; s += NW) {
        const int c0 = soff[s], c1 = soff[s + 1];
        const int a = max(c0, c_lo), b = min(c1, c_hi);
        if (a >= b) continue;
        const QrSrc q = src[s];
        const int nc = b - a, off = a - c0;
        const int a0 = a - c_lo;
        if (!q.transposed) {
            double* dp = q.blk + (size_t)off * q.ld;
            const int sld = q.ld;
            warp_tile_copy(rank, nc, lane,
                           [&](int i, int jj) {
                               const int p = pos[a0 + jj];
                               return (p >= rank || i <= p) ? P[i + (size_t)(a0 + jj) * ld] : 0.0;
                           },
                           [&](int i, int jj, double v) { dp[i + (size_t)jj * sld] = v; });
        } else {
            double* dp = q.blk + off;
            const int sld = q.ld;
            warp_tile_copy(nc, rank, lane,
                           [&](int jj, int i) {
                               const int p = pos[a0 + jj];
                               return (p >= rank || i <= p) ? P[i + (size_t)(a0 + jj) * ld] : 0.0;
                           },
                           [&](int jj, int i, double v) { dp[jj + (size_t)i * sld] = v; });
        }
    }
    if (crank == 0 && tid == 0) csize[t.cluster] = rank;  // nobody in this launch reads the size of this cluster
    HT(9)
}

// ------------------------------------------------------------------------------------------------
// Short panels (rows <= 128): one THREAD per column; panel in the shared memory of one CTA, or in L2 (GP).
// The lower levels have tens of thousands of panels of 9-47 rows and 100-600 columns. With lanes along the rows
// (the other kernels) such a panel pays warp reductions and several block barriers per Householder step for a few
// flops; here a thread owns whole columns and runs down their (short) rows serially: no reduction over rows at all,
// the reflector is rebuilt by every thread from the pivot column (broadcast reads), two block barriers per step.
// Unblocked QRCP with dlaqp2's arithmetic: reflector applied at once, partial norms downdated with the dlaqps /
// dlaqp2 safeguard (exact recomputation), LAPACK's first-index tie-breaking through virtual positions; stops at the
// first |R_kk| / |R_00| < tol (geqp3 + choose_rank, src/util.cpp:383-452).
// ------------------------------------------------------------------------------------------------
// GP: the panel lives in global scratch (t.W, L2-resident) instead of shared memory: for panels that do not fit one
// CTA's shared memory; a thread re-reads its own columns every step, which the L1 serves.
template <int NT, bool GP>
__global__ void __launch_bounds__(NT) rrqr_col_kernel(const QrTask* __restrict__ tasks, const QrSrc* __restrict__ srcs,
                                                       int* csize, double tol) {
    constexpr int NW = NT / 32;
    const QrTask t = tasks[blockIdx.x];
    const QrSrc* src = srcs + t.src0;
    const int rows = t.rows;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ Cand wbest[2][NW];
    __shared__ double rdiag[COL_MAXROWS];
    extern __shared__ __align__(16) double dsm[];
    const int ld = t.ld;  // odd: threads walk down neighbouring columns without bank conflicts
    const int mce = (t.maxcols + 3) & ~3;
    double* A = GP ? t.W : dsm;                    // ld x maxcols
    double* n1 = GP ? dsm : dsm + (((size_t)ld * t.maxcols + 1) & ~(size_t)1);
    double* n2 = n1 + mce;
    int* pos = (int*)(n2 + mce);
    int* soff = pos + mce;
    for (int s = tid; s < t.nsrc; s += NT) soff[s + 1] = csize[src[s].nbr];
    if (tid == 0) soff[0] = 0;
    __syncthreads();
    if (warp == 0) {
        const int per = (t.nsrc + 31) / 32;
        const int lo = 1 + lane * per, hi = min(t.nsrc + 1, lo + per);
        int sum = 0;
        for (int i = lo; i < hi; i++) sum += soff[i];
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
        }
        int run = incl - sum;
        for (int i = lo; i < hi; i++) {
            run += soff[i];
            soff[i] = run;
        }
    }
    __syncthreads();
    const int cols = soff[t.nsrc];
    if (rows == 0) return;
    if (tol >= 1.0 || cols == 0) {
        if (tid == 0) csize[t.cluster] = 0;
        return;
    }
    const int mn = min(rows, cols);
    // ---- gather (one warp per source block) ----
    for (int s = warp; s < t.nsrc; s += NW) {
        const int c0 = soff[s], nc = soff[s + 1] - c0;
        if (nc <= 0) continue;
        const QrSrc q = src[s];
        double* dst = A + (size_t)c0 * ld;
        const int sld = q.ld;
        if (!q.transposed) {
            const double* sp = q.blk;
            warp_tile_copy(rows, nc, lane, [&](int i, int jj) { return sp[i + (size_t)jj * sld]; },
                           [&](int i, int jj, double v) { dst[i + (size_t)jj * ld] = v; });
        } else {
            const double* sp = q.blk;
            warp_tile_copy(nc, rows, lane, [&](int jj, int i) { return sp[jj + (size_t)i * sld]; },
                           [&](int jj, int i, double v) { dst[i + (size_t)jj * ld] = v; });
        }
    }
    __syncthreads();
    for (int c = tid; c < cols; c += NT) {
        const double* a = A + (size_t)c * ld;
        double sm = 0.0;
#pragma unroll 4
        for (int i = 0; i < rows; i++) sm = fma(a[i], a[i], sm);
        n1[c] = sm;
        n2[c] = sm;
        pos[c] = c;
    }
    const double tol3z = sqrt(DBL_EPSILON);
    double r00 = 0.0;
    int rank = mn;
    for (int k = 0; k < mn; k++) {
        // ---- pivot: largest partial norm, smallest position on ties ----
        Cand b{-1.0, INT_MAX, -1};
        for (int c = tid; c < cols; c += NT) {
            const int p = pos[c];
            if (p >= k) {
                const double v = n1[c];
                if (better(v, p, b.val, b.pos)) b = Cand{v, p, c};
            }
        }
        {
            const int par = k & 1;
            b = warp_best(b);
            if (lane == 0) wbest[par][warp] = b;
            __syncthreads();  // also: every column update of the previous step is visible
            Cand bb = (lane < NW) ? wbest[par][lane] : Cand{-1.0, INT_MAX, -1};
            b = warp_best(bb);
        }
        const int pc = b.col, ppos = b.pos;
        if (pc < 0) {
            rank = k;
            break;
        }
        // ---- its reflector, rebuilt by every thread from the pivot column (nobody writes that column any more) ----
        const double* pv = A + (size_t)pc * ld;
        const double alpha = pv[k];
        double ss;
        {  // four independent partial sums: the serial runs down a column are latency chains otherwise
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int i = k + 1;
            for (; i + 3 < rows; i += 4) {
                s0 = fma(pv[i], pv[i], s0);
                s1 = fma(pv[i + 1], pv[i + 1], s1);
                s2 = fma(pv[i + 2], pv[i + 2], s2);
                s3 = fma(pv[i + 3], pv[i + 3], s3);
            }
            for (; i < rows; i++) s0 = fma(pv[i], pv[i], s0);
            ss = (s0 + s1) + (s2 + s3);
        }
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
        if (k == 0) r00 = fabs(beta);
        if (tol != 0.0 && !(fabs(beta) / r00 >= tol)) {
            rank = k;
            break;
        }
        if (tid == 0) {
            rdiag[k] = beta;
            t.tau[k] = tau;
        }
        for (int i = k + 1 + tid; i < rows; i += NT) t.V[i + (size_t)k * rows] = pv[i] * scal;
        // ---- apply it to the own columns, downdate their norms ----
        for (int c = tid; c < cols; c += NT) {
            int p = pos[c];
            if (c == pc) {
                pos[c] = k;
                continue;
            }
            if (p == k) {  // virtual swap: the column at position k takes the pivot's old place
                p = ppos;
                pos[c] = p;
            }
            if (p < k) continue;
            double* a = A + (size_t)c * ld;
            double w;
            {
                double w0 = 0.0, w1 = 0.0, w2 = 0.0, w3 = 0.0;
                int i = k + 1;
                for (; i + 3 < rows; i += 4) {
                    w0 = fma(pv[i], a[i], w0);
                    w1 = fma(pv[i + 1], a[i + 1], w1);
                    w2 = fma(pv[i + 2], a[i + 2], w2);
                    w3 = fma(pv[i + 3], a[i + 3], w3);
                }
                for (; i < rows; i++) w0 = fma(pv[i], a[i], w0);
                w = (w0 + w1) + (w2 + w3);
            }
            w = fma(w, scal, a[k]);
            const double f = tau * w;
            const double ak = a[k] - f;
            a[k] = ak;
            const double fs = f * scal;
            double q;
            {
                double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
                int i = k + 1;
                for (; i + 3 < rows; i += 4) {
                    const double x0 = fma(-fs, pv[i], a[i]), x1 = fma(-fs, pv[i + 1], a[i + 1]);
                    const double x2 = fma(-fs, pv[i + 2], a[i + 2]), x3 = fma(-fs, pv[i + 3], a[i + 3]);
                    a[i] = x0;
                    a[i + 1] = x1;
                    a[i + 2] = x2;
                    a[i + 3] = x3;
                    q0 = fma(x0, x0, q0);
                    q1 = fma(x1, x1, q1);
                    q2 = fma(x2, x2, q2);
                    q3 = fma(x3, x3, q3);
                }
                for (; i < rows; i++) {
                    const double x = fma(-fs, pv[i], a[i]);
                    a[i] = x;
                    q0 = fma(x, x, q0);
                }
                q = (q0 + q1) + (q2 + q3);
            }
            const double o1 = n1[c];
            if (o1 != 0.0) {
                double nn = fmax(0.0, o1 - ak * ak);
                if (nn <= tol3z * n2[c]) {  // dlaqps / dlaqp2 safeguard: exact norm of the updated column
                    nn = q;
                    n2[c] = q;
                }
                n1[c] = nn;
            }
        }
        // (the barrier of the next pivot search orders these updates before anybody reads them)
        if (k + 1 >= mn) __syncthreads();
    }
    __syncthreads();
    if (rank >= rows) return;  // nothing to do (tree.cpp:1317-1319); csize unchanged
    // ---- scatter triu(R[:rank,:]) P^T back into the edge blocks, in place (one warp per source block) ----
    for (int s = warp; s < t.nsrc; s += NW) {
        const int c0 = soff[s], nc = soff[s + 1] - c0;
        if (nc <= 0) continue;
        const QrSrc q = src[s];
        const int sld = q.ld;
        double* dp = q.blk;
        auto value = [&](int i, int jj) {
            const int p = pos[c0 + jj];
            if (p < rank) {
                if (i > p) return 0.0;
                if (i == p) return rdiag[p];
            }
            return A[i + (size_t)(c0 + jj) * ld];
        };
        if (!q.transposed)
            warp_tile_copy(rank, nc, lane, [&](int i, int jj) { return value(i, jj); },
                           [&](int i, int jj, double v) { dp[i + (size_t)jj * sld] = v; });
        else
            warp_tile_copy(nc, rank, lane, [&](int jj, int i) { return value(i, jj); },
                           [&](int jj, int i, double v) { dp[jj + (size_t)i * sld] = v; });
    }
    if (tid == 0) csize[t.cluster] = rank;
}

template <int NT, bool GP>
void launch_col(const QrTask* t, int nt, const QrSrc* s, int* csize, double tol, int smem, cudaStream_t st) {
    auto kern = rrqr_col_kernel<NT, GP>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    kern<<<nt, NT, smem, st>>>(t, s, csize, tol);
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess)
        throw std::runtime_error(std::string("rrqr (column kernel) launch failed (threads=") + std::to_string(NT) +
                                 ", smem=" + std::to_string(smem) + "): " + cudaGetErrorString(err));
}

template <int G, int NT, int NB, int RP, int MINB>
void launch_one_hc2(const QrTask* t, int nt, const QrSrc* s, int* csize, double tol, double theta2, int smem,
                    cudaStream_t st) {
    auto kern = rrqr_hc2_kernel<G, NT, NB, RP, MINB>;
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
    // SPAND_HC2_TMA=1: cold refresh through TMA-staged strips shared by teams of warps (slower than the direct
    // per-warp path on the shapes measured so far, kept selectable: profiles/r2_rrqr.md)
    static const int staged = getenv("SPAND_HC2_TMA") ? atoi(getenv("SPAND_HC2_TMA")) : 0;
    const cudaError_t err = cudaLaunchKernelEx(&cfg, kern, t, s, csize, tol, theta2, staged);
    if (err != cudaSuccess)
        throw std::runtime_error(std::string("rrqr (hot set) launch failed (G=") + std::to_string(G) + ", threads=" +
                                 std::to_string(NT) + ", smem=" + std::to_string(smem) + "): " + cudaGetErrorString(err));
}

template <int G>
void launch_hc2_g(const QrTask* t, int nt, const QrSrc* s, int* csize, double tol, double th2, int rp, int smem,
                  cudaStream_t st) {
    // 512 threads, one CTA per SM (128 registers per thread, up to 224 KB of shared memory for the hot set)
    // SPAND_HC2_NT=256: panels of up to 256 rows on 256-thread CTAs, two per SM (the pivot loop of one overlaps the
    // cold refresh of the other), with half the shared memory for the hot set
    if (rp <= 4 && hc2_threads(128) == 256) {
        if (rp <= 2) launch_one_hc2<G, 256, HC2_NB, 2, 2>(t, nt, s, csize, tol, th2, smem, st);
        else launch_one_hc2<G, 256, HC2_NB, 4, 2>(t, nt, s, csize, tol, th2, smem, st);
        return;
    }
    if (rp <= 2) launch_one_hc2<G, 512, HC2_NB, 2, 1>(t, nt, s, csize, tol, th2, smem, st);
    else if (rp <= 4) launch_one_hc2<G, 512, HC2_NB, 4, 1>(t, nt, s, csize, tol, th2, smem, st);
    else if (rp <= 6) launch_one_hc2<G, 512, HC2_NB, 6, 1>(t, nt, s, csize, tol, th2, smem, st);
    else launch_one_hc2<G, 512, HC2_NB, 10, 1>(t, nt, s, csize, tol, th2, smem, st);
}

}  // namespace

int hc2_row_pairs(int rows) {
    const int need = (((rows + 1) & ~1) / 2 + 31) / 32;
    if (need <= 2) return 2;
    if (need <= 4) return 4;
    if (need <= 6) return 6;
    if (need <= 10) return 10;
    return 0;  // too tall for the register-resident reflector: use the other kernels
}

int hc2_threads(int rows) {
    static const int nt_small = getenv("SPAND_HC2_NT") ? atoi(getenv("SPAND_HC2_NT")) : 512;
    return (nt_small == 256 && hc2_row_pairs(rows) <= 4 && hc2_row_pairs(rows) > 0) ? 256 : 512;
}

size_t hc2_smem_bytes(int rows, int maxcols, int G, int hcap, int nsrc) {
    const size_t ldv = ((size_t)rows + 1) & ~(size_t)1;
    const size_t cpcm = (size_t)(maxcols + G - 1) / G;
    const size_t cpce = (cpcm + 3) & ~(size_t)3;
    const size_t hcape = ((size_t)hcap + 3) & ~(size_t)3;
    const size_t nw = hc2_threads(rows) / 32;
    // the scratch of the cold refresh (per warp: F rows + Y tile of one strip) aliases the hot columns
    const size_t hot = hc2_hot_doubles(ldv, hcap, (int)nw);
    const size_t doubles = ldv * HC2_NB + hot + 2 * hcape + cpce;
    // + source offsets (ints) + two candidate reflector slots
    return doubles * sizeof(double) + (2 * cpce + 2 * hcape + (((size_t)nsrc + 2) & ~(size_t)1)) * sizeof(int) +
           2 * ldv * sizeof(double);
}

size_t hc2_exchange_doubles(int rows, int G) {  // candidate reflectors of wide clusters travel through L2
    const size_t ldv = ((size_t)rows + 1) & ~(size_t)1;
    return G >= 8 ? (size_t)G * 2 * ldv : 0;
}

void hc2_stats(unsigned long long* out16, bool reset) {
    cudaMemcpyFromSymbol(out16, g_hc2_stat, sizeof(unsigned long long) * 16);
    if (reset) {
        unsigned long long z[16] = {};
        cudaMemcpyToSymbol(g_hc2_stat, z, sizeof(z));
    }
}

// ---- column kernel (rows <= 64, panel in the shared memory of one CTA) ----
int rrqr_col_ld(int rows) { return rows | 1; }
int rrqr_col_max_rows() { return COL_MAXROWS; }
size_t rrqr_col_smem_bytes(int rows, int maxcols, int nsrc, bool global_panel) {
    const size_t ld = (size_t)rrqr_col_ld(rows), mce = ((size_t)maxcols + 3) & ~(size_t)3;
    const size_t doubles = (global_panel ? 0 : ((ld * maxcols + 1) & ~(size_t)1)) + 2 * mce;
    return doubles * sizeof(double) + (mce + (size_t)nsrc + 2) * sizeof(int);
}
void launch_rrqr_col(const QrTask* t, int nt, const QrSrc* s, int* csize, double tol, int nthreads, int smem,
                     bool global_panel, cudaStream_t st) {
    if (nt <= 0) return;
    if (global_panel) {
        if (nthreads <= 128) launch_col<128, true>(t, nt, s, csize, tol, smem, st);
        else if (nthreads <= 256) launch_col<256, true>(t, nt, s, csize, tol, smem, st);
        else launch_col<512, true>(t, nt, s, csize, tol, smem, st);
    } else {
        if (nthreads <= 128) launch_col<128, false>(t, nt, s, csize, tol, smem, st);
        else if (nthreads <= 256) launch_col<256, false>(t, nt, s, csize, tol, smem, st);
        else launch_col<512, false>(t, nt, s, csize, tol, smem, st);
    }
}

void launch_rrqr_hc2(const QrTask* t, int nt, const QrSrc* s, int* csize, double tol, int G, int row_pairs, int smem,
                     double theta, cudaStream_t st) {
    if (nt <= 0) return;
    const double th2 = theta * theta;
    switch (G) {
        case 1: launch_hc2_g<1>(t, nt, s, csize, tol, th2, row_pairs, smem, st); break;
        case 2: launch_hc2_g<2>(t, nt, s, csize, tol, th2, row_pairs, smem, st); break;
        case 4: launch_hc2_g<4>(t, nt, s, csize, tol, th2, row_pairs, smem, st); break;
        case 8: launch_hc2_g<8>(t, nt, s, csize, tol, th2, row_pairs, smem, st); break;
        default: launch_hc2_g<16>(t, nt, s, csize, tol, th2, row_pairs, smem, st); break;
    }
}

}  // namespace spand
