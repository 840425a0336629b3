#include "tree.hpp"
#include "host/parallel.hpp"

#include <sys/time.h>

#include <algorithm>
#include <cmath>
#include <limits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <unordered_map>
#include <stdexcept>
#include <string>
#include <thread>

namespace spand {



namespace {
double wtime() {
    timeval t;
    gettimeofday(&t, nullptr);
    return t.tv_sec + 1e-6 * t.tv_usec;
}
void cuda_check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e));
}
#define CK(x) cuda_check((x), #x)
}  // namespace

// ------------------------------------------------------------------------------------------------
// memory
// ------------------------------------------------------------------------------------------------
// device / pinned allocations of the process (first-call cost of a factorization), reported with SPAND_TIMING
static double g_malloc_seconds = 0;
static long g_malloc_calls = 0;
static size_t g_malloc_bytes = 0;

void* DeviceArena::alloc(size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes == 0) bytes = 256;
    while (true) {
        if (cur_ < chunks_.size()) {
            Chunk& c = chunks_[cur_];
            if (c.top + bytes <= c.cap) {
                void* p = c.p + c.top;
                c.top += bytes;
                used_ += bytes;
                return p;
            }
            cur_++;
            continue;
        }
        Chunk c;
        c.cap = std::max(chunk_, bytes);
        c.top = 0;
        const double t0 = wtime();
        CK(cudaMalloc((void**)&c.p, c.cap));
        g_malloc_seconds += wtime() - t0;
        g_malloc_calls++;
        g_malloc_bytes += c.cap;
        chunks_.push_back(c);
    }
}
void DeviceArena::reset() {
    for (auto& c : chunks_) c.top = 0;
    cur_ = 0;
    used_ = 0;
}
void DeviceArena::release() {
    for (auto& c : chunks_) cudaFree(c.p);
    chunks_.clear();
    cur_ = 0;
    used_ = 0;
}
size_t DeviceArena::capacity() const {
    size_t s = 0;
    for (auto& c : chunks_) s += c.cap;
    return s;
}

Stager::~Stager() {
    if (pinned_) cudaFreeHost(pinned_);
}
void Stager::reserve(size_t bytes) {
    if (bytes <= cap_) return;
    if (pinned_) cudaFreeHost(pinned_);
    cap_ = bytes;
    const double t0 = wtime();
    CK(cudaMallocHost((void**)&pinned_, cap_));
    g_malloc_seconds += wtime() - t0;
    g_malloc_calls++;
    top_ = 0;
}
void Stager::upload(void* dst, const void* src, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return;
    if (top_ + bytes > cap_) {
        // everything staged so far must have left the pinned buffer before it is recycled
        CK(cudaStreamSynchronize(st));
        if (bytes > cap_) reserve(std::max(bytes, 2 * cap_));
        top_ = 0;
    }
    // large descriptor arrays (tens of MB per level on the low levels): the staging copy on host threads
    char* const stage = pinned_ + top_;
    const char* const from = static_cast<const char*>(src);
    parallel_chunks(bytes, [&](int, size_t b, size_t e) { std::memcpy(stage + b, from + b, e - b); }, (size_t)4 << 20);
    CK(cudaMemcpyAsync(dst, pinned_ + top_, bytes, cudaMemcpyHostToDevice, st));
    top_ += (bytes + 63) & ~(size_t)63;
}

// ------------------------------------------------------------------------------------------------
// construction / partition / analyze / assemble
// ------------------------------------------------------------------------------------------------
Tree::Tree(int nlevels_) : nlevels(nlevels_) {
    if (nlevels <= 0) throw std::runtime_error("nlevels must be > 0");
    log.assign(nlevels, LevelLog());
}

Tree::~Tree() { free_device(); }

// Every public entry point that touches the device runs with the tree's device current and restores the caller's
// device afterwards (the host application may switch devices between calls, e.g. a second Tree on another GPU).
struct Tree::DeviceGuard {
    int prev = -1;
    bool active = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) {
            if (cudaSetDevice(dev) == cudaSuccess) active = true;
        } else {
            cudaGetLastError();  // no device at all: the entry point reports it itself
        }
    }
    ~DeviceGuard() {
        if (active) cudaSetDevice(prev);
    }
};

void Tree::ensure_device() {
    if (st_) return;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw std::runtime_error("spand_b200: no CUDA device available (this library has no CPU fallback)");
    CK(cudaSetDevice(device));
    CK(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
    arena_ = new DeviceArena((size_t)1 << 30);
    scratch_ = new DeviceArena((size_t)256 << 20);
    sym_arena_ = new DeviceArena((size_t)256 << 20);
    stager_.reserve((size_t)64 << 20);
    CK(cudaMalloc((void**)&d_err_, sizeof(int)));
    CK(cudaMalloc((void**)&d_cnt_, sizeof(int) * 64));
    for (int i = 0; i < kSide; i++) {
        CK(cudaStreamCreateWithFlags(&side_[i], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ev_join_[i], cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
}

void Tree::free_device() {
    DeviceGuard dev_guard(device);
    if (!st_) return;
    cudaStreamSynchronize(st_);
    for (int r = 0; r < mg_nranks; r++) {
        if (!mg_base_[r]) continue;
        if (r == mg_rank) cudaFree(mg_base_[r]);
        else cudaIpcCloseMemHandle(mg_base_[r]);
        mg_base_[r] = nullptr;
    }
    delete arena_;
    delete scratch_;
    delete sym_arena_;
    arena_ = scratch_ = sym_arena_ = nullptr;
    plan_valid_ = false;
    if (d_err_) cudaFree(d_err_);
    if (d_cnt_) cudaFree(d_cnt_);
    d_err_ = d_cnt_ = nullptr;
    for (int i = 0; i < kSide; i++) {
        cudaStreamDestroy(side_[i]);
        cudaEventDestroy(ev_join_[i]);
    }
    cudaEventDestroy(ev_fork_);
    for (auto e : ev_pool_) cudaEventDestroy(e);
    ev_pool_.clear();
    cudaStreamDestroy(st_);
    st_ = nullptr;
}

const char* Tree::family_name(int f) {
    static const char* names[F_COUNT] = {"potrf", "trsm", "gemm", "rrqr", "copy"};
    return (f >= 0 && f < F_COUNT) ? names[f] : "";
}

cudaEvent_t Tree::fam_begin(int fam) {
    family_launches[fam]++;
    if (!profile_families) return nullptr;
    cudaEvent_t a;
    if (ev_pool_.empty()) CK(cudaEventCreate(&a));
    else {
        a = ev_pool_.back();
        ev_pool_.pop_back();
    }
    CK(cudaEventRecord(a, st_));
    return a;
}

void Tree::fam_end(int fam, cudaEvent_t a) {
    if (!profile_families) return;
    cudaEvent_t b;
    if (ev_pool_.empty()) CK(cudaEventCreate(&b));
    else {
        b = ev_pool_.back();
        ev_pool_.pop_back();
    }
    CK(cudaEventRecord(b, st_));
    fam_events_.push_back({fam, a, b});
}

void Tree::fam_resolve() {
    for (auto& e : fam_events_) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e.a, e.b));
        family_ms[e.fam] += ms;
        ev_pool_.push_back(e.a);
        ev_pool_.push_back(e.b);
    }
    fam_events_.clear();
}

template <class T>
T* Tree::to_device(const std::vector<T>& v, DeviceArena* where) {
    if (v.empty()) return nullptr;
    T* d = where->alloc_n<T>(v.size());
    stager_.upload(d, v.data(), v.size() * sizeof(T), st_);
    return d;
}

int* Tree::next_counter() {
    if (cnt_next_ >= 64) throw std::runtime_error("internal: out of work-list counters");
    return d_cnt_ + cnt_next_++;
}

void Tree::set_coords(int dim, int N_, const double* X) {
    Xcoo_ = DenseMat(dim, N_);
    std::copy(X, X + (size_t)dim * N_, Xcoo_.a.begin());
    have_coords_ = true;
}

// src/tree.cpp:306-419
void Tree::partition(const SpMat& A) {
    if (use_geo && !have_coords_) throw std::runtime_error("use_geo set without coordinates");
    ord = build_ordering(A, nlevels, use_geo ? &Xcoo_ : nullptr, verb);
    ord_serial_++;
    N = A.rows;
    // a new partition invalidates everything built on the previous one (plan, blocks, factors, solve batches)
    assembled_ = false;
    factorized_ = false;
    plan_valid_ = false;
    state_level_ = -1;
    log.assign(nlevels, LevelLog());
    for (auto& p : ord.part) log[p.self.lvl].dofs_nd += 1;
    for (int l = nlevels - 2; l >= 0; l--) log[l].dofs_left_nd = log[l + 1].dofs_nd + log[l + 1].dofs_left_nd;
}

// Symbolic analysis (once per partition + pattern): leaf block structure (src/tree.cpp:505-575), the map from the
// non-zeros of A to their position inside the dense leaf blocks (util.cpp:454-486 block2dense), and the plan of
// every level (symbolic.hpp). Everything is uploaded into sym_arena_ and reused by later assemble() calls.
void Tree::analyze_host(const SpMat& A, std::vector<unsigned>& valmap) {
    const bool timing = getenv("SPAND_TIMING") != nullptr;
    double tm0 = wtime();
    auto lap = [&](const char* what) {
        if (!timing) return;
        const double now = wtime();
        fprintf(stderr, "[spand] analyze: %-30s %8.2f ms\n", what, (now - tm0) * 1e3);
        tm0 = now;
    };
    const bool symm = symmetry();
    const int ncl = ord.norders;
    std::vector<int> pinv(N), cmap(N);
    for (int i = 0; i < N; i++) pinv[ord.perm[i]] = i;
    for (int c : bottoms_[0])
        for (int k = cl_[c].start; k < cl_[c].start + cl_[c].size; k++) cmap[k] = c;
    // pass 1: neighbours of every leaf column cluster, pivot first then increasing order (cluster.cpp:113-176);
    // the leaves are independent: chunks of them on host threads, stitched together in leaf order afterwards
    std::vector<int> leaf_n1, leaf_n2;
    std::vector<int> blk_begin(ncl + 1, 0);
    leaf_off_.clear();
    size_t total = 0;
    {
        const std::vector<int>& leaves = bottoms_[0];
        struct Part {
            std::vector<int> n2;     // neighbour lists, pivot first
            std::vector<int> count;  // per leaf
        };
        std::vector<Part> parts(8);
        const int nth = parallel_chunks(leaves.size(), [&](int t, size_t b, size_t e) {
            Part& P = parts[t];
            std::vector<int> mark(ncl, -1), nb;
            for (size_t x = b; x < e; x++) {
                const int s = leaves[x];
                nb.clear();
                const Cluster& cs = cl_[s];
                for (int pj = cs.start; pj < cs.start + cs.size; pj++) {
                    const int j = ord.perm[pj];
                    for (int k = A.colptr[j]; k < A.colptr[j + 1]; k++) {
                        const int pi = pinv[A.rowind[k]];
                        if (symm && pi < pj) continue;
                        const int n = cmap[pi];
                        if (mark[n] != s) {
                            mark[n] = s;
                            nb.push_back(n);
                        }
                    }
                }
                mark[s] = s;
                std::sort(nb.begin(), nb.end());
                const size_t before = P.n2.size();
                P.n2.push_back(s);
                for (int n : nb)
                    if (n != s) P.n2.push_back(n);
                P.count.push_back((int)(P.n2.size() - before));
            }
        }, 4096);
        size_t nblocks = 0;
        for (int t = 0; t < nth; t++) nblocks += parts[t].n2.size();
        leaf_n1.reserve(nblocks);
        leaf_n2.reserve(nblocks);
        leaf_off_.reserve(nblocks);
        size_t x = 0;
        for (int t = 0; t < nth; t++) {
            const Part& P = parts[t];
            size_t q = 0;
            for (size_t i = 0; i < P.count.size(); i++, x++) {
                const int s = leaves[x];
                const size_t ssz = cl_[s].size;
                blk_begin[s] = (int)leaf_n1.size();
                for (int c = 0; c < P.count[i]; c++, q++) {
                    const int n = P.n2[q];
                    leaf_n1.push_back(s);
                    leaf_n2.push_back(n);
                    leaf_off_.push_back(total);
                    total += (((size_t)cl_[n].size * ssz) + 31) & ~(size_t)31;
                }
                blk_begin[s + 1] = (int)leaf_n1.size();
            }
        }
    }
    if (total >= 0xffffffffull) throw std::runtime_error("assemble: leaf blocks exceed the 32-bit value map");
    leaf_total_ = total;
    lap("leaf block structure");
    // pass 2: value map (independent per leaf: host threads)
    valmap.assign(A.nnz(), 0xffffffffu);
    {
        const std::vector<int>& leaves = bottoms_[0];
        parallel_chunks(leaves.size(), [&](int, size_t b, size_t e) {
            for (size_t x = b; x < e; x++) {
                const int s = leaves[x];
                const Cluster& cs = cl_[s];
                const int b0 = blk_begin[s], b1 = blk_begin[s + 1];
                for (int pj = cs.start; pj < cs.start + cs.size; pj++) {
                    const int j = ord.perm[pj];
                    for (int k = A.colptr[j]; k < A.colptr[j + 1]; k++) {
                        const int pi = pinv[A.rowind[k]];
                        if (symm && pi < pj) continue;
                        const int n = cmap[pi];
                        int bb = b0;
                        if (n != s)
                            bb = (int)(std::lower_bound(leaf_n2.begin() + b0 + 1, leaf_n2.begin() + b1, n) - leaf_n2.begin());
                        valmap[k] = (unsigned)(leaf_off_[bb] + (size_t)(pi - cl_[n].start) + (size_t)(pj - cs.start) * cl_[n].size);
                    }
                }
            }
        }, 4096);
    }
    lap("value map");
    // plan
    std::vector<SymCluster> sc(ncl);
    for (int c = 0; c < ncl; c++)
        sc[c] = SymCluster{cl_[c].level, cl_[c].hlevel, cl_[c].parent, cl_[c].child_begin, cl_[c].child_end, cl_[c].sparsify};
    build_symbolic(sc, bottoms_, leaf_n1, leaf_n2, symm, use_want_sparsify, plan_);
    lap("build_symbolic (all levels)");
    pat_colptr_ = A.colptr;
    pat_rowind_ = A.rowind;
    plan_ord_serial_ = ord_serial_;
    plan_host_valid_ = true;
}

void Tree::analyze(const SpMat& A) {
    DeviceGuard dev_guard(device);
    const double t0 = wtime();
    const int ncl = ord.norders;
    std::vector<unsigned> valmap;
    plan_valid_ = false;
    analyze_host(A, valmap);
    // upload
    CK(cudaStreamSynchronize(st_));
    sym_arena_->reset();
    d_valmap_ = to_device(valmap, sym_arena_);
    d_en1_ = to_device(plan_.en1, sym_arena_);
    d_en2_ = to_device(plan_.en2, sym_arena_);
    {
        std::vector<int> par(ncl);
        for (int c = 0; c < ncl; c++) par[c] = cl_[c].parent;
        d_parent_ = to_device(par, sym_arena_);
    }
    dplan_.assign(nlevels, DevLevel());
    size_t maxlist = 1;
    for (int l = 0; l < nlevels; l++) {
        const SymLevel& L = plan_.lv[l];
        DevLevel& D = dplan_[l];
        D.E = to_device(L.E, sym_arena_);
        D.e_piv = to_device(L.e_piv, sym_arena_);
        D.S = to_device(L.S, sym_arena_);
        D.s_piv = to_device(L.s_piv, sym_arena_);
        D.e_out = to_device(L.e_out, sym_arena_);
        D.e_in = to_device(L.e_in, sym_arena_);
        D.s_right = to_device(L.s_right, sym_arena_);
        D.s_left = to_device(L.s_left, sym_arena_);
        D.e_gemm = to_device(L.e_gemm, sym_arena_);
        D.e_con = to_device(L.e_con, sym_arena_);
        D.e_gf = to_device(L.e_gf, sym_arena_);
        D.e_gb = to_device(L.e_gb, sym_arena_);
        D.e_gfc = to_device(L.e_gfc, sym_arena_);
        D.e_gbc = to_device(L.e_gbc, sym_arena_);
        D.qs = to_device(L.qs, sym_arena_);
        D.m_copy = to_device(L.m_copy, sym_arena_);
        if (l + 1 < nlevels) {
            std::vector<int> ch;
            for (int p : bottoms_[l + 1])
                for (int c = cl_[p].child_begin; c < cl_[p].child_end; c++) ch.push_back(c);
            D.children = to_device(ch, sym_arena_);
            D.n_children = (int)ch.size();
        }
        maxlist = std::max({maxlist, L.E.size(), L.S.size(), L.e_out.size(), L.s_right.size(), L.e_gemm.size()});
    }
    d_mid_ = sym_arena_->alloc_n<int>(maxlist);
    CK(cudaStreamSynchronize(st_));
    stager_.reset();
    plan_valid_ = true;
    t_analyze_ = wtime() - t0;
    if (verb)
        printf("symbolic analysis: %zu edges, plan %.1f MB, %.3f s\n", plan_.en1.size(), plan_.bytes() / 1e6, t_analyze_);
}

// clusters from the ordering; id == order (src/tree.cpp:360-415)
void Tree::build_clusters() {
    cl_.assign(ord.norders, Cluster());
    bottoms_.assign(nlevels, {});
    std::vector<int> first_id(nlevels, 0);
    for (int h = 0; h < nlevels; h++) {
        first_id[h] = ord.levels[h].empty() ? 0 : ord.levels[h][0].order;
        for (auto& cn : ord.levels[h]) {
            Cluster& c = cl_[cn.order];
            c.start = cn.start;
            c.size = c.orig_size = cn.size;
            c.level = cn.level;
            c.sparsify = cn.sparsify;
            c.eliminated = false;
            c.parent = -1;
            c.hlevel = h;
            c.child_begin = c.child_end = 0;
            if (h > 0) {
                c.child_begin = first_id[h - 1] + cn.child_begin;
                c.child_end = first_id[h - 1] + cn.child_end;
                for (int k = c.child_begin; k < c.child_end; k++) cl_[k].parent = cn.order;
            }
            bottoms_[h].push_back(cn.order);
        }
    }
}


// Host-only symbolic analysis (no device needed): used by the CPU tests of the planner.
void Tree::analyze_only(const SpMat& A) {
    if (N == 0 || A.rows != N) throw std::runtime_error("analyze: call partition first with a matrix of the same size");
    build_clusters();
    std::vector<unsigned> valmap;
    plan_valid_ = false;
    analyze_host(A, valmap);
}

// Blocks alive after phase `phase` of level `level` (0: eliminate, 1: scale, 2: sparsify, 3: merge; level < 0: as
// assembled), as (column cluster, row cluster) pairs.
void Tree::plan_live_edges(int level, int phase, std::vector<int>& n1, std::vector<int>& n2) const {
    if (!plan_host_valid_) throw std::runtime_error("plan_live_edges: no plan");
    std::vector<int> live;
    if (level < 0) {
        for (int e = 0; e < plan_.nleaf_edges; e++) live.push_back(e);
    } else if (phase == 3) {
        const SymLevel& L = plan_.lv[level];
        for (int e = L.medge0; e < L.medge1; e++) live.push_back(e);
    } else {
        const SymLevel& L = plan_.lv[level];
        for (int e : L.s_piv) live.push_back(e);
        for (const SymTrsm& r : L.s_right) live.push_back(r.eB);
    }
    n1.clear();
    n2.clear();
    for (int e : live) {
        n1.push_back(plan_.en1[e]);
        n2.push_back(plan_.en2[e]);
    }
}

void Tree::plan_counts(int level, long long out[12]) const {
    if (!plan_host_valid_) throw std::runtime_error("plan_counts: no plan");
    const SymLevel& L = plan_.lv[level];
    long long v[12] = {(long long)L.E.size(), (long long)L.e_out.size(), (long long)L.e_in.size(), L.fill1 - L.fill0,
                       (long long)L.e_gemm.size(), (long long)L.e_con.size(), (long long)L.S.size(),
                       (long long)L.s_right.size(), (long long)L.q.size(), L.ncolors, L.medge1 - L.medge0,
                       (long long)L.m_copy.size()};
    for (int i = 0; i < 12; i++) out[i] = v[i];
}

// src/tree.cpp:505-575 — the values go to the device as they are (CSC order) and are scattered into the dense
// leaf blocks by one kernel
void Tree::assemble(const SpMat& A) { assemble_impl(&A, A.rows, A.colptr.data(), A.rowind.data(), A.val.data()); }

// Borrowed CSC arrays (the C ABI's spand_assemble): when the pattern is the one the plan was built for — the usual
// case of repeated factorizations — nothing is copied on the host, the values go straight from the caller's buffer
// to the device. Otherwise the matrix is canonicalised (sorted, duplicates summed) and analysed first.
void Tree::assemble_csc(int n, const int* colptr, const int* rowind, const double* val) {
    const bool same = plan_valid_ && n == N && pat_colptr_.size() == (size_t)n + 1 &&
                      std::memcmp(pat_colptr_.data(), colptr, sizeof(int) * ((size_t)n + 1)) == 0 &&
                      std::memcmp(pat_rowind_.data(), rowind, sizeof(int) * pat_rowind_.size()) == 0;
    if (same) {
        assemble_impl(nullptr, n, colptr, rowind, val);
    } else {
        SpMat A = from_csc(n, colptr, rowind, val);
        assemble_impl(&A, n, A.colptr.data(), A.rowind.data(), A.val.data());
    }
}

// A == nullptr: the caller has verified that (colptr, rowind) equal the planned pattern
void Tree::assemble_impl(const SpMat* Afull, int n_in, const int* colptr, const int* rowind, const double* val) {
    DeviceGuard dev_guard(device);
    (void)rowind;
    if (N == 0 || n_in != N) throw std::runtime_error("assemble: call partition first with a matrix of the same size");
    const size_t nnz_in = (size_t)colptr[n_in];
    ensure_device();
    CK(cudaStreamSynchronize(st_));
    const bool timing = getenv("SPAND_TIMING") != nullptr;
    double tm0 = wtime();
    auto lap = [&](const char* what) {
        if (!timing) return;
        const double now = wtime();
        fprintf(stderr, "[spand] assemble: %-28s %8.2f ms\n", what, (now - tm0) * 1e3);
        tm0 = now;
    };
    arena_->reset();
    scratch_->reset();
    stager_.reset();
    solve_.assign(nlevels, SolveLevel());
    nnz_ = 0;
    launches_total = 0;
    factorized_ = false;
    current_bottom_ = 0;
    ilvl_ = 0;
    state_level_ = state_phase_ = -1;
    logs_final_ = true;
    for (auto& l : log) {
        LevelLog fresh;
        fresh.dofs_nd = l.dofs_nd;
        fresh.dofs_left_nd = l.dofs_left_nd;
        l = fresh;
    }
    build_clusters();
    lap("reset + build_clusters");
    const bool reuse = plan_valid_ && plan_ord_serial_ == ord_serial_ && plan_.symmetric == symmetry() &&
                       plan_.want_flag == use_want_sparsify &&
                       (Afull == nullptr || (pat_colptr_ == Afull->colptr && pat_rowind_ == Afull->rowind));
    if (!reuse) {
        if (Afull != nullptr) {
            analyze(*Afull);
        } else {  // same pattern but the plan depends on a setting that changed: analyse a copy
            SpMat A = from_csc(n_in, colptr, rowind, val);
            analyze(A);
        }
    }
    lap(reuse ? "pattern check (plan reused)" : "symbolic analysis");

    const int ncl = ord.norders;
    const size_t nedges = plan_.en1.size();
    d_pos_ = arena_->alloc_n<int>(ncl);
    d_xptr_ = arena_->alloc_n<double*>(ncl);
    d_ud_ = arena_->alloc_n<double*>(ncl);
    d_pperm_ = arena_->alloc_n<int*>(ncl);
    d_eptr_ = arena_->alloc_n<double*>(nedges);
    d_eld_ = arena_->alloc_n<int>(nedges);
    d_perm_ = arena_->alloc_n<int>(N);
    d_xnat_ = arena_->alloc_n<double>(N);
    d_owner_ = nullptr;
    d_dof_owner_ = nullptr;
    double* leaf_base[MG_MAX_RANKS] = {};
    if (!mg()) {
        d_csize_ = arena_->alloc_n<int>(ncl);
        d_xleaf_ = arena_->alloc_n<double>(N);
        leaf_base[0] = arena_->alloc_n<double>(leaf_total_);
    } else {
        // Symmetric layout of every rank's shared arena: [flags | cluster sizes | leaf solution segments | the leaf
        // blocks (every rank assembles all of them, only the owner's copy is used) | blocks created later].
        if (!mg_peers_set_) throw std::runtime_error("assemble: mg_setup / mg_set_peers must be called first");
        auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
        mg_off_csize_ = 4096;
        const size_t off_xleaf = up(mg_off_csize_ + sizeof(int) * ncl);
        mg_off_leaf_ = up(off_xleaf + sizeof(double) * N);
        mg_off_blocks_ = up(mg_off_leaf_ + sizeof(double) * leaf_total_);
        if (mg_off_blocks_ > mg_size_) throw std::runtime_error("multi-GPU shared arena too small (SPAND_MG_ARENA_GB)");
        for (int r = 0; r < mg_nranks; r++) {
            mg_top_[r] = mg_off_blocks_;
            mg_csize_.p[r] = mg_base_[r] + mg_off_csize_;
            mg_leaf_.p[r] = mg_base_[r] + off_xleaf;
            leaf_base[r] = reinterpret_cast<double*>(mg_base_[r] + mg_off_leaf_);
        }
        d_csize_ = reinterpret_cast<int*>(mg_base_[mg_rank] + mg_off_csize_);
        d_xleaf_ = reinterpret_cast<double*>(mg_base_[mg_rank] + off_xleaf);
        h_owner_ = owner_map(mg_nranks);
        d_owner_ = arena_->alloc_n<int>(ncl);
        stager_.upload(d_owner_, h_owner_.data(), sizeof(int) * ncl, st_);
        std::vector<signed char> dof_owner(N);
        for (int c : bottoms_[0])
            for (int k = cl_[c].start; k < cl_[c].start + cl_[c].size; k++) dof_owner[k] = (signed char)h_owner_[c];
        d_dof_owner_ = reinterpret_cast<signed char*>(arena_->alloc(N));
        stager_.upload(d_dof_owner_, dof_owner.data(), N, st_);
    }
    tab_ = DevTables{d_csize_, d_eptr_, d_eld_, d_en1_, d_en2_, d_xptr_, d_pos_, d_parent_, d_owner_, mg() ? mg_rank : -1};
    h_csize_.assign(ncl, 0);
    h_pos_.assign(ncl, 0);
    h_xptr_.assign(ncl, nullptr);
    h_eptr_.assign(nedges, nullptr);
    h_eld_.assign(nedges, 1);
    h_ud_.assign(ncl, nullptr);
    h_ipiv_.assign(ncl, nullptr);
    h_pperm_.assign(ncl, nullptr);
    size_pre_.assign(nlevels, {});
    size_post_.assign(nlevels, {});
    phases_done_.assign(nlevels, 0);
    for (int c : bottoms_[0]) {
        h_xptr_[c] = (mg() ? reinterpret_cast<double*>(mg_leaf_.p[h_owner_[c]]) : d_xleaf_) + cl_[c].start;
        h_csize_[c] = cl_[c].size;
    }
    double* dblocks = leaf_base[mg() ? mg_rank : 0];
    parallel_chunks((size_t)plan_.nleaf_edges, [&](int, size_t b, size_t e1) {
        for (size_t e = b; e < e1; e++) {
            h_eptr_[e] = leaf_base[mg() ? h_owner_[plan_.en1[e]] : 0] + leaf_off_[e];
            h_eld_[e] = std::max(1, cl_[plan_.en2[e]].size);
        }
    });
    lap("host tables");
    stager_.upload(d_perm_, ord.perm.data(), sizeof(int) * N, st_);
    stager_.upload(d_csize_, h_csize_.data(), sizeof(int) * ncl, st_);
    stager_.upload(d_xptr_, h_xptr_.data(), sizeof(double*) * ncl, st_);
    stager_.upload(d_eptr_, h_eptr_.data(), sizeof(double*) * plan_.nleaf_edges, st_);
    stager_.upload(d_eld_, h_eld_.data(), sizeof(int) * plan_.nleaf_edges, st_);
    // values
    double* d_val = scratch_->alloc_n<double>(nnz_in);
    CK(cudaMemcpyAsync(d_val, val, sizeof(double) * nnz_in, cudaMemcpyHostToDevice, st_));
    CK(cudaMemsetAsync(dblocks, 0, leaf_total_ * sizeof(double), st_));
    launch_scatter_values(d_val, d_valmap_, nnz_in, dblocks, st_);
    CK(cudaStreamSynchronize(st_));
    lap("uploads + scatter kernel");
    stager_.reset();
    scratch_->reset();
    assembled_ = true;
}

int Tree::ndofs_left() const {
    int n = 0;
    for (int c : bottoms_[current_bottom_])
        if (!cl_[c].eliminated) n += cl_[c].size;
    return n;
}

int Tree::level_max_size() const {
    int m = 0;
    for (int c : bottoms_[current_bottom_])
        if (!cl_[c].eliminated) m = std::max(m, h_csize_[c]);
    return m;
}

void Tree::check_error() {
    // launch-configuration errors are only reported by the launch itself / cudaGetLastError
    const cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) throw std::runtime_error(std::string("CUDA launch error: ") + cudaGetErrorString(le));
    if (mg()) {
        // Sharded: a non-SPD / singular pivot found by one rank must stop all of them at the same point (a rank that
        // left the level loop alone would leave its peers spinning in the next peer barrier). Word 512 of every
        // rank's shared arena carries its flag.
        PeerPtrs words{};
        for (int r = 0; r < mg_nranks; r++) words.p[r] = mg_base_[r] + 2048;
        launch_err_publish(d_err_, reinterpret_cast<int*>(mg_base_[mg_rank] + 2048), st_);
        mg_barrier();
        launch_err_or(words, mg_nranks, d_err_, st_);
        mg_barrier();  // nobody publishes again before everybody has read
    }
    int err = 0;
    CK(cudaMemcpyAsync(&err, d_err_, sizeof(int), cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));
    if (err & 1) throw std::runtime_error("Error: Non-SPD Pivot\n");
    if (err & 2) throw std::runtime_error("Error: Singular Pivot\n");
}

TrsmTask Tree::host_trsm(const SymTrsm& t, const double* diag) const {
    TrsmTask r{};
    r.B = h_eptr_[t.eB];
    r.ldb = h_eld_[t.eB];
    r.T = h_eptr_[t.eT];
    r.ldt = h_eld_[t.eT];
    r.m = h_csize_[t.cm];
    r.n = h_csize_[t.cn];
    r.diag = diag;
    return r;
}

// A block that peers may touch: from this rank's arena on one GPU, from the owner's shared arena otherwise (every
// rank replays every allocation, so that all pointer tables hold valid peer addresses).
double* Tree::alloc_block(int owner, size_t doubles) {
    if (!mg()) return arena_->alloc_n<double>(doubles);
    size_t bytes = (doubles * sizeof(double) + 255) & ~(size_t)255;
    if (bytes == 0) bytes = 256;
    if (mg_top_[owner] + bytes > mg_size_) throw std::runtime_error("multi-GPU shared arena too small (SPAND_MG_ARENA_GB)");
    double* p = reinterpret_cast<double*>(mg_base_[owner] + mg_top_[owner]);
    mg_top_[owner] += bytes;
    return p;
}

void Tree::mg_barrier() {
    if (!mg()) return;
    launch_peer_barrier(mg_flags_, mg_rank, mg_nranks, ++mg_epoch_, st_);
}

// Blocks of the edges [e0, e1) (fill-in of a level, or the parents' blocks of a merge) at the current sizes.
void Tree::alloc_edges(int e0, int e1, bool zero, LevelLog& lg) {
    if (e1 <= e0) return;
    if (mg()) {
        const size_t top0 = mg_top_[mg_rank];
        for (int e = e0; e < e1; e++) {
            const size_t rows = h_csize_[plan_.en2[e]], cols = h_csize_[plan_.en1[e]];
            h_eld_[e] = (int)std::max<size_t>(1, rows);
            h_eptr_[e] = alloc_block(h_owner_[plan_.en1[e]], rows * cols);
        }
        if (zero && mg_top_[mg_rank] > top0) {
            CK(cudaMemsetAsync(mg_base_[mg_rank] + top0, 0, mg_top_[mg_rank] - top0, st_));
            lg.launches++;
        }
    } else {
        size_t total = 0;
        for (int e = e0; e < e1; e++) {
            const size_t rows = h_csize_[plan_.en2[e]], cols = h_csize_[plan_.en1[e]];
            h_eld_[e] = (int)std::max<size_t>(1, rows);
            h_eptr_[e] = (double*)total;  // offset for now
            total += (rows * cols + 31) & ~(size_t)31;
        }
        double* base = arena_->alloc_n<double>(total + 32);
        for (int e = e0; e < e1; e++) h_eptr_[e] = base + (size_t)h_eptr_[e];
        if (zero) {
            auto ev = fam_begin(F_COPY);
            CK(cudaMemsetAsync(base, 0, total * sizeof(double), st_));
            fam_end(F_COPY, ev);
            lg.launches++;
        }
    }
    stager_.upload(d_eptr_ + e0, h_eptr_.data() + e0, sizeof(double*) * (e1 - e0), st_);
    stager_.upload(d_eld_ + e0, h_eld_.data() + e0, sizeof(int) * (e1 - e0), st_);
}

// ------------------------------------------------------------------------------------------------
// multi-GPU setup (one process per GPU; the handles travel through the caller's process group)
// ------------------------------------------------------------------------------------------------
void Tree::mg_setup(int rank, int nranks, size_t arena_bytes) {
    DeviceGuard dev_guard(device);
    if (nranks < 1 || nranks > MG_MAX_RANKS || (nranks & (nranks - 1)) != 0)
        throw std::runtime_error("mg_setup: the number of ranks must be a power of two <= 16");
    if (nranks > 1 && nranks > (1 << (nlevels - 1))) throw std::runtime_error("mg_setup: more ranks than sub-trees");
    ensure_device();
    mg_rank = rank;
    mg_nranks = nranks;
    mg_peers_set_ = false;
    if (nranks == 1) return;
    if (mg_base_[rank]) cudaFree(mg_base_[rank]);
    mg_size_ = arena_bytes;
    CK(cudaMalloc((void**)&mg_base_[rank], mg_size_));
    CK(cudaMemset(mg_base_[rank], 0, 4096));
    mg_epoch_ = 0;
}

void Tree::mg_get_handle(void* out64) const {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, mg_base_[mg_rank]));
    std::memcpy(out64, &h, 64);
}

void Tree::mg_set_peers(const void* handles) {
    DeviceGuard dev_guard(device);
    for (int r = 0; r < mg_nranks; r++) {
        if (r == mg_rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, (const char*)handles + 64 * r, 64);
        CK(cudaIpcOpenMemHandle((void**)&mg_base_[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    for (int r = 0; r < mg_nranks; r++) mg_flags_.p[r] = mg_base_[r];
    mg_peers_set_ = true;
}

// Owner of every cluster: the rank of the depth-g sub-tree (g = log2 ranks) its separator belongs to; the 2^g - 1
// separators above the sub-trees go to distinct ranks (the middle sub-tree below them). Parents keep the owner of
// their children (same separator), so merges and solution-segment copies are local.
std::vector<int> Tree::owner_map(int nranks) const {
    std::vector<int> own(ord.norders, 0);
    if (nranks <= 1) return own;
    int g = 0;
    while ((1 << g) < nranks) g++;
    const int Ls = nlevels - 1 - g;
    if (Ls < 0) throw std::runtime_error("owner_map: more ranks than sub-trees");
    for (int h = 0; h < nlevels; h++)
        for (auto& cn : ord.levels[h]) {
            const int lvl = cn.id.self.lvl, sep = cn.id.self.sep;
            int o;
            if (lvl <= Ls) o = sep >> (Ls - lvl);
            else {
                const int shift = lvl - Ls;
                o = (sep << shift) + (1 << (shift - 1));
            }
            own[cn.order] = o;
        }
    return own;
}

// ------------------------------------------------------------------------------------------------
// batched dense building blocks (blocked right-looking over 64-wide steps)
// ------------------------------------------------------------------------------------------------
void Tree::run_gemm(std::vector<GemmTask>& tasks, std::vector<GemmContrib>& contribs, LevelLog& lg) {
    if (tasks.empty()) return;
    std::vector<GemmTask> small, big;
    for (auto& t : tasks) {
        if (t.m == 0 || t.n == 0) continue;
        if ((long)t.m * t.n <= 1024) small.push_back(t);
        else big.push_back(t);
    }
    GemmContrib* dc = to_device(contribs, scratch_);
    if (!small.empty()) {
        GemmTask* dt = to_device(small, scratch_);
        auto ev = fam_begin(F_GEMM);
        launch_gemm_small(dt, (int)small.size(), dc, st_);
        fam_end(F_GEMM, ev);
        lg.launches++;
    }
    if (!big.empty()) {
        std::vector<int> prefix(big.size() + 1, 0);
        for (size_t i = 0; i < big.size(); i++)
            prefix[i + 1] = prefix[i] + ((big[i].m + 63) / 64) * ((big[i].n + 63) / 64);
        GemmTask* dt = to_device(big, scratch_);
        int* dp = to_device(prefix, scratch_);
        auto ev = fam_begin(F_GEMM);
        launch_gemm_tiled(dt, (int)big.size(), dc, dp, prefix.back(), st_);
        fam_end(F_GEMM, ev);
        lg.launches++;
    }
}

void Tree::run_potrf(std::vector<PotrfTask>& tasks, LevelLog& lg) {
    if (tasks.empty()) return;
    int maxn = 0;
    for (auto& t : tasks) maxn = std::max(maxn, t.n);
    PotrfTask* dt = to_device(tasks, scratch_);
    for (int j0 = 0; j0 < maxn; j0 += NB) {
        auto ev = fam_begin(F_POTRF);
        launch_potrf_step(dt, (int)tasks.size(), j0, d_err_, st_);
        fam_end(F_POTRF, ev);
        lg.launches++;
        if (j0 + NB >= maxn) break;
        std::vector<TrsmTask> panel;
        std::vector<GemmTask> upd;
        std::vector<GemmContrib> con;
        int max_m = 0;
        for (auto& t : tasks) {
            int rem = t.n - j0 - NB;
            if (rem <= 0) continue;
            TrsmTask p{};
            p.B = t.A + (j0 + NB) + (size_t)j0 * t.ld;
            p.T = t.A + j0 + (size_t)j0 * t.ld;
            p.ldb = p.ldt = t.ld;
            p.m = rem;
            p.n = NB;
            panel.push_back(p);
            max_m = std::max(max_m, rem);
            GemmTask g;
            g.C = t.A + (j0 + NB) + (size_t)(j0 + NB) * t.ld;
            g.ldc = t.ld;
            g.m = g.n = rem;
            g.c0 = (int)con.size();
            g.nc = 1;
            g.flags = GEMM_LOWER;
            upd.push_back(g);
            con.push_back({p.B, p.B, t.ld, t.ld, NB});
        }
        TrsmTask* dp = to_device(panel, scratch_);
        ev = fam_begin(F_TRSM);
        launch_trsm_step(TRSM_RLT, dp, (int)panel.size(), 0, max_m, st_);
        fam_end(F_TRSM, ev);
        lg.launches++;
        run_gemm(upd, con, lg);
    }
}

void Tree::run_trsm(int mode, std::vector<TrsmTask>& all, LevelLog& lg) {
    if (all.empty()) return;
    if (mode == TRSM_RLT || mode == TRSM_LLN) {
        // GEMM-only solves: invert the 64 x 64 diagonal blocks of every distinct triangle, then one CTA per strip
        std::unordered_map<const double*, double*> invs;
        std::vector<TrtriTask> tt;
        std::vector<TrsmTask> tasks;
        std::vector<int> prefix(1, 0);
        int max_n = 0;
        for (auto& t : all) {
            if (t.m == 0 || t.n == 0) continue;
            auto it = invs.find(t.T);
            if (it == invs.end()) {
                int nblk = (t.n + NB - 1) / NB;
                double* inv = scratch_->alloc_n<double>((size_t)nblk * NB * NB);
                it = invs.emplace(t.T, inv).first;
                tt.push_back({t.T, t.ldt, t.n, inv});
                max_n = std::max(max_n, t.n);
            }
            t.inv = it->second;
            tasks.push_back(t);
            prefix.push_back(prefix.back() + (t.m + NB - 1) / NB);
        }
        if (tasks.empty()) return;
        TrtriTask* dtt = to_device(tt, scratch_);
        TrsmTask* dt = to_device(tasks, scratch_);
        int* dp = to_device(prefix, scratch_);
        auto ev = fam_begin(F_TRSM);
        launch_trtri(dtt, (int)tt.size(), max_n, st_);
        launch_trsm_strip(mode, dt, (int)tasks.size(), dp, prefix.back(), st_);
        fam_end(F_TRSM, ev);
        lg.launches += 2;
        return;
    }
    // bin by the free dimension so that the 2-D grid is not dominated by empty strips
    std::vector<TrsmTask> bins[2];
    for (auto& t : all) {
        if (t.m == 0 || t.n == 0) continue;
        bins[t.m <= NB ? 0 : 1].push_back(t);
    }
    for (auto& tasks : bins) {
        if (tasks.empty()) continue;
        int maxn = 0, max_m = 0;
        for (auto& t : tasks) {
            maxn = std::max(maxn, t.n);
            max_m = std::max(max_m, t.m);
        }
        TrsmTask* dt = to_device(tasks, scratch_);
        for (int j0 = 0; j0 < maxn; j0 += NB) {
            auto ev = fam_begin(F_TRSM);
            launch_trsm_step(mode, dt, (int)tasks.size(), j0, max_m, st_);
            fam_end(F_TRSM, ev);
            lg.launches++;
            if (j0 + NB >= maxn) break;
            std::vector<GemmTask> upd;
            std::vector<GemmContrib> con;
            for (auto& t : tasks) {
                int rem = t.n - j0 - NB;
                if (rem <= 0) continue;
                GemmTask g;
                g.c0 = (int)con.size();
                g.nc = 1;
                g.ldc = t.ldb;
                if (mode == TRSM_RLT) {
                    g.C = t.B + (size_t)(j0 + NB) * t.ldb;
                    g.m = t.m;
                    g.n = rem;
                    g.flags = 0;
                    con.push_back({t.B + (size_t)j0 * t.ldb, t.T + (j0 + NB) + (size_t)j0 * t.ldt, t.ldb, t.ldt, NB});
                } else if (mode == TRSM_LLN) {
                    g.C = t.B + (j0 + NB);
                    g.m = rem;
                    g.n = t.m;
                    g.flags = GEMM_NN;
                    con.push_back({t.T + (j0 + NB) + (size_t)j0 * t.ldt, t.B + j0, t.ldt, t.ldb, NB});
                } else {
                    g.C = t.B + (size_t)(j0 + NB) * t.ldb;
                    g.m = t.m;
                    g.n = rem;
                    g.flags = GEMM_NN;
                    con.push_back({t.B + (size_t)j0 * t.ldb, t.T + j0 + (size_t)(j0 + NB) * t.ldt, t.ldb, t.ldt, NB});
                }
                upd.push_back(g);
            }
            run_gemm(upd, con, lg);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// ELIMINATE — src/tree.cpp:895-967 for every cluster of level ilvl (mutually non-adjacent). The batches are the
// id arrays of the plan; tasks with a dimension above SMALL_DIM go through the host-driven blocked path.
// ------------------------------------------------------------------------------------------------
void Tree::phase_eliminate(LevelLog& lg, SolveLevel& sl) {
    if (scale_kind == PLU) {
        phase_eliminate_plu(lg, sl);
        return;
    }
    const SymLevel& L = plan_.lv[ilvl_];
    const DevLevel& D = dplan_[ilvl_];
    if (L.E.empty()) return;
    const bool big = level_max_size() > SMALL_DIM;
    alloc_edges(L.fill0, L.fill1, false, lg);
    // pivots
    auto ev = fam_begin(F_POTRF);
    launch_potrf_sym(tab_, D.E, D.e_piv, (int)L.E.size(), d_mid_, next_counter(), d_err_, st_);
    fam_end(F_POTRF, ev);
    lg.launches += 2;
    if (big) {
        std::vector<PotrfTask> bp;
        for (size_t i = 0; i < L.E.size(); i++)
            if (h_csize_[L.E[i]] > SMALL_DIM && mine(L.E[i]))
                bp.push_back({h_eptr_[L.e_piv[i]], h_eld_[L.e_piv[i]], h_csize_[L.E[i]]});
        run_potrf(bp, lg);
    }
    // panels
    ev = fam_begin(F_TRSM);
    launch_trsm_sym(TRSM_RLT, tab_, D.e_out, (int)L.e_out.size(), d_mid_, next_counter(), st_);
    fam_end(F_TRSM, ev);
    lg.launches += 2;
    if (big) {
        std::vector<TrsmTask> bt;
        for (const SymTrsm& t : L.e_out) {
            const int m = h_csize_[t.cm], n = h_csize_[t.cn];
            if ((m > SMALL_DIM || n > SMALL_DIM) && m > 0 && n > 0 && mine(plan_.en1[t.eB]))
                bt.push_back(host_trsm(t, nullptr));
        }
        run_trsm(TRSM_RLT, bt, lg);
    }
    // Schur complement: a target is updated by its owner, which reads the panels of other ranks through NVLink
    mg_barrier();
    ev = fam_begin(F_GEMM);
    launch_gemm_sym(tab_, D.e_gemm, (int)L.e_gemm.size(), D.e_con, d_mid_, next_counter(), st_);
    fam_end(F_GEMM, ev);
    lg.launches += 2;
    if (big) {
        std::vector<GemmTask> tasks;
        std::vector<GemmContrib> con;
        for (const SymGemm& g : L.e_gemm) {
            const int m = h_csize_[plan_.en2[g.target]], n = h_csize_[plan_.en1[g.target]];
            if (!(m > SMALL_DIM || n > SMALL_DIM) || m == 0 || n == 0 || !mine(plan_.en1[g.target])) continue;
            GemmTask t;
            t.C = h_eptr_[g.target];
            t.ldc = h_eld_[g.target];
            t.m = m;
            t.n = n;
            t.c0 = (int)con.size();
            t.nc = g.nc;
            t.flags = g.flags;
            for (int ci = 0; ci < g.nc; ci++) {
                const SymCon& c = L.e_con[g.c0 + ci];
                con.push_back({h_eptr_[c.e1], h_eptr_[c.e2], h_eld_[c.e1], h_eld_[c.e2], h_csize_[plan_.en1[c.e1]]});
            }
            tasks.push_back(t);
        }
        run_gemm(tasks, con, lg);
    }
    // recorded operations (tree.cpp:909-910, :885-892) as solve batches; sizes are captured now
    sl.max_e = level_max_size();  // bounds every pivot / panel dimension of this level's recorded operations
    sl.n_e_trsv = (int)L.E.size();
    sl.e_trsv = arena_->alloc_n<TrsvTask>(L.E.size());
    launch_expand_trsv(tab_, D.E, D.e_piv, sl.n_e_trsv, sl.e_trsv, st_);
    sl.n_e_gemv_f = (int)L.e_gf.size();
    sl.e_gemv_f = arena_->alloc_n<GemvTask>(L.e_gf.size());
    sl.e_gemv_fc = arena_->alloc_n<GemvContrib>(L.e_gfc.size());
    launch_expand_gemv(tab_, D.e_gf, sl.n_e_gemv_f, D.e_gfc, (int)L.e_gfc.size(), sl.e_gemv_f, sl.e_gemv_fc, st_);
    sl.n_e_gemv_b = (int)L.e_gb.size();
    sl.e_gemv_b = arena_->alloc_n<GemvTask>(L.e_gb.size());
    sl.e_gemv_bc = arena_->alloc_n<GemvContrib>(L.e_gbc.size());
    launch_expand_gemv(tab_, D.e_gb, sl.n_e_gemv_b, D.e_gbc, (int)L.e_gbc.size(), sl.e_gemv_b, sl.e_gemv_bc, st_);
    lg.launches += 3;
    for (int s : L.E) cl_[s].eliminated = true;
}

// ------------------------------------------------------------------------------------------------
// GEN / PLU variants (src/tree.cpp:614-689, :735-742, :929-956, :838-853; src/util.cpp:183-227): same plan,
// pointer descriptors expanded on the host for every task
// ------------------------------------------------------------------------------------------------
void Tree::run_getrf(std::vector<GetrfTask>& tasks, LevelLog& lg) {
    if (tasks.empty()) return;
    std::vector<GetrfTask> small, big;
    int maxn = 0;
    for (auto& t : tasks) {
        if (t.n <= 0) continue;
        if (t.n > 4096) throw std::runtime_error("PLU pivot larger than 4096 is not supported by the solve kernels");
        if (t.n <= NB) small.push_back(t);
        else {
            big.push_back(t);
            maxn = std::max(maxn, t.n);
        }
    }
    if (!small.empty()) {
        GetrfTask* dt = to_device(small, scratch_);
        auto ev = fam_begin(F_POTRF);
        launch_getrf_small(dt, (int)small.size(), d_err_, st_);
        fam_end(F_POTRF, ev);
        lg.launches++;
    }
    if (big.empty()) return;
    GetrfTask* dt = to_device(big, scratch_);
    for (int j0 = 0; j0 < maxn; j0 += NB) {
        auto ev = fam_begin(F_POTRF);
        launch_getrf_panel(dt, (int)big.size(), j0, d_err_, st_);
        launch_getrf_laswp(dt, (int)big.size(), j0, maxn, st_);
        fam_end(F_POTRF, ev);
        lg.launches += 2;
        std::vector<TrsmTask> u12;
        std::vector<GemmTask> upd;
        std::vector<GemmContrib> con;
        int max_m = 0;
        for (auto& t : big) {
            int rem = t.n - j0 - NB;
            if (rem <= 0) continue;
            TrsmTask p{};
            p.B = t.A + j0 + (size_t)(j0 + NB) * t.ld;  // U12: NB x rem
            p.T = t.A + j0 + (size_t)j0 * t.ld;
            p.ldb = p.ldt = t.ld;
            p.m = rem;
            p.n = NB;
            u12.push_back(p);
            max_m = std::max(max_m, rem);
            GemmTask g;
            g.C = t.A + (j0 + NB) + (size_t)(j0 + NB) * t.ld;
            g.ldc = t.ld;
            g.m = g.n = rem;
            g.c0 = (int)con.size();
            g.nc = 1;
            g.flags = GEMM_NN;
            upd.push_back(g);
            con.push_back({t.A + (j0 + NB) + (size_t)j0 * t.ld, p.B, t.ld, t.ld, NB});
        }
        if (u12.empty()) break;
        TrsmTask* dp = to_device(u12, scratch_);
        ev = fam_begin(F_TRSM);
        launch_trsm_step(TRSM_LLU, dp, (int)u12.size(), 0, max_m, st_);
        fam_end(F_TRSM, ev);
        lg.launches++;
        run_gemm(upd, con, lg);
    }
    auto ev = fam_begin(F_POTRF);
    launch_getrf_finish(dt, (int)big.size(), st_);
    fam_end(F_POTRF, ev);
    lg.launches++;
}

void Tree::run_rowperm(std::vector<RowPermTask>& tasks, LevelLog& lg) {
    if (tasks.empty()) return;
    for (auto& t : tasks)
        if (t.n > 6144) throw std::runtime_error("PLU pivot larger than 6144 is not supported by the row permutation kernel");
    RowPermTask* dt = to_device(tasks, scratch_);
    auto ev = fam_begin(F_TRSM);
    launch_rowperm(dt, (int)tasks.size(), st_);
    fam_end(F_TRSM, ev);
    lg.launches++;
}

void Tree::alloc_plu(int c) {
    size_t n = std::max(1, h_csize_[c]);
    if (!mg()) {
        h_ud_[c] = arena_->alloc_n<double>(n);
        h_ipiv_[c] = arena_->alloc_n<int>(n);
        h_pperm_[c] = arena_->alloc_n<int>(n);
        return;
    }
    // sharded: diag(U) and the row permutation of a pivot are read by the owners of the blocks in its row / column
    // (peer memory), so they live in the owner's shared arena; every rank replays the allocation
    const int o = h_owner_[c];
    h_ud_[c] = alloc_block(o, n);
    h_ipiv_[c] = reinterpret_cast<int*>(alloc_block(o, (n + 1) / 2));
    h_pperm_[c] = reinterpret_cast<int*>(alloc_block(o, (n + 1) / 2));
}

// diag(U) / row permutation pointers of the pivots of the current level -> device tables read by the plan-driven kernels
void Tree::upload_plu_tables() {
    const std::vector<int>& bottom = bottoms_[current_bottom_];
    if (bottom.empty()) return;
    const int first = bottom.front(), span = bottom.back() - first + 1;
    stager_.upload(d_ud_ + first, h_ud_.data() + first, sizeof(double*) * span, st_);
    stager_.upload(d_pperm_ + first, h_pperm_.data() + first, sizeof(int*) * span, st_);
}

void Tree::phase_eliminate_plu(LevelLog& lg, SolveLevel& sl) {
    const SymLevel& L = plan_.lv[ilvl_];
    const DevLevel& D = dplan_[ilvl_];
    if (L.E.empty()) return;
    const bool big = level_max_size() > SMALL_DIM;
    alloc_edges(L.fill0, L.fill1, false, lg);
    std::vector<GetrfTask> getrf;
    std::vector<TrsvTask> trsv;
    // Sharded over several GPUs: a pivot is factored by the owner of its cluster, a block is transformed / updated by
    // the owner of its column cluster (which reads the factor and the permutation of the row cluster through peer
    // memory); the solve batches keep every entry, with size 0 on the ranks that do not own the segment.
    for (size_t i = 0; i < L.E.size(); i++) {
        const int s = L.E[i], piv = L.e_piv[i];
        alloc_plu(s);
        if (mine(s)) getrf.push_back({h_eptr_[piv], h_eld_[piv], h_csize_[s], h_ud_[s], h_ipiv_[s], h_pperm_[s]});
        trsv.push_back({h_eptr_[piv], h_xptr_[s], h_eld_[piv], mine(s) ? h_csize_[s] : 0, h_ud_[s], h_pperm_[s]});
    }
    run_getrf(getrf, lg);
    upload_plu_tables();
    mg_barrier();  // the in-edges of a pivot may belong to other ranks
    // panels with both dimensions <= 64: plan-driven (ids resolved on the device), the others through the blocked path
    auto ev = fam_begin(F_TRSM);
    launch_plu_sym(PLU_LEFT, tab_, nullptr, D.e_in, (int)L.e_in.size(), d_ud_, d_pperm_, d_mid_, next_counter(), st_);
    launch_plu_sym(PLU_RIGHT, tab_, D.e_out, nullptr, (int)L.e_out.size(), d_ud_, d_pperm_, d_mid_, next_counter(), st_);
    fam_end(F_TRSM, ev);
    lg.launches += 4;
    if (big) {
        std::vector<RowPermTask> rperm;
        std::vector<TrsmTask> left, right;
        auto is_big = [&](const SymTrsm& t) {
            const int m = h_csize_[t.cm], n = h_csize_[t.cn];
            return (m > SMALL_DIM || n > SMALL_DIM) && m > 0 && n > 0 && mine(plan_.en1[t.eB]);
        };
        for (const SymTrsm& t : L.e_in) {  // A[s,n] <- L^-1 P^T A[s,n]   (tree.cpp:668-676)
            if (!is_big(t)) continue;
            rperm.push_back({h_eptr_[t.eB], h_eld_[t.eB], h_csize_[t.cn], h_csize_[t.cm], h_pperm_[t.cn]});
            left.push_back(host_trsm(t, nullptr));
        }
        for (const SymTrsm& t : L.e_out)  // A[n,s] <- A[n,s] U^-1 (tree.cpp:681-689)
            if (is_big(t)) right.push_back(host_trsm(t, h_ud_[t.cn]));
        run_rowperm(rperm, lg);
        run_trsm(TRSM_LLN, left, lg);
        run_trsm(TRSM_RUN, right, lg);
    }
    mg_barrier();  // a Schur target is updated by its owner, which reads the panels of other ranks
    // Schur complement A[n1,n2] -= A[n1,s] A[s,n2] (tree.cpp:943-947): targets up to 64 x 64 plan-driven
    ev = fam_begin(F_GEMM);
    launch_gemm_sym(tab_, D.e_gemm, (int)L.e_gemm.size(), D.e_con, d_mid_, next_counter(), st_);
    fam_end(F_GEMM, ev);
    lg.launches += 2;
    if (big) {
        std::vector<GemmTask> tasks;
        std::vector<GemmContrib> con;
        for (const SymGemm& g : L.e_gemm) {
            const int m = h_csize_[plan_.en2[g.target]], n = h_csize_[plan_.en1[g.target]];
            if (!(m > SMALL_DIM || n > SMALL_DIM) || m == 0 || n == 0 || !mine(plan_.en1[g.target])) continue;
            GemmTask t;
            t.C = h_eptr_[g.target];
            t.ldc = h_eld_[g.target];
            t.m = m;
            t.n = n;
            t.c0 = (int)con.size();
            t.nc = g.nc;
            t.flags = g.flags;
            for (int ci = 0; ci < g.nc; ci++) {
                const SymCon& c = L.e_con[g.c0 + ci];
                con.push_back({h_eptr_[c.e1], h_eptr_[c.e2], h_eld_[c.e1], h_eld_[c.e2], h_csize_[plan_.en1[c.e1]]});
            }
            tasks.push_back(t);
        }
        run_gemm(tasks, con, lg);
    }
    // recorded operations: ScalingPLUQ, GemmOut (forward only), GemmIn (backward only)
    sl.e_trsv = to_device(trsv, arena_);
    sl.max_e = level_max_size();
    sl.n_e_trsv = (int)trsv.size();
    sl.n_e_gemv_f = (int)L.e_gf.size();
    sl.e_gemv_f = arena_->alloc_n<GemvTask>(L.e_gf.size());
    sl.e_gemv_fc = arena_->alloc_n<GemvContrib>(L.e_gfc.size());
    launch_expand_gemv(tab_, D.e_gf, sl.n_e_gemv_f, D.e_gfc, (int)L.e_gfc.size(), sl.e_gemv_f, sl.e_gemv_fc, st_);
    sl.n_e_gemv_b = (int)L.e_gb.size();
    sl.e_gemv_b = arena_->alloc_n<GemvTask>(L.e_gb.size());
    sl.e_gemv_bc = arena_->alloc_n<GemvContrib>(L.e_gbc.size());
    launch_expand_gemv(tab_, D.e_gb, sl.n_e_gemv_b, D.e_gbc, (int)L.e_gbc.size(), sl.e_gemv_b, sl.e_gemv_bc, st_);
    lg.launches += 2;
    for (int s : L.E) cl_[s].eliminated = true;
}

void Tree::phase_scale_plu(LevelLog& lg, SolveLevel& sl) {
    const SymLevel& L = plan_.lv[ilvl_];
    const DevLevel& D = dplan_[ilvl_];
    if (L.S.empty()) return;
    const bool big = level_max_size() > SMALL_DIM;
    std::vector<GetrfTask> getrf;
    std::vector<TrsvTask> trsv;
    for (size_t i = 0; i < L.S.size(); i++) {
        const int c = L.S[i], piv = L.s_piv[i];
        alloc_plu(c);
        if (mine(c)) getrf.push_back({h_eptr_[piv], h_eld_[piv], h_csize_[c], h_ud_[c], h_ipiv_[c], h_pperm_[c]});
        trsv.push_back({h_eptr_[piv], h_xptr_[c], h_eld_[piv], mine(c) ? h_csize_[c] : 0, h_ud_[c], h_pperm_[c]});
    }
    run_getrf(getrf, lg);
    upload_plu_tables();
    mg_barrier();  // a block needs the factor and the permutation of its row cluster too, possibly from another rank
    // block A[n2,n1] (|n2| x |n1|): out-edge of n1 -> B U_n1^-1 ; in-edge of n2 -> L_n2^-1 P_n2^T B. Both
    // transformations of a block run on the owner of its column cluster. Blocks up to 64 x 64: one fused plan-driven
    // pass (permuted load, two substitutions, one store); larger ones through the blocked path.
    auto ev = fam_begin(F_TRSM);
    launch_plu_sym(PLU_BOTH, tab_, D.s_right, D.s_left, (int)L.s_right.size(), d_ud_, d_pperm_, d_mid_, next_counter(), st_);
    fam_end(F_TRSM, ev);
    lg.launches += 2;
    if (scale_inv_mode_ < 0) {
        const char* e = std::getenv("SPAND_SCALE_INV");
        scale_inv_mode_ = e ? std::atoi(e) : 1;
    }
    if (big && scale_inv_mode_ >= 1) {
        run_scale_inv(SMALL_DIM, lg);  // explicit inverses + two grouped GEMM launches (row permutation in between)
    } else if (big) {
        std::vector<RowPermTask> rperm;
        std::vector<TrsmTask> right, left;
        for (size_t i = 0; i < L.s_right.size(); i++) {
            const SymTrsm& r = L.s_right[i];
            const SymTrsm& l = L.s_left[i];
            const int m = h_csize_[r.cm], n = h_csize_[r.cn];
            if (!((m > SMALL_DIM || n > SMALL_DIM) && m > 0 && n > 0 && mine(plan_.en1[r.eB]))) continue;
            right.push_back(host_trsm(r, h_ud_[r.cn]));
            rperm.push_back({h_eptr_[l.eB], h_eld_[l.eB], h_csize_[l.cn], h_csize_[l.cm], h_pperm_[l.cn]});
            left.push_back(host_trsm(l, nullptr));
        }
        run_trsm(TRSM_RUN, right, lg);
        run_rowperm(rperm, lg);
        run_trsm(TRSM_LLN, left, lg);
    }
    sl.s_trsv = to_device(trsv, arena_);
    sl.max_s = level_max_size();
    sl.n_s_trsv = (int)trsv.size();
}

// ------------------------------------------------------------------------------------------------
// SCALE — src/tree.cpp:796-856 for every remaining cluster: pivot = L L^T, every incident block
// A[n2,n1] <- L_n2^-1 A[n2,n1] L_n1^-T in one pass, pivot := I (kept implicit)
// ------------------------------------------------------------------------------------------------
void Tree::phase_scale(LevelLog& lg, SolveLevel& sl) {
    if (scale_kind == PLU) {
        phase_scale_plu(lg, sl);
        return;
    }
    const SymLevel& L = plan_.lv[ilvl_];
    const DevLevel& D = dplan_[ilvl_];
    if (L.S.empty()) return;
    const bool big = level_max_size() > SMALL_DIM;
    auto ev = fam_begin(F_POTRF);
    launch_potrf_sym(tab_, D.S, D.s_piv, (int)L.S.size(), d_mid_, next_counter(), d_err_, st_);
    fam_end(F_POTRF, ev);
    lg.launches += 2;
    if (big) {
        std::vector<PotrfTask> bp;
        for (size_t i = 0; i < L.S.size(); i++)
            if (h_csize_[L.S[i]] > SMALL_DIM && mine(L.S[i]))
                bp.push_back({h_eptr_[L.s_piv[i]], h_eld_[L.s_piv[i]], h_csize_[L.S[i]]});
        run_potrf(bp, lg);
    }
    mg_barrier();  // a block needs the factor of its row cluster too, possibly from another rank
    if (scale_inv_mode_ < 0) {
        const char* e = std::getenv("SPAND_SCALE_INV");
        scale_inv_mode_ = e ? std::atoi(e) : 1;  // measured best that keeps rank parity
    }
    const int lmax = level_max_size();
    // blocks with a dimension above inv_dim go through the explicit-inverse GEMM path (0: none on this level)
    const int inv_dim = scale_inv_mode_ == 2 ? (lmax > 32 ? 32 : 0) : (scale_inv_mode_ == 1 && big ? SMALL_DIM : 0);
    ev = fam_begin(F_TRSM);
    launch_scale_sym(tab_, D.s_right, D.s_left, (int)L.s_right.size(), d_mid_, next_counter(), st_, inv_dim == 32);
    fam_end(F_TRSM, ev);
    lg.launches += 2;
    if (inv_dim > 0) {
        run_scale_inv(inv_dim, lg);
    } else if (big) {
        std::vector<TrsmTask> right, left;
        for (size_t i = 0; i < L.s_right.size(); i++) {
            const SymTrsm& r = L.s_right[i];
            const int m = h_csize_[r.cm], n = h_csize_[r.cn];
            if ((m > SMALL_DIM || n > SMALL_DIM) && m > 0 && n > 0 && mine(plan_.en1[r.eB])) {
                right.push_back(host_trsm(r, nullptr));
                left.push_back(host_trsm(L.s_left[i], nullptr));
            }
        }
        run_trsm(TRSM_RLT, right, lg);
        run_trsm(TRSM_LLN, left, lg);
    }
    sl.max_s = level_max_size();
    sl.n_s_trsv = (int)L.S.size();
    sl.s_trsv = arena_->alloc_n<TrsvTask>(L.S.size());
    launch_expand_trsv(tab_, D.S, D.s_piv, sl.n_s_trsv, sl.s_trsv, st_);
    lg.launches++;
}

// Two-sided scaling of the blocks with a dimension above min_dim (src/tree.cpp:796-856: two cblas_dtrsm per block)
// as products with the explicit inverses W_c = L_c^-1 of the scaled pivots: Y = B W_c1^T into scratch, then
// B = W_c2 Y in place. Both passes are ONE grouped tensor-core GEMM launch over all blocks of the level with nothing
// sequential inside a block (a triangular solve sweeps the 64-wide steps of the triangle one after the other); the
// triangular shape of W is used through the inner limits GEMM_TRIB / GEMM_TRIA, so the flop count is that of the
// solves. W_c is computed once per cluster: triangles up to 64 by forward substitution on I (trtri_kernel), larger
// ones by the strip solve L X = I on top of the inverted diagonal blocks. When sharded, a rank inverts the pivots its
// own blocks need (the factor of a remote row cluster is read through peer memory, after the barrier above).
void Tree::run_scale_inv(int min_dim, LevelLog& lg) {
    const bool plu = scale_kind == PLU;
    const SymLevel& L = plan_.lv[ilvl_];
    const size_t nb_all = L.s_right.size();
    if (nb_all == 0) return;
    // 1. blocks of this rank above the size class of the warp kernel (threaded scan of the level's block list)
    std::vector<std::vector<int>> part(8);
    const int nth = parallel_chunks(nb_all, [&](int t, size_t b, size_t e) {
        std::vector<int>& out = part[t];
        for (size_t i = b; i < e; i++) {
            const SymTrsm& r = L.s_right[i];
            const int m = h_csize_[r.cm], n = h_csize_[r.cn];
            if ((m > min_dim || n > min_dim) && m > 0 && n > 0 && mine(plan_.en1[r.eB])) out.push_back((int)i);
        }
    });
    std::vector<int> idx;
    for (int t = 0; t < nth; t++) idx.insert(idx.end(), part[t].begin(), part[t].end());
    const size_t nt = idx.size();
    if (nt == 0) return;
    // 2. inverses of the pivots these blocks touch. LLT: W = L^-1 for both sides. PLU: the row side uses L^-1 of the
    // pivot (lower part of the block, its own diagonal), the column side (U^T)^-1 with U^T materialised as a lower
    // triangle (strictly upper part of the block + diag(U) kept beside it), so that B U^-1 = B ((U^T)^-1)^T is the same
    // NT product as in the LLT case.
    struct Inv {
        const double* W = nullptr;
        int ld = 0;
    };
    std::unordered_map<long long, Inv> inv;  // key: 2 * cluster + (1: column side of a PLU pivot)
    std::vector<TrtriTask> tt;
    std::vector<UtTask> ut;
    std::vector<EyeTask> eye;
    std::vector<TrsmTask> solve;
    std::vector<int> sprefix(1, 0);
    int max_n = 0;
    auto need = [&](int c, int eT, bool upper) {
        const long long key = 2LL * c + (upper ? 1 : 0);
        if (inv.count(key)) return;
        const int n = h_csize_[c];
        const double* T = h_eptr_[eT];
        int ldt = h_eld_[eT];
        if (upper) {
            double* Lt = scratch_->alloc_n<double>((size_t)n * n);
            ut.push_back({T, h_ud_[c], Lt, ldt, n});
            T = Lt;
            ldt = n;
        }
        Inv w;
        if (n <= NB) {
            const int ld = (n + 1) & ~1;
            double* W = scratch_->alloc_n<double>((size_t)ld * n);
            tt.push_back({T, ldt, n, W, ld});
            w.W = W;
            w.ld = ld;
        } else {
            const int nblk = (n + NB - 1) / NB;
            double* blocks = scratch_->alloc_n<double>((size_t)nblk * NB * NB);
            double* W = scratch_->alloc_n<double>((size_t)n * n);
            tt.push_back({T, ldt, n, blocks, 0});
            eye.push_back({W, n, n});
            TrsmTask s{};
            s.B = W;
            s.ldb = n;
            s.T = T;
            s.ldt = ldt;
            s.m = n;
            s.n = n;
            s.inv = blocks;
            s.tri = 1;
            solve.push_back(s);
            sprefix.push_back(sprefix.back() + nblk);
            w.W = W;
            w.ld = n;
        }
        max_n = std::max(max_n, n);
        inv.emplace(key, w);
    };
    for (size_t q = 0; q < nt; q++) {
        need(L.s_right[idx[q]].cn, L.s_right[idx[q]].eT, plu);
        need(L.s_left[idx[q]].cn, L.s_left[idx[q]].eT, false);
    }
    // 3. the two grouped products
    std::vector<size_t> yoff(nt + 1, 0);
    std::vector<int> prefix(nt + 1, 0);
    for (size_t q = 0; q < nt; q++) {
        const SymTrsm& r = L.s_right[idx[q]];
        const size_t m = h_csize_[r.cm], n = h_csize_[r.cn];
        yoff[q + 1] = yoff[q] + ((m * n + 31) & ~(size_t)31);
        prefix[q + 1] = prefix[q] + (int)(((m + 63) / 64) * ((n + 63) / 64));
    }
    // one scratch block per product (bump allocations inside the arena's chunks, which later levels reuse: one
    // allocation of the level's total would cost a multi-GB cudaMalloc at every level of the first factorization)
    std::vector<double*> Yptr(nt);
    for (size_t q = 0; q < nt; q++) Yptr[q] = scratch_->alloc_n<double>(yoff[q + 1] - yoff[q]);
    std::vector<GemmTask> g1(nt), g2(nt);
    std::vector<GemmContrib> c1(nt), c2(nt);
    std::vector<RowPermTask> rperm(plu ? nt : 0);
    std::vector<int> tile_task(prefix[nt]);
    parallel_chunks(nt, [&](int, size_t b, size_t e) {
        for (size_t q = b; q < e; q++) {
            const SymTrsm& r = L.s_right[idx[q]];
            const SymTrsm& l = L.s_left[idx[q]];
            const int m = h_csize_[r.cm], n = h_csize_[r.cn];
            const Inv& w1 = inv.find(2LL * r.cn + (plu ? 1 : 0))->second;
            const Inv& w2 = inv.find(2LL * l.cn)->second;
            double* B = h_eptr_[r.eB];
            const int ldb = h_eld_[r.eB];
            double* Y = Yptr[q];
            g1[q] = GemmTask{Y, m, m, n, (int)q, 1, GEMM_ZERO_INIT | GEMM_POS | GEMM_TRIB};
            c1[q] = GemmContrib{B, w1.W, ldb, w1.ld, n};
            g2[q] = GemmTask{B, ldb, m, n, (int)q, 1, GEMM_NN | GEMM_ZERO_INIT | GEMM_POS | GEMM_TRIA};
            c2[q] = GemmContrib{w2.W, Y, w2.ld, m, m};
            if (plu) rperm[q] = RowPermTask{Y, m, m, n, h_pperm_[l.cn]};  // rows of Y <- rows perm[.] (P^T of the row pivot)
            for (int x = prefix[q]; x < prefix[q + 1]; x++) tile_task[x] = (int)q;
        }
    });
    if (plu)
        for (auto& t : rperm)
            if (t.n > 6144) throw std::runtime_error("PLU pivot larger than 6144 is not supported by the row permutation kernel");
    UtTask* dut = to_device(ut, scratch_);
    TrtriTask* dtt = to_device(tt, scratch_);
    EyeTask* deye = to_device(eye, scratch_);
    TrsmTask* dsolve = to_device(solve, scratch_);
    int* dsp = to_device(sprefix, scratch_);
    GemmTask* dg1 = to_device(g1, scratch_);
    GemmTask* dg2 = to_device(g2, scratch_);
    GemmContrib* dc1 = to_device(c1, scratch_);
    GemmContrib* dc2 = to_device(c2, scratch_);
    RowPermTask* drp = to_device(rperm, scratch_);
    int* dp = to_device(prefix, scratch_);
    int* dtile = to_device(tile_task, scratch_);
    auto ev = fam_begin(F_TRSM);
    if (!ut.empty()) {
        launch_ut(dut, (int)ut.size(), st_);
        lg.launches++;
    }
    launch_trtri(dtt, (int)tt.size(), max_n, st_);
    if (!solve.empty()) {
        launch_eye(deye, (int)eye.size(), st_);
        launch_trsm_strip(TRSM_LLN, dsolve, (int)solve.size(), dsp, sprefix.back(), st_);
        lg.launches += 2;
    }
    launch_gemm_tiled(dg1, (int)nt, dc1, dp, prefix[nt], st_, dtile);
    if (plu) {
        launch_rowperm(drp, (int)nt, st_);
        lg.launches++;
    }
    launch_gemm_tiled(dg2, (int)nt, dc2, dp, prefix[nt], st_, dtile);
    fam_end(F_TRSM, ev);
    lg.launches += 3;
}

// ------------------------------------------------------------------------------------------------
// SPARSIFY — src/tree.cpp:1417-1433 -> 1292-1347. The reference sweeps the clusters in list order
// (Gauss-Seidel: a cluster sees the already-shrunk blocks of earlier neighbours). The same result is
// obtained by wavefronts of the dependency DAG (colours of the plan); clusters of one wavefront are mutually
// non-adjacent.
// ------------------------------------------------------------------------------------------------
void Tree::phase_sparsify(LevelLog& lg, SolveLevel& sl) {
    const double plan0 = wtime();
    const SymLevel& L = plan_.lv[ilvl_];
    const DevLevel& D = dplan_[ilvl_];
    const std::vector<int>& bottom = bottoms_[current_bottom_];
    const int first = bottom.empty() ? 0 : bottom.front();
    const int span = bottom.empty() ? 0 : bottom.back() - first + 1;
    lg.ignored = L.ignored;
    const int ncolors = L.ncolors;
    if (!L.q.empty()) {
        QrSrc* ds = scratch_->alloc_n<QrSrc>(L.qs.size());
        launch_expand_qsrc(tab_, D.qs, (int)L.qs.size(), ds, st_);
        lg.launches++;
        // per_color counts the tasks of the whole wavefront (all ranks): the launch shapes are the single-GPU ones
        std::vector<int> per_color(std::max(1, ncolors), 0);
        for (const SymQr& q : L.q) per_color[q.color]++;
        // ... except the cluster width G of the streaming shape, which does not change a task's arithmetic (same
        // columns, same 32-lane reduction trees): it is sized from this rank's share of the wavefront, so a rank that
        // holds a quarter of the tasks spreads each of them over four times as many CTAs
        std::vector<int> per_color_own(std::max(1, ncolors), 0);
        std::vector<const SymQr*> own;
        for (const SymQr& q : L.q)
            if (mine(q.cluster)) {
                own.push_back(&q);
                per_color_own[q.color]++;
            }
        const size_t nq = own.size();
        std::vector<QrTask> tasks(nq);
        std::vector<int> task_color(nq);
        size_t vtotal = 0, ttotal = 0;
        std::vector<size_t> voff(nq), toff(nq);
        // sizes of every task (sum over its ~10-30 neighbours: the bulk of the planning time on the low levels): host threads
        parallel_chunks(nq, [&](int, size_t b, size_t e) {
            for (size_t i = b; i < e; i++) {
                const SymQr& q = *own[i];
                QrTask& t = tasks[i];
                t.cluster = q.cluster;
                t.rows = h_csize_[q.cluster];
                t.src0 = q.src0;
                t.nsrc = q.nsrc;
                int maxcols = 0;
                for (int k = 0; k < q.nsrc; k++) maxcols += h_csize_[L.qs[q.src0 + k].nbr];
                t.maxcols = maxcols;
                t.W = nullptr;
            }
        }, 16384);
        for (size_t i = 0; i < nq; i++) {
            const SymQr& q = *own[i];
            QrTask& t = tasks[i];
            const int maxcols = t.maxcols;
            const size_t kmax = std::min(t.rows, maxcols);
            voff[i] = vtotal;
            toff[i] = ttotal;
            vtotal += ((size_t)t.rows * kmax + 31) & ~(size_t)31;
            ttotal += (kmax + 31) & ~(size_t)31;
            task_color[i] = q.color;
        }
        {
            double* vbase = arena_->alloc_n<double>(vtotal + 32);
            double* tbase = arena_->alloc_n<double>(ttotal + 32);
            for (size_t i = 0; i < nq; i++) {
                tasks[i].V = vbase + voff[i];
                tasks[i].tau = tbase + toff[i];
            }
        }
        // Launch classes: (threads, cluster size G, shared-memory bucket). The panel stays resident in (distributed)
        // shared memory whenever it fits in 16 CTAs; otherwise it lives in the (L2-resident) scratch arena.
        const int MAXS = rrqr_max_smem();
        static const int kBuckets[] = {12 << 10, 24 << 10, 40 << 10, 72 << 10, 112 << 10, 224 << 10};
        auto pow2floor = [](int x) { int p = 1; while (2 * p <= x) p *= 2; return p; };
        auto pow2ceil = [](int x) { int p = 1; while (p < x) p *= 2; return p; };
        // class = (mode << 8) | (log2 G << 4) | bucket; mode 0: 128 threads (G = 1), 1: 256 threads, 2: panel in the
        // global scratch (512 threads, G = 8), 3: 512 threads
        std::vector<int> klass(nq);
        std::vector<int> smem_need(nq);
        const int kNB = 6;
        // test hooks: force every task through one kernel shape (tests/test_gpu_parity.py)
        const bool force_global = getenv("SPAND_RRQR_FORCE_GLOBAL") != nullptr;
        const int force_g = getenv("SPAND_RRQR_FORCE_G") ? atoi(getenv("SPAND_RRQR_FORCE_G")) : 0;
        const char* mode_env = getenv("SPAND_RRQR_MODE");  // "smem": cluster kernels with the panel in shared memory
        const bool smem_mode = mode_env != nullptr && std::string(mode_env) == "smem";
        // hot / cold threshold of the global-panel kernel (rrqr.cu); 0 = sweep every column every step
        const double qr_theta = getenv("SPAND_RRQR_THETA") ? atof(getenv("SPAND_RRQR_THETA")) : 0.5;
        // hot-set kernel (rrqr_hc2.cu) for panels in global memory: on unless SPAND_RRQR_HC2=0; per-CTA shared memory
        // budget (two 256-thread CTAs per SM by default), CTAs wanted per wavefront, smallest panel it takes
        const bool hc2_on = !getenv("SPAND_RRQR_HC2") || atoi(getenv("SPAND_RRQR_HC2")) != 0;
        const int top_g = getenv("SPAND_RRQR_GTOP") ? atoi(getenv("SPAND_RRQR_GTOP")) : 0;
        const int gtop_switch = getenv("SPAND_RRQR_GTOP_SWITCH") ? atoi(getenv("SPAND_RRQR_GTOP_SWITCH")) : 8;
        const int cluster_ctas = getenv("SPAND_RRQR_WANT") ? atoi(getenv("SPAND_RRQR_WANT")) : 296;
        const int colk_maxrows = std::min(rrqr_col_max_rows(),
                                          getenv("SPAND_RRQR_COLROWS") ? atoi(getenv("SPAND_RRQR_COLROWS")) : 64);
        // panel of the column kernel in L2 instead of shared memory: 0 never (default: measured 2.7x slower at level 3 of
        // C4, thread-private column walks thrash the L1), 1 always, 2 when it does not fit shared memory
        const int colk_gp = getenv("SPAND_RRQR_COLGP") ? atoi(getenv("SPAND_RRQR_COLGP")) : 0;
        const bool colk_on = !getenv("SPAND_RRQR_COL") || atoi(getenv("SPAND_RRQR_COL")) != 0;
        const long colk_max = (getenv("SPAND_RRQR_COLKB") ? atol(getenv("SPAND_RRQR_COLKB")) : 222) * 1024;
        const bool force_hc2 = getenv("SPAND_RRQR_HC2") && atoi(getenv("SPAND_RRQR_HC2")) == 2;  // every eligible panel
        const int hc2_tmin = getenv("SPAND_HC2_TMIN") ? atoi(getenv("SPAND_HC2_TMIN")) : 48;
        const int hc2_maxrows = getenv("SPAND_HC2_MAXROWS") ? atoi(getenv("SPAND_HC2_MAXROWS")) : 256;
        const long hc2_budget = (getenv("SPAND_HC2_KB") ? atol(getenv("SPAND_HC2_KB")) : 105) * 1024;
        const long hc2_budget_big = (getenv("SPAND_HC2_BIGKB") ? atol(getenv("SPAND_HC2_BIGKB")) : 215) * 1024;
        const int hc2_ctas = getenv("SPAND_HC2_CTAS") ? atoi(getenv("SPAND_HC2_CTAS")) : 296;  // twice the CTAs wanted
        // hot set = columns within this factor of the largest norm (as far as the capacity goes)
        const double hc2_theta = getenv("SPAND_HC2_THETA") ? atof(getenv("SPAND_HC2_THETA")) : 0.5;
        const int hc2_hmax = getenv("SPAND_HC2_HMAX") ? atoi(getenv("SPAND_HC2_HMAX")) : 128;
        const double hc2_min_bytes = (getenv("SPAND_HC2_MINKB") ? atof(getenv("SPAND_HC2_MINKB")) : 300.0) * 1024.0;
        const int stream_tmin = getenv("SPAND_RRQR_TMIN") ? atoi(getenv("SPAND_RRQR_TMIN")) : 48;
        const int stream_ctas = getenv("SPAND_RRQR_CTAS") ? atoi(getenv("SPAND_RRQR_CTAS")) : 1776;
        const long smem1_max = (getenv("SPAND_RRQR_SMEM1KB") ? atol(getenv("SPAND_RRQR_SMEM1KB")) : 200) * 1024;
        const double l2_budget = (getenv("SPAND_RRQR_L2MB") ? atof(getenv("SPAND_RRQR_L2MB")) : 1.0e5) * 1048576.0;
        // panels smaller than this stay in (distributed) shared memory even in wide wavefronts
        const double stream_min_bytes = (getenv("SPAND_RRQR_MINKB") ? atof(getenv("SPAND_RRQR_MINKB")) : 300.0) * 1024.0;
        std::vector<double> color_work(std::max(1, ncolors), 0.0);  // rows x columns of this rank's tasks, per wavefront
        for (size_t i = 0; i < nq; i++) color_work[task_color[i]] += (double)tasks[i].rows * tasks[i].maxcols;
        for (size_t i = 0; i < nq; i++) {
            QrTask& t = tasks[i];
            int mn = std::max(1, std::min(t.rows, t.maxcols));
            t.nb = std::min(QR_NB, mn);
            auto config = [&](int NT, int G, bool in_smem) {
                int cpcm = std::max(1, (t.maxcols + G - 1) / G);
                int Ln = in_smem ? std::min(32, std::max(1, pow2floor(NT / cpcm))) : 32;
                Ln = std::min(Ln, pow2ceil(std::max(1, (t.rows + 1) / 2)));  // a lane works on pairs of rows
                int ld = (t.rows + 1) & ~1;
                if (in_smem && Ln < 8)
                    while (ld % 16 != (2 * Ln) % 16) ld += 2;  // conflict-free 128-bit shared loads
                t.L = Ln;
                t.ld = ld;
                t.in_smem = in_smem ? 1 : 0;
                return (long)rrqr_smem_bytes(t.rows, t.maxcols, G, t.nb, t.ld, in_smem);
            };
            auto bucket_of = [&](long nd) {
                int b = 0;
                while (b < kNB - 1 && nd > kBuckets[b]) b++;
                return b;
            };
            // Short panels (the lower levels: tens of thousands of panels of 9-47 rows) that fit the shared memory of
            // one CTA: one thread per column, no reductions over the rows (rrqr_col_kernel)
            if (colk_on && !force_global && force_g == 0 && !smem_mode && t.rows <= colk_maxrows && t.rows > 0 &&
                t.maxcols > 0) {
                const long cb = (long)rrqr_col_smem_bytes(t.rows, t.maxcols, t.nsrc, false);
                const bool gp = colk_gp == 1 || (colk_gp != 0 && cb > colk_max);  // panel in L2 when it does not fit
                if (gp || cb <= colk_max) {
                    const long need = gp ? (long)rrqr_col_smem_bytes(t.rows, t.maxcols, t.nsrc, true) : cb;
                    t.ld = rrqr_col_ld(t.rows);
                    t.L = 1;
                    t.nb = 1;
                    t.in_smem = gp ? 0 : 1;
                    if (gp) t.W = scratch_->alloc_n<double>((size_t)t.ld * t.maxcols);
                    klass[i] = ((gp ? 7 : 6) << 8) | ((t.maxcols <= 128 ? 0 : (t.maxcols <= 256 ? 1 : 2)) << 4) |
                               bucket_of(need);
                    smem_need[i] = (int)need;
                    continue;
                }
            }
            long nd = config(128, 1, true);
            if (!force_global && force_g == 0 && t.rows <= 64 && nd <= kBuckets[2]) {
                klass[i] = bucket_of(nd);
                smem_need[i] = (int)nd;
                continue;
            }
            // Hot-set kernel: the pivot search runs on a shared-memory copy of the few columns that can win it, the
            // rest of the panel (global scratch) is refreshed once per block on the tensor cores. The capacity of the
            // hot set and the smallest cluster width depend on the task alone (not on the wavefront or on how many GPUs
            // share it), so the arithmetic of a task is the same however the launch is shaped.
            const int rp = hc2_row_pairs(t.rows);
            // Measured on C4 (profiles/r2_rrqr.md): it wins on the wide wavefronts of the middle levels (hundreds of
            // panels of up to 256 rows, one CTA each); the few tall panels of the upper levels are still served faster
            // by the cluster kernel that keeps all columns current.
            const bool hc2_fits = force_hc2 || (per_color_own[task_color[i]] >= hc2_tmin && t.rows <= hc2_maxrows);
            if (hc2_on && hc2_fits && !smem_mode && !force_global && force_g == 0 && rp > 0 &&
                8.0 * t.rows * t.maxcols >= hc2_min_bytes) {
                const bool big = hc2_threads(t.rows) == 512;  // 512 threads, one CTA per SM; else 256 threads, two per SM
                const long budget = big ? hc2_budget_big : hc2_budget;
                const long ldv = (t.rows + 1) & ~1;
                auto hcap_for = [&](int G) {  // largest hot set that fits the budget with this cluster width
                    int h = (int)std::min<long>(hc2_hmax, std::max<long>(0, budget / (8 * ldv)));
                    while (h > 0 && (long)hc2_smem_bytes(t.rows, t.maxcols, G, h, t.nsrc) > budget) h--;
                    return h;
                };
                int gmin = 1;
                while (gmin < 16 && hcap_for(gmin) < 32) gmin *= 2;
                int hcap = std::min(std::min(hcap_for(gmin), hc2_hmax), t.maxcols);
                if (hcap >= 8 || hcap == t.maxcols) {
                    // cluster width: the wavefront should cover the SMs without exceeding them (one CTA per SM; a
                    // 16-CTA cluster needs a whole GPC, of which there are 8: keep 16 for wavefronts of up to 6 tasks)
                    // its share of the SMs is its share of the work of the wavefront (rows x columns: the big panels
                    // of a wavefront are its critical path), rounded down to a power of two
                    int G = gmin;
                    const int nown = std::max(1, per_color_own[task_color[i]]);
                    const double share = (double)t.rows * t.maxcols / std::max(1.0, color_work[task_color[i]]);
                    while (G < 16 && 2 * G <= share * (hc2_ctas / 2)) G *= 2;
                    if (G == 16 && nown > 6 && gmin <= 8) G = 8;
                    int g = 0;
                    while ((1 << g) < G) g++;
                    t.hcap = hcap;
                    t.nb = HC2_NB;
                    t.L = 32;
                    t.ld = (int)ldv;
                    t.in_smem = 0;
                    t.W = scratch_->alloc_n<double>((size_t)t.ld * t.maxcols);
                    t.X = G >= 8 ? scratch_->alloc_n<double>(hc2_exchange_doubles(t.rows, G)) : nullptr;
                    nd = (long)hc2_smem_bytes(t.rows, t.maxcols, G, hcap, t.nsrc);
                    klass[i] = (5 << 8) | (g << 4) | (rp <= 2 ? 0 : (rp <= 4 ? 1 : (rp <= 6 ? 2 : 3)));
                    smem_need[i] = (int)nd;
                    continue;
                }
            }
            bool stream = !smem_mode && !force_global && force_g == 0 && per_color[task_color[i]] >= stream_tmin;
            if (stream && config(256, 1, true) <= smem1_max) stream = false;  // fits one CTA's shared memory
            if (stream && 8.0 * t.rows * t.maxcols < stream_min_bytes) stream = false;
            if (stream) {
                // Streaming shape: the panel lives in the scratch arena and is read once per Householder step by
                // small CTAs (256 threads, 64 registers, several per SM); the cluster is as wide as needed for the
                // wavefront to put about four CTAs on every SM and for the per-CTA state to stay small.
                t.nb = std::min(QR_NBS, mn);
                int G = 1;
                while (G < 16 && per_color_own[task_color[i]] * G < stream_ctas) G *= 2;
                // keep the panels of the CTAs resident at the same time (about 592) inside the L2
                const double panel_bytes = 8.0 * t.rows * t.maxcols;
                while (G < 16 && ((double)stream_ctas / G) * panel_bytes > l2_budget) G *= 2;
                nd = config(256, G, false);
                while (G < 16 && nd > (56 << 10)) {
                    G *= 2;
                    nd = config(256, G, false);
                }
                while (nd > MAXS && t.nb > 2) {
                    t.nb /= 2;
                    nd = config(256, G, false);
                }
                if (nd > MAXS) throw std::runtime_error("sparsify: interface cluster too large for the RRQR kernel");
                int g = 0;
                while ((1 << g) < G) g++;
                t.W = scratch_->alloc_n<double>((size_t)t.ld * t.maxcols);
                klass[i] = (4 << 8) | (g << 4) | bucket_of(nd);
                smem_need[i] = (int)nd;
                continue;
            }
            // Cluster size: wide enough that the wavefront fills the GPU (about two CTAs per SM), at least what the
            // panel needs to stay in shared memory.
            int want = 1;
            while (want < 16 && per_color[task_color[i]] * want < cluster_ctas) want *= 2;
            if (force_g > 0) want = force_g;
            int g = 0;
            while ((1 << g) < want) g++;
            int gi = -1, nt = 256;
            for (; g <= 4 && gi < 0 && !force_global; g++) {
                nd = config(256, 1 << g, true);
                if (nd <= kBuckets[4]) gi = g;
                else if (nd <= MAXS) {
                    // 512 threads change the lanes per column and with them the padding of ld: check again
                    const long nd512 = config(512, 1 << g, true);
                    if (nd512 <= MAXS) {
                        nt = 512;
                        nd = nd512;
                        gi = g;
                    }
                }
            }
            if (gi >= 0) {
                klass[i] = ((nt == 256 ? 1 : 3) << 8) | (gi << 4) | bucket_of(nd);
                smem_need[i] = (int)nd;
                continue;
            }
            // panel stays in the scratch arena: a cluster of 512-thread CTAs streams its slabs from L2. 16 CTAs for
            // the narrowest wavefronts; a 16-CTA cluster needs a whole GPC (8 of them), so wavefronts of more than
            // gtop_switch panels use 8 (two clusters per GPC, shorter cluster barriers)
            // (the hot/cold kernel keeps 20 KB of static shared memory: per-warp Y tiles, T, records)
            const long MAXS_HC = MAXS - 24 * 1024;
            int gtop = top_g > 0 ? top_g : (per_color_own[task_color[i]] > gtop_switch ? 8 : 16);
            nd = config(512, gtop, false);
            while (nd > MAXS_HC && gtop < 16) {
                gtop *= 2;
                nd = config(512, gtop, false);
            }
            while (nd > MAXS_HC && t.nb > 2) {
                t.nb /= 2;
                nd = config(512, gtop, false);
            }
            if (nd > MAXS_HC) throw std::runtime_error("sparsify: interface cluster too large for the RRQR kernel");
            t.W = scratch_->alloc_n<double>((size_t)t.ld * t.maxcols);
            int glog = 0;
            while ((1 << glog) < gtop) glog++;
            klass[i] = (2 << 8) | (glog << 4) | bucket_of(nd);
            smem_need[i] = (int)nd;
        }
        // order tasks by (colour, class), stable
        // (a stable counting sort: the keys are small integers, 10^5 tasks on the low levels)
        std::vector<int> idx(nq);
        {
            int kmax = 0;
            for (size_t i = 0; i < nq; i++) kmax = std::max(kmax, klass[i]);
            const size_t nk = (size_t)kmax + 1;
            std::vector<size_t> start((size_t)std::max(1, ncolors) * nk + 1, 0);
            for (size_t i = 0; i < nq; i++) start[(size_t)task_color[i] * nk + klass[i] + 1]++;
            for (size_t k = 1; k < start.size(); k++) start[k] += start[k - 1];
            for (size_t i = 0; i < nq; i++) idx[start[(size_t)task_color[i] * nk + klass[i]]++] = (int)i;
        }
        std::vector<QrTask> sorted(nq);
        for (size_t i = 0; i < nq; i++) sorted[i] = tasks[idx[i]];
        QrTask* dt = to_device(sorted, scratch_);
        // One launch per (colour, class). The classes of a colour are independent: they run concurrently on side
        // streams forked from / joined into the factorization stream. With several GPUs every colour ends with the
        // exchange of the new ranks (min over the replicas of csize) between two peer barriers: the next colour reads
        // the shrunk blocks and sizes of its earlier neighbours wherever they live.
        mg_barrier();
        size_t b = 0;
        for (int color = 0; color < ncolors; color++) {
            size_t cend = b;
            while (cend < nq && task_color[idx[cend]] == color) cend++;
            if (cend > b) {
                auto ev = fam_begin(F_RRQR);
                CK(cudaEventRecord(ev_fork_, st_));
                int nside = 0;
                while (b < cend) {
                    size_t e = b;
                    int smem = 0;
                    while (e < cend && klass[idx[e]] == klass[idx[b]]) {
                        smem = std::max(smem, smem_need[idx[e]]);
                        e++;
                    }
                    int k = klass[idx[b]];
                    int mode = k >> 8, G = 1 << ((k >> 4) & 15);
                    smem = (smem + 1023) & ~1023;
                    cudaStream_t s = side_[nside % kSide];
                    CK(cudaStreamWaitEvent(s, ev_fork_, 0));
                    const bool trace = getenv("SPAND_QR_TRACE") != nullptr;  // debug: one line per launch, serialised
                    cudaEvent_t tr0 = nullptr, tr1 = nullptr;
                    if (trace) {
                        CK(cudaDeviceSynchronize());
                        CK(cudaEventCreate(&tr0));
                        CK(cudaEventCreate(&tr1));
                        CK(cudaEventRecord(tr0, s));
                    }
                    if (mode == 6 || mode == 7) {
                        launch_rrqr_col(dt + b, (int)(e - b), ds, d_csize_, tol, 128 << ((k >> 4) & 15), smem, mode == 7, s);
                    } else if (mode == 5) {
                        static const int kRowPairs[4] = {2, 4, 6, 10};
                        launch_rrqr_hc2(dt + b, (int)(e - b), ds, d_csize_, tol, G, kRowPairs[k & 15], smem, hc2_theta, s);
                    } else
                    launch_rrqr(dt + b, (int)(e - b), ds, d_csize_, tol, G,
                                mode == 0 ? 128 : ((mode == 1 || mode == 4) ? 256 : 512), mode != 2 && mode != 4, smem, s,
                                qr_theta);
                    if (trace) {
                        CK(cudaEventRecord(tr1, s));
                        CK(cudaEventSynchronize(tr1));
                        float ms = 0;
                        CK(cudaEventElapsedTime(&ms, tr0, tr1));
                        int mr = 0, mc = 0, mh = 0;
                        long sr = 0, sc = 0;
                        for (size_t q = b; q < e; q++) {
                            const QrTask& tq = sorted[q];
                            mr = std::max(mr, tq.rows);
                            mc = std::max(mc, tq.maxcols);
                            mh = std::max(mh, tq.hcap);
                            sr += tq.rows;
                            sc += tq.maxcols;
                        }
                        fprintf(stderr, "QRTRACE lvl %d color %d mode %d G %d tasks %d smem %d rows max %d avg %ld cols max %d avg %ld hcap %d: %.3f ms\n",
                                ilvl_, color, mode, G, (int)(e - b), smem, mr, sr / (long)(e - b), mc, sc / (long)(e - b), mh, ms);
                        cudaEventDestroy(tr0);
                        cudaEventDestroy(tr1);
                    }
                    nside++;
                    lg.launches++;
                    family_launches[F_RRQR]++;
                    b = e;
                }
                for (int i = 0; i < std::min(nside, kSide); i++) {
                    CK(cudaEventRecord(ev_join_[i], side_[i]));
                    CK(cudaStreamWaitEvent(st_, ev_join_[i], 0));
                }
                family_launches[F_RRQR]--;  // fam_begin counted the colour once
                fam_end(F_RRQR, ev);
            }
            if (mg()) {
                mg_barrier();
                launch_csize_min(mg_csize_, mg_rank, mg_nranks, first, span, st_);
                mg_barrier();
                lg.launches += 3;
            }
        }
        lg.wavefronts = ncolors;
        // Orthogonal ops (tree.cpp:1322-1331): one entry per task, no-ops where the rank did not drop
        sl.n_house = (int)nq;
        sl.house = arena_->alloc_n<HouseTask>(std::max<size_t>(1, nq));
        launch_expand_house(tab_, dt, (int)nq, sl.house, st_);
        lg.launches++;
        lg.t_plan_spars = wtime() - plan0;
        // ranks back to the host: the one synchronisation of the level
        CK(cudaMemcpyAsync(h_csize_.data() + first, d_csize_ + first, sizeof(int) * span, cudaMemcpyDeviceToHost, st_));
        check_error();
        if (const char* fn = getenv("SPAND_DUMP_QR")) {
            if (FILE* dump = fopen(fn, "a")) {
                for (size_t i = 0; i < nq; i++)
                    fprintf(dump, "%d %d %d %d %d %d\n", ilvl_, tasks[i].cluster, tasks[i].rows, tasks[i].maxcols,
                            h_csize_[tasks[i].cluster], task_color[i]);
                fclose(dump);
            }
        }
        for (const SymQr& q : L.q) cl_[q.cluster].size = h_csize_[q.cluster];
    } else {
        check_error();
    }
    CK(cudaStreamSynchronize(st_));
    stager_.reset();
    scratch_->reset();
}

// ------------------------------------------------------------------------------------------------
// MERGE — src/tree.cpp:1435-1445 (reset_size :1106-1131, update_edges :1133-1184)
// ------------------------------------------------------------------------------------------------
void Tree::phase_merge(LevelLog& lg, SolveLevel& sl) {
    const SymLevel& L = plan_.lv[ilvl_];
    const DevLevel& D = dplan_[ilvl_];
    const std::vector<int>& children = bottoms_[current_bottom_];
    current_bottom_++;
    const std::vector<int>& parents = bottoms_[current_bottom_];
    if (parents.empty()) return;
    size_t xtotal = 0;
    int maxchild = 0;
    for (int p : parents) {
        Cluster& cp = cl_[p];
        int size = 0;
        for (int c = cp.child_begin; c < cp.child_end; c++) {
            h_pos_[c] = size;
            size += h_csize_[c];
            maxchild = std::max(maxchild, h_csize_[c]);
        }
        cp.size = cp.orig_size = size;
        h_csize_[p] = size;
        xtotal += size;
    }
    if (!mg()) {
        double* xbase = arena_->alloc_n<double>(xtotal + 1);
        size_t off = 0;
        for (int p : parents) {
            h_xptr_[p] = xbase + off;
            off += cl_[p].size;
        }
    } else {
        for (int p : parents) h_xptr_[p] = alloc_block(h_owner_[p], cl_[p].size);
    }
    const int pfirst = parents.front(), np = parents.back() - pfirst + 1;
    stager_.upload(d_csize_ + pfirst, h_csize_.data() + pfirst, sizeof(int) * np, st_);
    stager_.upload(d_xptr_ + pfirst, h_xptr_.data() + pfirst, sizeof(double*) * np, st_);
    if (!children.empty()) {
        const int cfirst = children.front(), nc = children.back() - cfirst + 1;
        stager_.upload(d_pos_ + cfirst, h_pos_.data() + cfirst, sizeof(int) * nc, st_);
    }
    alloc_edges(L.medge0, L.medge1, true, lg);
    const bool ident = ilvl_ >= skip;  // children were scaled at this level: their pivots are I (tree.cpp:811)
    auto evc = fam_begin(F_COPY);
    launch_copy_sym(tab_, D.m_copy, (int)L.m_copy.size(), ident ? 1 : 0, st_);
    fam_end(F_COPY, evc);
    lg.launches++;
    if ((long)maxchild * maxchild > COPY_SMALL) {
        std::vector<CopyTask> copies;
        const int CHUNK = 4096;
        for (const SymCopy& t : L.m_copy) {
            const int rows = h_csize_[t.c2], cols = h_csize_[t.c1];
            if ((long)rows * cols <= COPY_SMALL || (ident && t.c1 == t.c2) || !mine(t.c1)) continue;
            const int ldn = h_eld_[t.enew], lds = h_eld_[t.eold];
            double* dst = h_eptr_[t.enew] + h_pos_[t.c2] + (size_t)h_pos_[t.c1] * ldn;
            const double* src = h_eptr_[t.eold];
            int cstep = std::max(1, CHUNK / rows);
            for (int c0 = 0; c0 < cols; c0 += cstep) {
                int w = std::min(cstep, cols - c0);
                copies.push_back({src + (size_t)c0 * lds, dst + (size_t)c0 * ldn, lds, ldn, rows, w});
            }
        }
        if (!copies.empty()) {
            CopyTask* dc = to_device(copies, scratch_);
            evc = fam_begin(F_COPY);
            launch_copy(dc, (int)copies.size(), st_);
            fam_end(F_COPY, evc);
            lg.launches++;
        }
    }
    sl.n_merge = D.n_children;
    sl.m_fwd = arena_->alloc_n<XCopyTask>(std::max(1, D.n_children));
    sl.m_bwd = arena_->alloc_n<XCopyTask>(std::max(1, D.n_children));
    launch_expand_xcopy(tab_, D.children, D.n_children, sl.m_fwd, sl.m_bwd, st_);
    lg.launches++;
}

// ------------------------------------------------------------------------------------------------
// FACTORIZE — src/tree.cpp:1447-1551
// ------------------------------------------------------------------------------------------------
void Tree::factorize() {
    DeviceGuard dev_guard(device);
    if (symm_kind == SPD && scale_kind != LLT) throw std::runtime_error("SPD requires LLT scaling");
    if (symm_kind == GEN && scale_kind != PLU) throw std::runtime_error("GEN requires PLU scaling (PLUQ is out of scope)");
    if (symm_kind == SYM) throw std::runtime_error("SYM/LDLT is out of scope (SURVEY.md section 2)");
    if (!assembled_ || factorized_ || state_level_ >= 0) throw std::runtime_error("factorize: call assemble first");
    if (plan_.symmetric != symmetry() || plan_.want_flag != use_want_sparsify)
        throw std::runtime_error("factorize: symm_kind / use_sparsify changed after assemble");
    ensure_device();
    CK(cudaMemsetAsync(d_err_, 0, sizeof(int), st_));
    for (int f = 0; f < F_COUNT; f++) {
        family_ms[f] = 0;
        family_launches[f] = 0;
    }
    struct Events {  // destroyed on every exit path (a non-SPD pivot throws out of the level loop)
        std::vector<cudaEvent_t> v;
        ~Events() {
            for (auto e : v)
                if (e) cudaEventDestroy(e);
        }
    } evs;
    evs.v.assign(nlevels * 5 + 2, nullptr);
    for (auto& e : evs.v) CK(cudaEventCreate(&e));
    cudaEvent_t* ev = evs.v.data();
    cudaEvent_t ev_begin = evs.v[nlevels * 5], ev_end = evs.v[nlevels * 5 + 1];
    CK(cudaEventRecord(ev_begin, st_));
    bool stopped = false;
    int last_level = -1;
    logs_final_ = false;
    mg_barrier();  // every rank has assembled
    auto snapshot = [&](std::vector<int>& dst) {
        const std::vector<int>& bottom = bottoms_[current_bottom_];
        if (bottom.empty()) {
            dst.clear();
            return;
        }
        dst.assign(h_csize_.begin() + bottom.front(), h_csize_.begin() + bottom.back() + 1);
    };
    for (ilvl_ = 0; ilvl_ < nlevels && !stopped; ilvl_++) {
        LevelLog& lg = log[ilvl_];
        SolveLevel& sl = solve_[ilvl_];
        last_level = ilvl_;
        double h0 = wtime();
        if (verb) printf("Level %d, %d dofs left\n", ilvl_, ndofs_left());
        cnt_next_ = 0;
        CK(cudaMemsetAsync(d_cnt_, 0, sizeof(int) * 64, st_));
        snapshot(size_pre_[ilvl_]);
        CK(cudaEventRecord(ev[ilvl_ * 5 + 0], st_));
        double p0 = wtime();
        phase_eliminate(lg, sl);
        phases_done_[ilvl_] |= 1;
        state_level_ = ilvl_;
        state_phase_ = 0;
        lg.t_plan_elim = wtime() - p0;
        CK(cudaEventRecord(ev[ilvl_ * 5 + 1], st_));
        lg.dofs_left_elim = ndofs_left();
        if (ilvl_ == stop_level && stop_phase == 0) stopped = true;
        if (!stopped && ilvl_ >= skip) {
            p0 = wtime();
            phase_scale(lg, sl);
            phases_done_[ilvl_] |= 2;
            state_phase_ = 1;
            lg.t_plan_scale = wtime() - p0;
            CK(cudaEventRecord(ev[ilvl_ * 5 + 2], st_));
            if (ilvl_ == stop_level && stop_phase == 1) stopped = true;
            if (!stopped) {
                phase_sparsify(lg, sl);
                phases_done_[ilvl_] |= 4;
                state_phase_ = 2;
                if (ilvl_ == stop_level && stop_phase == 2) stopped = true;
            }
        } else {
            CK(cudaEventRecord(ev[ilvl_ * 5 + 2], st_));
            check_error();
            stager_.reset();
            scratch_->reset();
        }
        snapshot(size_post_[ilvl_]);
        CK(cudaEventRecord(ev[ilvl_ * 5 + 3], st_));
        p0 = wtime();
        if (!stopped && ilvl_ < nlevels - 1) {
            phase_merge(lg, sl);
            phases_done_[ilvl_] |= 8;
            state_phase_ = 3;
        }
        lg.t_plan_merge = wtime() - p0;
        CK(cudaEventRecord(ev[ilvl_ * 5 + 4], st_));
        lg.dofs_left_spars = ndofs_left();
        lg.t_host = wtime() - h0;
        if (ilvl_ == stop_level && stop_phase == 3) stopped = true;
    }
    CK(cudaEventRecord(ev_end, st_));
    CK(cudaStreamSynchronize(st_));
    check_error();
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ev_begin, ev_end));
    t_factorize_device = ms * 1e-3;
    fam_resolve();
    for (int l = 0; l <= last_level; l++) {
        float a, b, c, d;
        CK(cudaEventElapsedTime(&a, ev[l * 5 + 0], ev[l * 5 + 1]));
        CK(cudaEventElapsedTime(&b, ev[l * 5 + 1], ev[l * 5 + 2]));
        CK(cudaEventElapsedTime(&c, ev[l * 5 + 2], ev[l * 5 + 3]));
        CK(cudaEventElapsedTime(&d, ev[l * 5 + 3], ev[l * 5 + 4]));
        log[l].t_elim = a * 1e-3;
        log[l].t_scale = b * 1e-3;
        log[l].t_spars = c * 1e-3;
        log[l].t_merge = d * 1e-3;
        launches_total += log[l].launches;
        if (verb)
            printf("  lvl %d: elim %.2e scale %.2e spars %.2e merge %.2e host %.2e | left %d -> %d, %d launches, %d waves\n",
                   l, log[l].t_elim, log[l].t_scale, log[l].t_spars, log[l].t_merge, log[l].t_host,
                   log[l].dofs_left_elim, log[l].dofs_left_spars, log[l].launches, log[l].wavefronts);
    }
    stager_.reset();
    scratch_->reset();
    if (getenv("SPAND_TIMING"))
        fprintf(stderr, "[spand] factorize: %.1f ms on the device; allocations of this process so far: %ld calls, %.2f GB, %.1f ms\n",
                t_factorize_device * 1e3, g_malloc_calls, g_malloc_bytes / 1e9, g_malloc_seconds * 1e3);
    factorized_ = !stopped;
}

// Flop / byte / nnz model of SURVEY.md 8(d) and the reference's Log (util.h:366-401), evaluated after the fact from
// the plan and the cluster sizes recorded before the elimination and after the sparsification of every level.
void Tree::finalize_logs() {
    if (logs_final_) return;
    logs_final_ = true;
    const bool plu = scale_kind == PLU;
    nnz_ = 0;
    flop_log_.clear();
    for (int l = 0; l < nlevels; l++) {
        if (!phases_done_[l]) break;
        LevelLog& lg = log[l];
        const SymLevel& L = plan_.lv[l];
        const std::vector<int>& bottom = bottoms_[l];
        if (bottom.empty()) continue;
        const int first = bottom.front();
        const std::vector<int>& pre = size_pre_[l];
        const std::vector<int>& post = size_post_[l];
        auto P = [&](int c) { return (double)pre[c - first]; };
        auto Q = [&](int c) { return (double)post[c - first]; };
        auto tuple = [&](int kind, double rows, double cols, double inner) {
            if (monitor_flops) flop_log_.push_back({(long long)l, (long long)kind, (long long)rows, (long long)cols, (long long)inner});
        };
        lg.fl_pivot = lg.fl_panel = lg.fl_schur = lg.fl_rrqr_rank = lg.fl_rrqr_full = 0;
        lg.by_scale = lg.by_rrqr = lg.by_merge = 0;
        lg.rank_before = lg.rank_after = 0;
        lg.nspars = 0;
        lg.nbrs = 0;
        long long nz = 0;
        // eliminate
        for (int s : L.E) {
            const double n = P(s);
            tuple(0, n, 0, 0);
            lg.fl_pivot += (plu ? 2.0 : 1.0) * n * n * n / 3.0;
            nz += plu ? (long long)(n * n + 2 * n) : (long long)(n * (n + 1) / 2);
        }
        for (const SymTrsm& t : L.e_out) {
            tuple(1, P(t.cm), P(t.cn), 0);
            lg.fl_panel += P(t.cm) * P(t.cn) * P(t.cn);
            nz += (long long)(P(t.cm) * P(t.cn));
        }
        for (const SymTrsm& t : L.e_in) {
            tuple(1, P(t.cm), P(t.cn), 0);
            lg.fl_panel += P(t.cm) * P(t.cn) * P(t.cn);
            nz += (long long)(P(t.cm) * P(t.cn));
        }
        for (const SymGemm& g : L.e_gemm) {
            const double m = P(plan_.en2[g.target]), n = P(plan_.en1[g.target]);
            for (int ci = 0; ci < g.nc; ci++) {
                const SymCon& c = L.e_con[g.c0 + ci];
                const double k = P(plan_.en1[c.e1]);
                tuple(2, m, n, k);
                if (plan_.symmetric && plan_.en2[c.e1] == plan_.en2[c.e2]) lg.fl_schur += m * (m + 1) * k;
                else lg.fl_schur += 2.0 * m * n * k;
            }
        }
        // scale
        if (phases_done_[l] & 2) {
            for (int c : L.S) {
                const double n = P(c);
                tuple(0, n, 0, 0);
                lg.fl_pivot += (plu ? 2.0 : 1.0) * n * n * n / 3.0;
                nz += plu ? (long long)(n * n + 2 * n) : (long long)(n * (n + 1) / 2);
                lg.by_scale += 16.0 * n * n;
            }
            for (const SymTrsm& t : L.s_right) {
                const double rows = P(t.cm), cols = P(t.cn);
                tuple(1, rows, cols, 0);  // the two triangular solves of a scaled block (trsm_potf_edgeOut / edgeIn)
                tuple(1, cols, rows, 0);
                lg.fl_panel += rows * cols * cols + cols * rows * rows;
                lg.by_scale += 16.0 * rows * cols;
            }
        }
        // sparsify
        if (phases_done_[l] & 4) {
            std::vector<char> is_task(pre.size(), 0);
            for (const SymQr& q : L.q) is_task[q.cluster - first] = 1;
            for (const SymQr& q : L.q) {
                const double r = P(q.cluster), rk = Q(q.cluster);
                double cc = 0;  // columns seen: earlier sparsified neighbours had already shrunk (tree.cpp:1194-1200)
                for (int k = 0; k < q.nsrc; k++) {
                    const int nbr = L.qs[q.src0 + k].nbr;
                    cc += (is_task[nbr - first] && nbr < q.cluster) ? Q(nbr) : P(nbr);
                }
                double rf = std::min(r, cc);
                if (tol >= 1.0 || cc == 0) rf = 0;
                tuple(3, r, cc, 0);  // pushed after every geqp3 call, empty panels included (tree.cpp:1312)
                lg.rank_before += (long long)r;
                lg.nspars++;
                lg.nbrs += (long long)cc;
                lg.fl_rrqr_full += 4 * r * cc * rf - 2 * (r + cc) * rf * rf + (4.0 / 3.0) * rf * rf * rf;
                lg.fl_rrqr_rank += 4 * r * cc * rk - 2 * (r + cc) * rk * rk + (4.0 / 3.0) * rk * rk * rk;
                lg.by_rrqr += 8 * r * cc + 8 * rk * cc + 8 * r * rk;
                if (rk < r) {
                    // the reference eliminates the dropped sibling right away (tree.cpp:1342 -> potf_cluster :592 pushes a
                    // pivot tuple for its identity pivot); nothing is computed for it here, the tuple is listed only
                    tuple(0, r - rk, 0, 0);
                    nz += (long long)(r * r);  // Orthogonal (operations.cpp:159-161)
                    const double m = r - rk;
                    // Scaling op of the dropped sibling (tree.cpp:1342): ScalingLLT(I) or ScalingPLUQ(I, I, id, id)
                    nz += plu ? (long long)(m * m + 2 * m) : (long long)(m * (m + 1) / 2);
                }
                lg.rank_after += (long long)rk;
            }
        }
        // merge
        if (phases_done_[l] & 8) {
            for (const SymCopy& t : L.m_copy) lg.by_merge += 8.0 * Q(t.c1) * Q(t.c2);
            // parent blocks (zero fill included): sizes of the parents = sums of the children's sizes
            const std::vector<int>& parents = bottoms_[l + 1];
            std::vector<double> psize(parents.empty() ? 0 : parents.back() - parents.front() + 1, 0.0);
            for (int p : parents)
                for (int c = cl_[p].child_begin; c < cl_[p].child_end; c++) psize[p - parents.front()] += Q(c);
            for (int e = L.medge0; e < L.medge1; e++)
                lg.by_merge += 8.0 * psize[plan_.en1[e] - parents.front()] * psize[plan_.en2[e] - parents.front()];
        }
        nnz_ += nz;
        lg.fact_nnz = nnz_;
    }
}

const std::vector<LevelLog>& Tree::logs() {
    finalize_logs();
    return log;
}

long long Tree::nnz() {
    finalize_logs();
    return nnz_;
}

int Tree::get_stop() const {
    int stop = N;
    for (auto& l : log) {
        if (l.dofs_left_elim > 0) stop = std::min(stop, l.dofs_left_elim);
        if (l.dofs_left_spars > 0) stop = std::min(stop, l.dofs_left_spars);
    }
    return stop;
}

// ------------------------------------------------------------------------------------------------
// SOLVE — src/tree.cpp:1610-1635: all fwd() in record order, then all bwd() in reverse
// ------------------------------------------------------------------------------------------------
void Tree::solve_device(double* x_dev) {
    DeviceGuard dev_guard(device);
    if (!factorized_) throw std::runtime_error("solve: call factorize first");
    double* xleaf = d_xleaf_;
    launch_gather(N, d_perm_, x_dev, xleaf, st_);  // b = P^T x
    // Several GPUs: every operation runs on the owner of the segment it writes; the two Gemm* sweeps read segments
    // (and panels) of other ranks through NVLink after a peer barrier.
    for (int l = 0; l < nlevels; l++) {
        SolveLevel& s = solve_[l];
        launch_trsv(s.e_trsv, s.n_e_trsv, 0, s.max_e, st_);
        mg_barrier();
        launch_gemv(s.e_gemv_f, s.n_e_gemv_f, s.e_gemv_fc, 0, s.max_e, st_);
        launch_trsv(s.s_trsv, s.n_s_trsv, 0, s.max_s, st_);
        launch_house(s.house, s.n_house, 1, s.max_s, st_);
        launch_xcopy(s.m_fwd, s.n_merge, st_);
    }
    for (int l = nlevels - 1; l >= 0; l--) {
        SolveLevel& s = solve_[l];
        launch_xcopy(s.m_bwd, s.n_merge, st_);
        launch_house(s.house, s.n_house, 0, s.max_s, st_);
        // LLT: x <- L^-T x, x_s -= A[n,s]^T x_n ; PLU: x <- U^-1 x, x_s -= A[s,n] x_n   (operations.cpp bwd)
        const bool plu = scale_kind == PLU;
        launch_trsv(s.s_trsv, s.n_s_trsv, plu ? 2 : 1, s.max_s, st_);
        mg_barrier();
        launch_gemv(s.e_gemv_b, s.n_e_gemv_b, s.e_gemv_bc, plu ? 0 : 1, s.max_e, st_);
        launch_trsv(s.e_trsv, s.n_e_trsv, plu ? 2 : 1, s.max_e, st_);
    }
    if (!mg()) {
        launch_scatter(N, d_perm_, xleaf, x_dev, st_);  // x = P b
    } else {
        mg_barrier();
        launch_scatter_owned(N, d_perm_, mg_leaf_, d_dof_owner_, x_dev, st_);
        mg_barrier();  // nobody starts the next solve before all peers have read these segments
    }
}

void Tree::solve(double* x_host) {
    DeviceGuard dev_guard(device);
    if (!factorized_) throw std::runtime_error("solve: call factorize first");
    CK(cudaMemcpyAsync(d_xnat_, x_host, sizeof(double) * N, cudaMemcpyHostToDevice, st_));
    solve_device(d_xnat_);
    CK(cudaMemcpyAsync(x_host, d_xnat_, sizeof(double) * N, cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));
}

// src/is.cpp:39-121 with every vector resident in HBM
int Tree::cg(const SpMat& A, const double* rhs, double* x, int iters, double tol_, bool verbose, double* seconds) {
    DeviceGuard dev_guard(device);
    if (!factorized_) throw std::runtime_error("cg: call factorize first");
    if (A.rows != N || A.cols != N) throw std::runtime_error("cg: the matrix does not have the size of the factorized one");
    int n = A.cols;
    SpMat At = transpose(A);  // CSR of A = CSC of A^T
    int *d_rp, *d_ci;
    double *d_v, *d_x, *d_r, *d_p, *d_z, *d_tmp, *d_b, *d_s, *d_part;
    CK(cudaMalloc((void**)&d_rp, sizeof(int) * (n + 1)));
    CK(cudaMalloc((void**)&d_ci, sizeof(int) * At.nnz()));
    CK(cudaMalloc((void**)&d_v, sizeof(double) * At.nnz()));
    double** vecs[] = {&d_x, &d_r, &d_p, &d_z, &d_tmp, &d_b};
    for (auto v : vecs) CK(cudaMalloc((void**)v, sizeof(double) * n));
    CK(cudaMalloc((void**)&d_s, sizeof(double) * 4));
    CK(cudaMalloc((void**)&d_part, sizeof(double) * (KR_GRID + 1)));
    CK(cudaMemcpyAsync(d_rp, At.colptr.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice, st_));
    CK(cudaMemcpyAsync(d_ci, At.rowind.data(), sizeof(int) * At.nnz(), cudaMemcpyHostToDevice, st_));
    CK(cudaMemcpyAsync(d_v, At.val.data(), sizeof(double) * At.nnz(), cudaMemcpyHostToDevice, st_));
    CK(cudaMemcpyAsync(d_b, rhs, sizeof(double) * n, cudaMemcpyHostToDevice, st_));
    CK(cudaMemcpyAsync(d_x, x, sizeof(double) * n, cudaMemcpyHostToDevice, st_));
    auto dot = [&](const double* a, const double* b) {
        double h;  // fixed-order reduction: the whole PCG is bit-reproducible
        launch_dot_det(n, a, b, d_part, d_s, st_);
        CK(cudaMemcpyAsync(&h, d_s, sizeof(double), cudaMemcpyDeviceToHost, st_));
        CK(cudaStreamSynchronize(st_));
        return h;
    };
    CK(cudaStreamSynchronize(st_));
    double t0 = wtime();
    int result;
    {
        launch_spmv(n, d_rp, d_ci, d_v, d_x, d_tmp, st_);
        CK(cudaMemcpyAsync(d_r, d_b, sizeof(double) * n, cudaMemcpyDeviceToDevice, st_));
        launch_axpy(n, -1.0, d_tmp, d_r, st_);
        double rhsNorm2 = dot(d_b, d_b);
        if (rhsNorm2 == 0) {
            CK(cudaMemsetAsync(d_x, 0, sizeof(double) * n, st_));
            result = 0;
        } else {
            double threshold = tol_ * tol_ * rhsNorm2;
            double residualNorm2 = dot(d_r, d_r);
            if (residualNorm2 < threshold) {
                result = 0;
            } else {
                CK(cudaMemcpyAsync(d_p, d_r, sizeof(double) * n, cudaMemcpyDeviceToDevice, st_));
                solve_device(d_p);
                double absNew = dot(d_r, d_p);
                int i = 0;
                while (i < iters) {
                    launch_spmv(n, d_rp, d_ci, d_v, d_p, d_tmp, st_);
                    double alpha = absNew / dot(d_p, d_tmp);
                    launch_axpy(n, alpha, d_p, d_x, st_);
                    launch_axpy(n, -alpha, d_tmp, d_r, st_);
                    residualNorm2 = dot(d_r, d_r);
                    if (verbose) printf("%d: |Ax-b|/|b| = %3.2e <? %3.2e\n", i, std::sqrt(residualNorm2 / rhsNorm2), tol_);
                    if (residualNorm2 < threshold) break;
                    CK(cudaMemcpyAsync(d_z, d_r, sizeof(double) * n, cudaMemcpyDeviceToDevice, st_));
                    solve_device(d_z);
                    double absOld = absNew;
                    absNew = dot(d_r, d_z);
                    double beta = absNew / absOld;
                    launch_xpay(n, d_z, beta, d_p, st_);
                    i++;
                }
                result = i + 1;
            }
        }
    }
    CK(cudaMemcpyAsync(x, d_x, sizeof(double) * n, cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));
    if (seconds) *seconds = wtime() - t0;
    cudaFree(d_rp);
    cudaFree(d_ci);
    cudaFree(d_v);
    for (auto v : vecs) cudaFree(*v);
    cudaFree(d_s);
    cudaFree(d_part);
    return result;
}

// src/is.cpp:123-300: Householder GMRES, left preconditioned, restarted. The Krylov basis is kept as the
// essential parts of the reflectors in an N x (restart + 1) array in HBM (column i holds H_i below row i); the
// small triangular factor, the Givens rotations and the rotated right-hand side w live on the host, which reads
// back the k + 1 leading entries of the new column once per iteration (the only synchronisation).
int Tree::gmres(const SpMat& A, const double* rhs, double* x, int iters, int restart, double tol_, bool verbose,
                double* seconds) {
    DeviceGuard dev_guard(device);
    if (!factorized_) throw std::runtime_error("gmres: call factorize first");
    if (A.rows != N || A.cols != N) throw std::runtime_error("gmres: the matrix does not have the size of the factorized one");
    const int m = A.cols;
    if (restart < 1) throw std::runtime_error("gmres: restart must be >= 1");
    if (restart > m) restart = m;
    SpMat At = transpose(A);  // CSR of A = CSC of A^T
    int *d_rp, *d_ci;
    double *d_v, *d_x, *d_b, *d_w, *d_t, *d_xn, *d_H, *d_tau, *d_beta, *d_part;
    CK(cudaMalloc((void**)&d_rp, sizeof(int) * (m + 1)));
    CK(cudaMalloc((void**)&d_ci, sizeof(int) * At.nnz()));
    CK(cudaMalloc((void**)&d_v, sizeof(double) * At.nnz()));
    double** vecs[] = {&d_x, &d_b, &d_w, &d_t, &d_xn};
    for (auto v : vecs) CK(cudaMalloc((void**)v, sizeof(double) * m));
    CK(cudaMalloc((void**)&d_H, sizeof(double) * (size_t)m * (restart + 1)));
    CK(cudaMalloc((void**)&d_tau, sizeof(double) * (restart + 1)));
    CK(cudaMalloc((void**)&d_beta, sizeof(double)));
    CK(cudaMalloc((void**)&d_part, sizeof(double) * (KR_GRID + 1)));
    CK(cudaMemcpyAsync(d_rp, At.colptr.data(), sizeof(int) * (m + 1), cudaMemcpyHostToDevice, st_));
    CK(cudaMemcpyAsync(d_ci, At.rowind.data(), sizeof(int) * At.nnz(), cudaMemcpyHostToDevice, st_));
    CK(cudaMemcpyAsync(d_v, At.val.data(), sizeof(double) * At.nnz(), cudaMemcpyHostToDevice, st_));
    CK(cudaMemcpyAsync(d_b, rhs, sizeof(double) * m, cudaMemcpyHostToDevice, st_));
    CK(cudaMemcpyAsync(d_x, x, sizeof(double) * m, cudaMemcpyHostToDevice, st_));
    auto Hess = [&](int i) { return d_H + (size_t)i * m + i + 1; };  // essential part of reflector i (length m-i-1)
    auto apply = [&](double* vec, int i) { launch_kr_house_apply(m - i, vec + i, Hess(i), d_tau + i, d_part, st_); };
    auto norm = [&](const double* a) {
        double h;
        launch_dot_det(m, a, a, d_part, d_beta, st_);
        CK(cudaMemcpyAsync(&h, d_beta, sizeof(double), cudaMemcpyDeviceToHost, st_));
        CK(cudaStreamSynchronize(st_));
        return std::sqrt(h);
    };
    // r0 = M^-1 (rhs - A x) into d_w, then the first reflector; returns beta
    auto start_cycle = [&]() {
        launch_spmv(m, d_rp, d_ci, d_v, d_x, d_w, st_);
        launch_kr_residual(m, d_b, d_w, st_);
        solve_device(d_w);
    };
    auto first_reflector = [&]() {
        double beta;
        CK(cudaMemsetAsync(d_tau, 0, sizeof(double) * (restart + 1), st_));
        launch_kr_house_make(m, d_w, Hess(0), d_tau, d_beta, d_part, st_);
        CK(cudaMemcpyAsync(&beta, d_beta, sizeof(double), cudaMemcpyDeviceToHost, st_));
        CK(cudaStreamSynchronize(st_));
        return beta;
    };
    struct Giv {
        double c = 1, s = 0;
        void make(double p, double q) {  // Eigen JacobiRotation::makeGivens, real scalars
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
        void apply_adjoint(double& xp, double& xq) const {
            double a = c * xp - s * xq, b = s * xp + c * xq;
            xp = a;
            xq = b;
        }
    };
    CK(cudaStreamSynchronize(st_));
    const double t0 = wtime();
    int result = 0;
    const int maxIters = iters;
    iters = 0;
    do {
        if (norm(d_b) <= std::numeric_limits<double>::min()) {
            CK(cudaMemsetAsync(d_x, 0, sizeof(double) * m, st_));
            result = 1;  // the reference returns `true`
            break;
        }
        start_cycle();
        const double r0Norm = norm(d_w);
        if (r0Norm == 0) {
            result = 1;
            break;
        }
        std::vector<double> R((size_t)(restart + 1) * (restart + 1), 0.0), w(restart + 1, 0.0), head(restart + 2, 0.0);
        std::vector<Giv> G(restart);
        w[0] = first_reflector();
        bool done = false;
        for (int k = 1; k <= restart && !done; ++k) {
            ++iters;
            launch_kr_unit(m, k - 1, d_w, st_);
            for (int i = k - 1; i >= 0; --i) apply(d_w, i);
            launch_spmv(m, d_rp, d_ci, d_v, d_w, d_t, st_);
            solve_device(d_t);
            for (int i = 0; i < k; ++i) apply(d_t, i);
            if (k < m) {  // new reflector from v.tail(m - k) (an all-zero tail gives tau = 0, like the skipped branch)
                launch_kr_house_make(m - k, d_t + k, Hess(k), d_tau + k, d_beta, d_part, st_);
                apply(d_t, k);
            }
            const int nh = std::min(k + 1, m);
            CK(cudaMemcpyAsync(head.data(), d_t, sizeof(double) * nh, cudaMemcpyDeviceToHost, st_));
            CK(cudaStreamSynchronize(st_));
            if (nh < k + 1) head[k] = 0.0;
            for (int i = 0; i < k - 1; ++i) G[i].apply_adjoint(head[i], head[i + 1]);
            if (k < m && head[k] != 0.0) {
                G[k - 1].make(head[k - 1], head[k]);
                G[k - 1].apply_adjoint(head[k - 1], head[k]);
                G[k - 1].apply_adjoint(w[k - 1], w[k]);
            }
            for (int i = 0; i < k; i++) R[i + (size_t)(k - 1) * (restart + 1)] = head[i];
            const double tol_error = std::fabs(w[k]) / r0Norm;
            const bool stop = (k == m || tol_error < tol_ || iters == maxIters);
            if (verbose) printf("%d: |Ax-b|/|b| = %3.2e <? %3.2e\n", iters, tol_error, tol_);
            if (stop || k == restart) {
                std::vector<double> y(w.begin(), w.begin() + k);
                for (int i = k - 1; i >= 0; --i) {
                    for (int j = i + 1; j < k; j++) y[i] -= R[i + (size_t)j * (restart + 1)] * y[j];
                    y[i] /= R[i + (size_t)i * (restart + 1)];
                }
                // x_new = H_0 ... H_{k-1} [y; 0]  (x_new(i) += y(i) before H_i only touches rows >= i)
                CK(cudaMemsetAsync(d_xn, 0, sizeof(double) * m, st_));
                CK(cudaMemcpyAsync(d_xn, y.data(), sizeof(double) * k, cudaMemcpyHostToDevice, st_));
                for (int i = k - 1; i >= 0; --i) apply(d_xn, i);
                launch_axpy(m, 1.0, d_xn, d_x, st_);
                CK(cudaStreamSynchronize(st_));  // y goes out of scope
                if (stop) {
                    done = true;
                    break;
                }
                k = 0;
                start_cycle();
                std::fill(R.begin(), R.end(), 0.0);
                std::fill(w.begin(), w.end(), 0.0);
                w[0] = first_reflector();
            }
        }
        result = iters;
    } while (false);
    CK(cudaMemcpyAsync(x, d_x, sizeof(double) * m, cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));
    if (seconds) *seconds = wtime() - t0;
    cudaFree(d_rp);
    cudaFree(d_ci);
    cudaFree(d_v);
    for (auto v : vecs) cudaFree(*v);
    cudaFree(d_H);
    cudaFree(d_tau);
    cudaFree(d_beta);
    cudaFree(d_part);
    return result;
}

// src/tree.cpp:1730-1763 (permuted ordering; both triangles for symmetric kinds). The live blocks at the point
// where factorize() stopped (parity-test hook) are read off the plan.
SpMat Tree::trailing_mat() {
    DeviceGuard dev_guard(device);
    std::vector<Triplet> t;
    std::vector<double> hb;
    if (!assembled_) throw std::runtime_error("trailing_mat: call assemble first");
    CK(cudaStreamSynchronize(st_));
    std::vector<int> live;
    bool piv_identity = false;
    if (state_level_ < 0) {
        for (int e = 0; e < plan_.nleaf_edges; e++) live.push_back(e);
    } else if (state_phase_ == 3) {
        const SymLevel& L = plan_.lv[state_level_];
        for (int e = L.medge0; e < L.medge1; e++) live.push_back(e);
    } else {
        const SymLevel& L = plan_.lv[state_level_];
        for (int e : L.s_piv) live.push_back(e);
        for (const SymTrsm& r : L.s_right) live.push_back(r.eB);
        piv_identity = state_phase_ >= 1;
    }
    for (int e : live) {
        const int n1 = plan_.en1[e], n2 = plan_.en2[e];
        const Cluster& cs = cl_[n1];
        const Cluster& c2 = cl_[n2];
        int rows = c2.size, cols = cs.size;
        hb.assign((size_t)rows * cols, 0.0);
        if (n1 == n2 && piv_identity) {
            for (int i = 0; i < rows; i++) hb[i + (size_t)i * rows] = 1.0;
        } else if (rows > 0 && cols > 0) {
            CK(cudaMemcpy2D(hb.data(), sizeof(double) * rows, h_eptr_[e], sizeof(double) * h_eld_[e], sizeof(double) * rows,
                            cols, cudaMemcpyDeviceToHost));
        }
        for (int j = 0; j < cols; j++)
            for (int i = 0; i < rows; i++) {
                int gi = c2.start + i, gj = cs.start + j;
                double v = hb[i + (size_t)j * rows];
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
    return from_triplets(N, N, t);
}

void Tree::cluster_layout(std::vector<int>& start, std::vector<int>& hlevel) const {
    start.clear();
    hlevel.clear();
    for (int h = 0; h < nlevels; h++)
        for (int c : bottoms_[h]) {
            start.push_back(cl_[c].start);
            hlevel.push_back(h);
        }
}

void Tree::stats(std::vector<int>& id, std::vector<int>& size, std::vector<int>& rank) const {
    id.clear();
    size.clear();
    rank.clear();
    for (int h = 0; h < nlevels; h++)
        for (int c : bottoms_[h]) {
            id.push_back(c);
            size.push_back(cl_[c].orig_size);
            rank.push_back(cl_[c].size);
        }
}

}  // namespace spand
