#include "tree.hpp"

#include <sys/time.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>

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
        CK(cudaMalloc((void**)&c.p, c.cap));
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
    CK(cudaMallocHost((void**)&pinned_, cap_));
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
    std::memcpy(pinned_ + top_, src, bytes);
    CK(cudaMemcpyAsync(dst, pinned_ + top_, bytes, cudaMemcpyHostToDevice, st));
    top_ += (bytes + 63) & ~(size_t)63;
}

// ------------------------------------------------------------------------------------------------
// construction / partition / assemble
// ------------------------------------------------------------------------------------------------
Tree::Tree(int nlevels_) : nlevels(nlevels_) {
    if (nlevels <= 0) throw std::runtime_error("nlevels must be > 0");
    log.assign(nlevels, LevelLog());
}

Tree::~Tree() { free_device(); }

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
    stager_.reserve((size_t)64 << 20);
    CK(cudaMalloc((void**)&d_err_, sizeof(int)));
    for (int i = 0; i < kSide; i++) {
        CK(cudaStreamCreateWithFlags(&side_[i], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ev_join_[i], cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
}

void Tree::free_device() {
    if (!st_) return;
    cudaStreamSynchronize(st_);
    delete arena_;
    delete scratch_;
    arena_ = scratch_ = nullptr;
    if (d_err_) cudaFree(d_err_);
    d_err_ = nullptr;
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

void Tree::set_coords(int dim, int N_, const double* X) {
    Xcoo_ = DenseMat(dim, N_);
    std::copy(X, X + (size_t)dim * N_, Xcoo_.a.begin());
    have_coords_ = true;
}

// src/tree.cpp:306-419
void Tree::partition(const SpMat& A) {
    if (use_geo && !have_coords_) throw std::runtime_error("use_geo set without coordinates");
    ord = build_ordering(A, nlevels, use_geo ? &Xcoo_ : nullptr, verb);
    N = A.rows;
    log.assign(nlevels, LevelLog());
    for (auto& p : ord.part) log[p.self.lvl].dofs_nd += 1;
    for (int l = nlevels - 2; l >= 0; l--) log[l].dofs_left_nd = log[l + 1].dofs_nd + log[l + 1].dofs_left_nd;
}

int Tree::new_edge(int n1, int n2, double* A, int ld, bool original) {
    Edge e;
    e.n1 = n1;
    e.n2 = n2;
    e.A = A;
    e.ld = ld;
    e.original = original;
    e.alive = true;
    e.identity = false;
    ed_.push_back(e);
    return (int)ed_.size() - 1;
}

int Tree::find_out(int c, int n2) const {
    for (int e : cl_[c].out)
        if (ed_[e].n2 == n2) return e;
    return -1;
}

// src/tree.cpp:505-575 — dense blocks are built on the host once and uploaded in one copy
void Tree::assemble(const SpMat& A) {
    if (N == 0 || A.rows != N) throw std::runtime_error("assemble: call partition first with a matrix of the same size");
    ensure_device();
    CK(cudaStreamSynchronize(st_));
    arena_->reset();
    scratch_->reset();
    stager_.reset();
    solve_.assign(nlevels, SolveLevel());
    nnz_ = 0;
    launches_total = 0;
    factorized_ = false;
    current_bottom_ = 0;
    ilvl_ = 0;
    for (auto& l : log) {
        LevelLog fresh;
        fresh.dofs_nd = l.dofs_nd;
        fresh.dofs_left_nd = l.dofs_left_nd;
        l = fresh;
    }
    // clusters from the ordering; id == order
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
    ed_.clear();
    d_csize_ = arena_->alloc_n<int>(ord.norders);
    d_perm_ = arena_->alloc_n<int>(N);
    d_xnat_ = arena_->alloc_n<double>(N);
    double* d_xleaf = arena_->alloc_n<double>(N);
    stager_.upload(d_perm_, ord.perm.data(), sizeof(int) * N, st_);
    h_csize_.assign(ord.norders, 0);
    for (int c : bottoms_[0]) {
        cl_[c].x = d_xleaf + cl_[c].start;
        h_csize_[c] = cl_[c].size;
    }
    stager_.upload(d_csize_, h_csize_.data(), sizeof(int) * ord.norders, st_);

    SpMat App = symm_perm(A, ord.perm);
    std::vector<int> cmap(N);
    for (int c : bottoms_[0])
        for (int k = cl_[c].start; k < cl_[c].start + cl_[c].size; k++) cmap[k] = c;
    // pass 1: structure
    struct Blk { int n1, n2; size_t off; };
    std::vector<Blk> blks;
    size_t total = 0;
    std::vector<int> mark(ord.norders, -1);
    std::vector<int> nb;
    for (int s : bottoms_[0]) {
        nb.clear();
        const Cluster& cs = cl_[s];
        for (int j = cs.start; j < cs.start + cs.size; j++)
            for (int k = App.colptr[j]; k < App.colptr[j + 1]; k++) {
                int row = App.rowind[k];
                if (symmetry() && row < j) continue;
                int n = cmap[row];
                if (mark[n] != s) {
                    mark[n] = s;
                    nb.push_back(n);
                }
            }
        if (mark[s] != s) {
            mark[s] = s;
            nb.push_back(s);
        }
        std::sort(nb.begin(), nb.end());
        // pivot first, the rest by increasing order (cluster.cpp:113-123,162-176)
        blks.push_back({s, s, total});
        total += (size_t)cs.size * cs.size;
        for (int n : nb)
            if (n != s) {
                blks.push_back({s, n, total});
                total += (size_t)cl_[n].size * cs.size;
            }
    }
    // pass 2: values (util.cpp:454-486 block2dense) into one host buffer
    double* dblocks = arena_->alloc_n<double>(total);
    std::vector<double> hb(total, 0.0);
    for (auto& b : blks) {
        const Cluster& c1 = cl_[b.n1];
        const Cluster& c2 = cl_[b.n2];
        double* dst = hb.data() + b.off;
        for (int col = 0; col < c1.size; col++) {
            int j = c1.start + col;
            const int* rb = App.rowind.data() + App.colptr[j];
            const int* re = App.rowind.data() + App.colptr[j + 1];
            const int* it = std::lower_bound(rb, re, c2.start);
            for (; it < re && *it < c2.start + c2.size; ++it)
                dst[(*it - c2.start) + (size_t)col * c2.size] = App.val[it - App.rowind.data()];
        }
        int e = new_edge(b.n1, b.n2, dblocks + b.off, c2.size, true);
        if (b.n1 == b.n2) cl_[b.n1].out.insert(cl_[b.n1].out.begin(), e);
        else {
            cl_[b.n1].out.push_back(e);
            cl_[b.n2].in.push_back(e);
        }
    }
    CK(cudaMemcpyAsync(dblocks, hb.data(), total * sizeof(double), cudaMemcpyHostToDevice, st_));
    CK(cudaStreamSynchronize(st_));
    stager_.reset();
}

int Tree::ndofs_left() const {
    int n = 0;
    for (int c : bottoms_[current_bottom_])
        if (!cl_[c].eliminated) n += cl_[c].size;
    return n;
}

void Tree::check_error() {
    int err = 0;
    CK(cudaMemcpyAsync(&err, d_err_, sizeof(int), cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));
    if (err & 1) throw std::runtime_error("Error: Non-SPD Pivot\n");
    if (err & 2) throw std::runtime_error("Error: Singular Pivot\n");
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
// ELIMINATE — src/tree.cpp:895-967 for every cluster of level ilvl (mutually non-adjacent)
// ------------------------------------------------------------------------------------------------
void Tree::phase_eliminate(LevelLog& lg, SolveLevel& sl) {
    if (scale_kind == PLU) {
        phase_eliminate_plu(lg, sl);
        return;
    }
    std::vector<int> E;
    for (int c : bottoms_[current_bottom_])
        if (cl_[c].level == ilvl_ && !cl_[c].eliminated) E.push_back(c);
    if (E.empty()) return;
    std::vector<PotrfTask> potrf;
    std::vector<TrsmTask> trsm;
    for (int s : E) {
        const Cluster& cs = cl_[s];
        const Edge& piv = ed_[cs.out[0]];
        double n = cs.size;
        potrf.push_back({piv.A, piv.ld, cs.size});
        lg.fl_pivot += n * n * n / 3.0;
        if (!cs.in.empty()) throw std::runtime_error("eliminate: unexpected in-edges on an SPD interior");
        for (size_t k = 1; k < cs.out.size(); k++) {
            const Edge& e = ed_[cs.out[k]];
            TrsmTask t{};
            t.B = e.A;
            t.ldb = e.ld;
            t.T = piv.A;
            t.ldt = piv.ld;
            t.m = cl_[e.n2].size;
            t.n = cs.size;
            trsm.push_back(t);
            lg.fl_panel += (double)t.m * n * n;
        }
    }
    run_potrf(potrf, lg);
    run_trsm(TRSM_RLT, trsm, lg);

    // Schur complement: targets are found / created exactly in the reference's loop order (tree.cpp:862-869,
    // gemm_edges :761-772) so that fill-in edges enter the out/in lists in the same sequence.
    struct Triple { int target, e1, e2; };
    std::vector<Triple> triples;
    std::vector<int> task_of_edge;  // lazily sized
    std::vector<int> targets;       // edge ids in first-seen order
    std::vector<char> fresh;
    for (int s : E) {
        const std::vector<int>& out = cl_[s].out;
        for (size_t a = 1; a < out.size(); a++) {
            int e1 = out[a];
            int n1 = ed_[e1].n2;
            for (size_t b = 1; b < out.size(); b++) {
                int e2 = out[b];
                int n2 = ed_[e2].n2;
                if (n1 < n2) continue;  // id == order
                int tg = find_out(n2, n1);
                bool is_new = false;
                if (tg < 0) {
                    double* A = arena_->alloc_n<double>((size_t)cl_[n1].size * cl_[n2].size);
                    tg = new_edge(n2, n1, A, cl_[n1].size, false);
                    cl_[n2].out.push_back(tg);
                    cl_[n1].in.push_back(tg);
                    is_new = true;
                }
                if ((int)task_of_edge.size() <= tg) task_of_edge.resize(ed_.size() + 1024, -1);
                if (task_of_edge[tg] < 0) {
                    task_of_edge[tg] = (int)targets.size();
                    targets.push_back(tg);
                    fresh.push_back(is_new);
                }
                triples.push_back({tg, e1, e2});
            }
        }
    }
    {
        std::vector<GemmTask> tasks(targets.size());
        std::vector<int> count(targets.size(), 0);
        for (auto& t : triples) count[task_of_edge[t.target]]++;
        int off = 0;
        for (size_t i = 0; i < targets.size(); i++) {
            const Edge& e = ed_[targets[i]];
            GemmTask& g = tasks[i];
            g.C = e.A;
            g.ldc = e.ld;
            g.m = cl_[e.n2].size;
            g.n = cl_[e.n1].size;
            g.c0 = off;
            g.nc = 0;
            g.flags = (e.n1 == e.n2 ? GEMM_LOWER : 0) | (fresh[i] ? GEMM_ZERO_INIT : 0);
            off += count[i];
        }
        std::vector<GemmContrib> con(triples.size());
        for (auto& t : triples) {
            GemmTask& g = tasks[task_of_edge[t.target]];
            const Edge& a = ed_[t.e1];
            const Edge& b = ed_[t.e2];
            int k = cl_[a.n1].size;
            con[g.c0 + g.nc++] = {a.A, b.A, a.ld, b.ld, k};
            if (a.n2 == b.n2) lg.fl_schur += (double)g.m * (g.m + 1) * k;
            else lg.fl_schur += 2.0 * g.m * g.n * k;
        }
        run_gemm(tasks, con, lg);
    }

    // Record the operations (tree.cpp:909-910, :885-892) as solve batches, count nnz, drop the clusters.
    std::vector<TrsvTask> trsv;
    std::map<int, std::vector<GemvContrib>> fwd_by_target;  // x_n -= A[n,s] x_s, in s order
    std::vector<GemvTask> bwd_tasks;
    std::vector<GemvContrib> bwd_con;
    for (int s : E) {
        Cluster& cs = cl_[s];
        const Edge& piv = ed_[cs.out[0]];
        trsv.push_back({piv.A, cs.x, piv.ld, cs.size});
        nnz_ += (long long)cs.size * (cs.size + 1) / 2;
        GemvTask bt;
        bt.y = cs.x;
        bt.m = cs.size;
        bt.c0 = (int)bwd_con.size();
        bt.nc = 0;
        for (size_t k = 1; k < cs.out.size(); k++) {
            const Edge& e = ed_[cs.out[k]];
            int nsz = cl_[e.n2].size;
            nnz_ += (long long)cs.size * nsz;
            if (nsz == 0 || cs.size == 0) continue;
            fwd_by_target[e.n2].push_back({e.A, cs.x, e.ld, cs.size});
            bwd_con.push_back({e.A, cl_[e.n2].x, e.ld, nsz});
            bt.nc++;
        }
        if (bt.nc > 0) bwd_tasks.push_back(bt);
    }
    std::vector<GemvTask> fwd_tasks;
    std::vector<GemvContrib> fwd_con;
    for (auto& kv : fwd_by_target) {
        GemvTask t;
        t.y = cl_[kv.first].x;
        t.m = cl_[kv.first].size;
        t.c0 = (int)fwd_con.size();
        t.nc = (int)kv.second.size();
        fwd_con.insert(fwd_con.end(), kv.second.begin(), kv.second.end());
        fwd_tasks.push_back(t);
    }
    sl.e_trsv = to_device(trsv, arena_);
    sl.n_e_trsv = (int)trsv.size();
    sl.e_gemv_f = to_device(fwd_tasks, arena_);
    sl.e_gemv_fc = to_device(fwd_con, arena_);
    sl.n_e_gemv_f = (int)fwd_tasks.size();
    sl.e_gemv_b = to_device(bwd_tasks, arena_);
    sl.e_gemv_bc = to_device(bwd_con, arena_);
    sl.n_e_gemv_b = (int)bwd_tasks.size();

    // set_eliminated (cluster.cpp:32-45)
    std::vector<int> touched;
    for (int s : E) {
        Cluster& cs = cl_[s];
        for (int e : cs.out) {
            ed_[e].alive = false;
            if (ed_[e].n2 != s) touched.push_back(ed_[e].n2);
        }
        cs.out.clear();
        cs.in.clear();
        cs.eliminated = true;
    }
    std::sort(touched.begin(), touched.end());
    touched.erase(std::unique(touched.begin(), touched.end()), touched.end());
    for (int n : touched) {
        auto& in = cl_[n].in;
        in.erase(std::remove_if(in.begin(), in.end(), [&](int e) { return !ed_[e].alive; }), in.end());
    }
}

// ------------------------------------------------------------------------------------------------
// GEN / PLU variants (src/tree.cpp:614-689, :735-742, :929-956, :838-853; src/util.cpp:183-227)
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

void Tree::alloc_plu(Cluster& cs) {
    size_t n = std::max(1, cs.size);
    cs.ud = arena_->alloc_n<double>(n);
    cs.ipiv = arena_->alloc_n<int>(n);
    cs.perm = arena_->alloc_n<int>(n);
}

void Tree::phase_eliminate_plu(LevelLog& lg, SolveLevel& sl) {
    std::vector<int> E;
    for (int c : bottoms_[current_bottom_])
        if (cl_[c].level == ilvl_ && !cl_[c].eliminated) E.push_back(c);
    if (E.empty()) return;
    std::vector<GetrfTask> getrf;
    std::vector<RowPermTask> rperm;
    std::vector<TrsmTask> left, right;
    for (int s : E) {
        Cluster& cs = cl_[s];
        const Edge& piv = ed_[cs.out[0]];
        double n = cs.size;
        alloc_plu(cs);
        getrf.push_back({piv.A, piv.ld, cs.size, cs.ud, cs.ipiv, cs.perm});
        lg.fl_pivot += 2.0 * n * n * n / 3.0;
        for (int eid : cs.in) {  // A[s,n] <- L^-1 P^T A[s,n]   (tree.cpp:668-676)
            const Edge& e = ed_[eid];
            int w = cl_[e.n1].size;
            rperm.push_back({e.A, e.ld, cs.size, w, cs.perm});
            TrsmTask t{};
            t.B = e.A;
            t.ldb = e.ld;
            t.T = piv.A;
            t.ldt = piv.ld;
            t.m = w;
            t.n = cs.size;
            left.push_back(t);
            lg.fl_panel += (double)w * n * n;
        }
        for (size_t k = 1; k < cs.out.size(); k++) {  // A[n,s] <- A[n,s] U^-1   (tree.cpp:681-689, Q = I)
            const Edge& e = ed_[cs.out[k]];
            TrsmTask t{};
            t.B = e.A;
            t.ldb = e.ld;
            t.T = piv.A;
            t.ldt = piv.ld;
            t.m = cl_[e.n2].size;
            t.n = cs.size;
            t.diag = cs.ud;
            right.push_back(t);
            lg.fl_panel += (double)t.m * n * n;
        }
    }
    run_getrf(getrf, lg);
    run_rowperm(rperm, lg);
    run_trsm(TRSM_LLN, left, lg);
    run_trsm(TRSM_RUN, right, lg);

    // Schur complement A[n1,n2] -= A[n1,s] A[s,n2] for every (out, in) pair in the reference's loop order
    // (tree.cpp:943-947); fill-in edges are created where gemm_edges would create them (:761-772).
    struct Triple { int target, e1, e2; };
    std::vector<Triple> triples;
    std::vector<int> task_of_edge, targets;
    std::vector<char> fresh;
    for (int s : E) {
        const std::vector<int>& out = cl_[s].out;
        const std::vector<int>& in = cl_[s].in;
        for (size_t a = 1; a < out.size(); a++) {
            int e1 = out[a];
            int n1 = ed_[e1].n2;
            for (size_t b = 0; b < in.size(); b++) {
                int e2 = in[b];
                int n2 = ed_[e2].n1;
                int tg = find_out(n2, n1);
                bool is_new = false;
                if (tg < 0) {
                    double* A = arena_->alloc_n<double>((size_t)cl_[n1].size * cl_[n2].size);
                    tg = new_edge(n2, n1, A, std::max(1, cl_[n1].size), false);
                    cl_[n2].out.push_back(tg);
                    cl_[n1].in.push_back(tg);
                    is_new = true;
                }
                if ((int)task_of_edge.size() <= tg) task_of_edge.resize(ed_.size() + 1024, -1);
                if (task_of_edge[tg] < 0) {
                    task_of_edge[tg] = (int)targets.size();
                    targets.push_back(tg);
                    fresh.push_back(is_new);
                }
                triples.push_back({tg, e1, e2});
            }
        }
    }
    {
        std::vector<GemmTask> tasks(targets.size());
        std::vector<int> count(targets.size(), 0);
        for (auto& t : triples) count[task_of_edge[t.target]]++;
        int off = 0;
        for (size_t i = 0; i < targets.size(); i++) {
            const Edge& e = ed_[targets[i]];
            GemmTask& g = tasks[i];
            g.C = e.A;
            g.ldc = e.ld;
            g.m = cl_[e.n2].size;
            g.n = cl_[e.n1].size;
            g.c0 = off;
            g.nc = 0;
            g.flags = GEMM_NN | (fresh[i] ? GEMM_ZERO_INIT : 0);
            off += count[i];
        }
        std::vector<GemmContrib> con(triples.size());
        for (auto& t : triples) {
            GemmTask& g = tasks[task_of_edge[t.target]];
            const Edge& a = ed_[t.e1];  // A[n1,s]
            const Edge& b = ed_[t.e2];  // A[s,n2]
            int k = cl_[a.n1].size;
            con[g.c0 + g.nc++] = {a.A, b.A, a.ld, b.ld, k};
            lg.fl_schur += 2.0 * g.m * g.n * k;
        }
        run_gemm(tasks, con, lg);
    }

    // recorded operations: ScalingPLUQ, GemmOut (forward only), GemmIn (backward only)
    std::vector<TrsvTask> trsv;
    std::map<int, std::vector<GemvContrib>> fwd_by_target;
    std::vector<GemvTask> bwd_tasks;
    std::vector<GemvContrib> bwd_con;
    for (int s : E) {
        Cluster& cs = cl_[s];
        const Edge& piv = ed_[cs.out[0]];
        trsv.push_back({piv.A, cs.x, piv.ld, cs.size, cs.ud, cs.perm});
        nnz_ += (long long)cs.size * cs.size + 2LL * cs.size;
        for (size_t k = 1; k < cs.out.size(); k++) {
            const Edge& e = ed_[cs.out[k]];
            int nsz = cl_[e.n2].size;
            nnz_ += (long long)cs.size * nsz;
            if (nsz == 0 || cs.size == 0) continue;
            fwd_by_target[e.n2].push_back({e.A, cs.x, e.ld, cs.size});
        }
        GemvTask bt;
        bt.y = cs.x;
        bt.m = cs.size;
        bt.c0 = (int)bwd_con.size();
        bt.nc = 0;
        for (int eid : cs.in) {
            const Edge& e = ed_[eid];
            int nsz = cl_[e.n1].size;
            nnz_ += (long long)cs.size * nsz;
            if (nsz == 0 || cs.size == 0) continue;
            bwd_con.push_back({e.A, cl_[e.n1].x, e.ld, nsz});  // x_s -= A[s,n] x_n  (no transpose)
            bt.nc++;
        }
        if (bt.nc > 0) bwd_tasks.push_back(bt);
    }
    std::vector<GemvTask> fwd_tasks;
    std::vector<GemvContrib> fwd_con;
    for (auto& kv : fwd_by_target) {
        GemvTask t;
        t.y = cl_[kv.first].x;
        t.m = cl_[kv.first].size;
        t.c0 = (int)fwd_con.size();
        t.nc = (int)kv.second.size();
        fwd_con.insert(fwd_con.end(), kv.second.begin(), kv.second.end());
        fwd_tasks.push_back(t);
    }
    sl.e_trsv = to_device(trsv, arena_);
    sl.n_e_trsv = (int)trsv.size();
    sl.e_gemv_f = to_device(fwd_tasks, arena_);
    sl.e_gemv_fc = to_device(fwd_con, arena_);
    sl.n_e_gemv_f = (int)fwd_tasks.size();
    sl.e_gemv_b = to_device(bwd_tasks, arena_);
    sl.e_gemv_bc = to_device(bwd_con, arena_);
    sl.n_e_gemv_b = (int)bwd_tasks.size();

    // set_eliminated (cluster.cpp:32-45): drop the cluster's out-edges and in-edges from their other endpoints
    std::vector<int> touched;
    for (int s : E) {
        Cluster& cs = cl_[s];
        for (int e : cs.out) {
            ed_[e].alive = false;
            if (ed_[e].n2 != s) touched.push_back(ed_[e].n2);
        }
        for (int e : cs.in) {
            ed_[e].alive = false;
            touched.push_back(ed_[e].n1);
        }
        cs.out.clear();
        cs.in.clear();
        cs.eliminated = true;
    }
    std::sort(touched.begin(), touched.end());
    touched.erase(std::unique(touched.begin(), touched.end()), touched.end());
    for (int n : touched) {
        auto& in = cl_[n].in;
        in.erase(std::remove_if(in.begin(), in.end(), [&](int e) { return !ed_[e].alive; }), in.end());
        auto& out = cl_[n].out;
        out.erase(std::remove_if(out.begin(), out.end(), [&](int e) { return !ed_[e].alive; }), out.end());
    }
}

void Tree::phase_scale_plu(LevelLog& lg, SolveLevel& sl) {
    std::vector<GetrfTask> getrf;
    std::vector<RowPermTask> rperm;
    std::vector<TrsmTask> right, left;
    std::vector<TrsvTask> trsv;
    const std::vector<int>& bottom = bottoms_[current_bottom_];
    for (int c : bottom) {
        Cluster& cs = cl_[c];
        if (cs.eliminated || cs.level <= ilvl_) continue;
        alloc_plu(cs);
    }
    for (int c : bottom) {
        Cluster& cs = cl_[c];
        if (cs.eliminated || cs.level <= ilvl_) continue;
        Edge& piv = ed_[cs.out[0]];
        double n = cs.size;
        getrf.push_back({piv.A, piv.ld, cs.size, cs.ud, cs.ipiv, cs.perm});
        trsv.push_back({piv.A, cs.x, piv.ld, cs.size, cs.ud, cs.perm});
        nnz_ += (long long)cs.size * cs.size + 2LL * cs.size;
        lg.fl_pivot += 2.0 * n * n * n / 3.0;
        lg.by_scale += 16.0 * n * n;
        for (size_t k = 1; k < cs.out.size(); k++) {
            const Edge& e = ed_[cs.out[k]];
            const Cluster& c2 = cl_[e.n2];
            const Edge& piv2 = ed_[c2.out[0]];
            // block A[n2,n1] (|n2| x |n1|): out-edge of n1 -> B U_n1^-1 ; in-edge of n2 -> L_n2^-1 P_n2^T B
            TrsmTask r{};
            r.B = e.A;
            r.ldb = e.ld;
            r.T = piv.A;
            r.ldt = piv.ld;
            r.m = c2.size;
            r.n = cs.size;
            r.diag = cs.ud;
            right.push_back(r);
            rperm.push_back({e.A, e.ld, c2.size, cs.size, c2.perm});
            TrsmTask l{};
            l.B = e.A;
            l.ldb = e.ld;
            l.T = piv2.A;
            l.ldt = piv2.ld;
            l.m = cs.size;
            l.n = c2.size;
            left.push_back(l);
            lg.fl_panel += (double)c2.size * n * n + (double)cs.size * c2.size * c2.size;
            lg.by_scale += 16.0 * c2.size * n;
        }
    }
    run_getrf(getrf, lg);
    run_trsm(TRSM_RUN, right, lg);
    run_rowperm(rperm, lg);
    run_trsm(TRSM_LLN, left, lg);
    for (int c : bottom) {
        Cluster& cs = cl_[c];
        if (cs.eliminated || cs.level <= ilvl_) continue;
        ed_[cs.out[0]].identity = true;  // tree.cpp:848 — never materialised
    }
    sl.s_trsv = to_device(trsv, arena_);
    sl.n_s_trsv = (int)trsv.size();
}

// ------------------------------------------------------------------------------------------------
// SCALE — src/tree.cpp:796-856 for every remaining cluster: pivot = L L^T, every incident block
// A[n2,n1] <- L_n2^-1 A[n2,n1] L_n1^-T, pivot := I (kept implicit)
// ------------------------------------------------------------------------------------------------
void Tree::phase_scale(LevelLog& lg, SolveLevel& sl) {
    if (scale_kind == PLU) {
        phase_scale_plu(lg, sl);
        return;
    }
    std::vector<PotrfTask> potrf;
    std::vector<TrsmTask> right, left;
    std::vector<TrsvTask> trsv;
    for (int c : bottoms_[current_bottom_]) {
        Cluster& cs = cl_[c];
        if (cs.eliminated || cs.level <= ilvl_) continue;
        Edge& piv = ed_[cs.out[0]];
        double n = cs.size;
        potrf.push_back({piv.A, piv.ld, cs.size});
        trsv.push_back({piv.A, cs.x, piv.ld, cs.size});
        nnz_ += (long long)cs.size * (cs.size + 1) / 2;
        lg.fl_pivot += n * n * n / 3.0;
        lg.by_scale += 16.0 * n * n;
        for (size_t k = 1; k < cs.out.size(); k++) {
            const Edge& e = ed_[cs.out[k]];
            const Cluster& c2 = cl_[e.n2];
            const Edge& piv2 = ed_[c2.out[0]];
            // right: B (|n2| x |n1|) <- B L_n1^-T ; left: B <- L_n2^-1 B
            right.push_back({e.A, piv.A, e.ld, piv.ld, c2.size, cs.size});
            left.push_back({e.A, piv2.A, e.ld, piv2.ld, cs.size, c2.size});
            lg.fl_panel += (double)c2.size * n * n + (double)cs.size * c2.size * c2.size;
            lg.by_scale += 16.0 * c2.size * n;
        }
    }
    run_potrf(potrf, lg);
    run_trsm(TRSM_RLT, right, lg);
    run_trsm(TRSM_LLN, left, lg);
    for (int c : bottoms_[current_bottom_]) {
        Cluster& cs = cl_[c];
        if (cs.eliminated || cs.level <= ilvl_) continue;
        ed_[cs.out[0]].identity = true;  // tree.cpp:811,853 — never materialised
    }
    sl.s_trsv = to_device(trsv, arena_);
    sl.n_s_trsv = (int)trsv.size();
}

// ------------------------------------------------------------------------------------------------
// SPARSIFY — src/tree.cpp:1417-1433 -> 1292-1347. The reference sweeps the clusters in list order
// (Gauss-Seidel: a cluster sees the already-shrunk blocks of earlier neighbours). The same result is
// obtained by wavefronts of the dependency DAG; clusters of one wavefront are mutually non-adjacent.
// ------------------------------------------------------------------------------------------------
void Tree::phase_sparsify(LevelLog& lg, SolveLevel& sl) {
    const double plan0 = wtime();
    const std::vector<int>& bottom = bottoms_[current_bottom_];
    std::vector<int> S;
    for (int c : bottom) {
        const Cluster& cs = cl_[c];
        if (cs.eliminated || cs.level <= ilvl_) continue;
        bool want = use_want_sparsify ? cs.sparsify : true;
        if (want) S.push_back(c);
        else lg.ignored++;
    }
    int first = bottom.empty() ? 0 : bottom.front();
    int span = bottom.empty() ? 0 : bottom.back() - first + 1;
    std::vector<int> color(span, -1);
    std::vector<int> old_size(span, 0);
    int ncolors = 0;
    if (!S.empty()) {
        std::vector<QrTask> tasks;
        std::vector<QrSrc> srcs;
        std::vector<int> task_color;
        for (int s : S) {
            const Cluster& cs = cl_[s];
            int col = 0;
            QrTask t;
            t.cluster = s;
            t.rows = cs.size;
            t.src0 = (int)srcs.size();
            int maxcols = 0;
            auto visit = [&](int nbr, const Edge& e, int transposed) {
                int cn = color[nbr - first];
                if (cn >= 0) col = std::max(col, cn + 1);  // earlier in list order and sparsified
                srcs.push_back({e.A, e.ld, nbr, transposed});
                maxcols += cl_[nbr].size;
            };
            for (int e : cs.in) visit(ed_[e].n1, ed_[e], 0);
            for (size_t k = 1; k < cs.out.size(); k++) visit(ed_[cs.out[k]].n2, ed_[cs.out[k]], 1);
            t.nsrc = (int)srcs.size() - t.src0;
            t.maxcols = maxcols;
            color[s - first] = col;
            ncolors = std::max(ncolors, col + 1);
            t.W = nullptr;
            int kmax = std::min(t.rows, maxcols);
            t.V = arena_->alloc_n<double>((size_t)t.rows * kmax);
            t.tau = arena_->alloc_n<double>(kmax);
            tasks.push_back(t);
            task_color.push_back(col);
        }
        // Launch classes: (threads, cluster size G, shared-memory bucket). The panel stays resident in (distributed)
        // shared memory whenever it fits in 16 CTAs; otherwise it lives in the (L2-resident) scratch arena.
        const int MAXS = rrqr_max_smem();
        static const int kBuckets[] = {12 << 10, 24 << 10, 40 << 10, 72 << 10, 112 << 10, 224 << 10};
        auto pow2floor = [](int x) { int p = 1; while (2 * p <= x) p *= 2; return p; };
        auto pow2ceil = [](int x) { int p = 1; while (p < x) p *= 2; return p; };
        // class = (mode << 8) | (log2 G << 4) | bucket; mode 0: 128 threads (G = 1), 1: 256 threads, 2: panel in the
        // global scratch (512 threads, G = 8), 3: 512 threads
        std::vector<int> klass(tasks.size());
        std::vector<int> smem_need(tasks.size());
        const int kNB = 6;
        // test hooks: force every task through one kernel shape (tests/test_gpu_parity.py)
        const bool force_global = getenv("SPAND_RRQR_FORCE_GLOBAL") != nullptr;
        const int force_g = getenv("SPAND_RRQR_FORCE_G") ? atoi(getenv("SPAND_RRQR_FORCE_G")) : 0;
        std::vector<int> per_color(ncolors, 0);
        for (int c : task_color) per_color[c]++;
        for (size_t i = 0; i < tasks.size(); i++) {
            QrTask& t = tasks[i];
            int mn = std::max(1, std::min(t.rows, t.maxcols));
            t.nb = std::min(QR_NB, mn);
            auto config = [&](int NT, int G, bool in_smem) {
                int cpcm = std::max(1, (t.maxcols + G - 1) / G);
                int L = in_smem ? std::min(32, std::max(1, pow2floor(NT / cpcm))) : 32;
                L = std::min(L, pow2ceil(std::max(1, (t.rows + 1) / 2)));  // a lane works on pairs of rows
                int ld = (t.rows + 1) & ~1;
                if (in_smem && L < 8)
                    while (ld % 16 != (2 * L) % 16) ld += 2;  // conflict-free 128-bit shared loads
                t.L = L;
                t.ld = ld;
                t.in_smem = in_smem ? 1 : 0;
                return (long)rrqr_smem_bytes(t.rows, t.maxcols, G, t.nb, t.ld, in_smem);
            };
            auto bucket_of = [&](long nd) {
                int b = 0;
                while (b < kNB - 1 && nd > kBuckets[b]) b++;
                return b;
            };
            long nd = config(128, 1, true);
            if (!force_global && force_g == 0 && t.rows <= 64 && nd <= kBuckets[2]) {
                klass[i] = bucket_of(nd);
                smem_need[i] = (int)nd;
                continue;
            }
            // Cluster size: wide enough that the wavefront fills the GPU (about two CTAs per SM), at least what the
            // panel needs to stay in shared memory.
            int want = 1;
            while (want < 16 && per_color[task_color[i]] * want < 296) want *= 2;
            if (force_g > 0) want = force_g;
            int g = 0;
            while ((1 << g) < want) g++;
            int gi = -1, nt = 256;
            for (; g <= 4 && gi < 0 && !force_global; g++) {
                nd = config(256, 1 << g, true);
                if (nd <= kBuckets[4]) gi = g;
                else if (nd <= MAXS) {
                    nt = 512;
                    nd = config(512, 1 << g, true);
                    gi = g;
                }
            }
            if (gi >= 0) {
                klass[i] = ((nt == 256 ? 1 : 3) << 8) | (gi << 4) | bucket_of(nd);
                smem_need[i] = (int)nd;
                continue;
            }
            // panel stays in the scratch arena: 16 CTAs stream their slabs from L2
            nd = config(512, 16, false);
            while (nd > MAXS && t.nb > 2) {
                t.nb /= 2;
                nd = config(512, 16, false);
            }
            if (nd > MAXS) throw std::runtime_error("sparsify: interface cluster too large for the RRQR kernel");
            t.W = scratch_->alloc_n<double>((size_t)t.ld * t.maxcols);
            klass[i] = (2 << 8) | (4 << 4) | bucket_of(nd);
            smem_need[i] = (int)nd;
        }
        // order tasks by (colour, class), stable
        std::vector<int> idx(tasks.size());
        for (size_t i = 0; i < idx.size(); i++) idx[i] = (int)i;
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) {
            if (task_color[a] != task_color[b]) return task_color[a] < task_color[b];
            return klass[a] < klass[b];
        });
        std::vector<QrTask> sorted(tasks.size());
        for (size_t i = 0; i < idx.size(); i++) sorted[i] = tasks[idx[i]];
        QrTask* dt = to_device(sorted, scratch_);
        QrSrc* ds = to_device(srcs, scratch_);
        // One launch per (colour, class). The classes of a colour are independent: they run concurrently on side
        // streams forked from / joined into the factorization stream.
        for (size_t b = 0; b < idx.size();) {
            size_t cend = b;
            while (cend < idx.size() && task_color[idx[cend]] == task_color[idx[b]]) cend++;
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
                launch_rrqr(dt + b, (int)(e - b), ds, d_csize_, tol, G, mode == 0 ? 128 : (mode == 1 ? 256 : 512),
                            mode != 2, smem, s);
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
        lg.wavefronts = ncolors;
        lg.t_plan_spars = wtime() - plan0;
        // ranks back to the host: the one synchronisation of the level
        for (int c : bottom) old_size[c - first] = cl_[c].size;
        CK(cudaMemcpyAsync(h_csize_.data() + first, d_csize_ + first, sizeof(int) * span, cudaMemcpyDeviceToHost, st_));
        check_error();
        std::vector<HouseTask> house;
        FILE* dump = nullptr;
        if (const char* fn = getenv("SPAND_DUMP_QR")) dump = fopen(fn, "a");
        for (size_t i = 0; i < tasks.size(); i++) {
            const QrTask& t = tasks[i];
            Cluster& cs = cl_[t.cluster];
            int rank = h_csize_[t.cluster];
            if (dump) fprintf(dump, "%d %d %d %d %d %d\n", ilvl_, t.cluster, t.rows, t.maxcols, rank, task_color[i]);
            // columns actually seen by this cluster (earlier sparsified neighbours had already shrunk)
            long cols = 0;
            for (int k = 0; k < t.nsrc; k++) {
                int nbr = srcs[t.src0 + k].nbr;
                // position in list order == id order inside one bottom
                cols += (color[nbr - first] >= 0 && nbr < t.cluster) ? h_csize_[nbr] : old_size[nbr - first];
            }
            double r = t.rows, cc = (double)cols, rf = std::min<double>(t.rows, cols), rk = rank;
            if (tol >= 1.0 || cols == 0) rf = 0;
            lg.rank_before += t.rows;
            lg.nspars++;
            lg.nbrs += cols;
            lg.fl_rrqr_full += 4 * r * cc * rf - 2 * (r + cc) * rf * rf + (4.0 / 3.0) * rf * rf * rf;
            lg.fl_rrqr_rank += 4 * r * cc * rk - 2 * (r + cc) * rk * rk + (4.0 / 3.0) * rk * rk * rk;
            lg.by_rrqr += 8 * r * cc + 8 * rk * cc + 8 * r * rk;
            if (rank < t.rows) {
                house.push_back({t.V, t.tau, cs.x, t.rows, rank});
                nnz_ += (long long)t.rows * t.rows;  // Orthogonal (operations.cpp:159-161)
                long long m = t.rows - rank;
                // Scaling op of the dropped sibling (tree.cpp:1342): ScalingLLT(I) or ScalingPLUQ(I, I, id, id)
                nnz_ += scale_kind == PLU ? m * m + 2 * m : m * (m + 1) / 2;
                cs.size = rank;
            }
            lg.rank_after += cs.size;
        }
        if (dump) fclose(dump);
        sl.house = to_device(house, arena_);
        sl.n_house = (int)house.size();
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
    current_bottom_++;
    const std::vector<int>& parents = bottoms_[current_bottom_];
    if (parents.empty()) return;
    std::vector<int> pos(ord.norders, 0);
    std::vector<XCopyTask> mf, mb;
    size_t xtotal = 0;
    for (int p : parents) {
        Cluster& cp = cl_[p];
        int size = 0;
        for (int c = cp.child_begin; c < cp.child_end; c++) {
            pos[c] = size;
            size += cl_[c].size;
        }
        cp.size = cp.orig_size = size;
        h_csize_[p] = size;
        xtotal += size;
    }
    double* xbase = arena_->alloc_n<double>(xtotal + 1);
    {
        size_t off = 0;
        for (int p : parents) {
            Cluster& cp = cl_[p];
            cp.x = xbase + off;
            off += cp.size;
            for (int c = cp.child_begin; c < cp.child_end; c++) {
                if (cl_[c].size == 0) continue;
                mf.push_back({cl_[c].x, cp.x + pos[c], cl_[c].size});
                mb.push_back({cp.x + pos[c], cl_[c].x, cl_[c].size});
            }
        }
    }
    int pfirst = parents.front();
    stager_.upload(d_csize_ + pfirst, h_csize_.data() + pfirst, sizeof(int) * parents.size(), st_);
    sl.m_fwd = to_device(mf, arena_);
    sl.m_bwd = to_device(mb, arena_);
    sl.n_merge = (int)mf.size();

    // new edges: structure first, then one zero-filled allocation, then the block copies
    struct NewEdge { int n1, n2; size_t off; bool original; };
    std::vector<NewEdge> ne;
    std::vector<CopyTask> copies;
    struct PendingCopy { int edge_old; int newidx; };
    std::vector<PendingCopy> pend;
    size_t total = 0;
    std::vector<int> slot(ord.norders, -1);
    std::vector<int> tlist;
    for (int p : parents) {
        Cluster& cp = cl_[p];
        tlist.clear();
        for (int c = cp.child_begin; c < cp.child_end; c++)
            for (int e : cl_[c].out) {
                int q = cl_[ed_[e].n2].parent;
                if (slot[q] == -1) {
                    slot[q] = -2;
                    tlist.push_back(q);
                }
            }
        std::sort(tlist.begin(), tlist.end());
        // pivot first, then by increasing order
        size_t base = ne.size();
        ne.push_back({p, p, 0, false});
        slot[p] = (int)base;
        for (int q : tlist)
            if (q != p) {
                slot[q] = (int)ne.size();
                ne.push_back({p, q, 0, false});
            }
        for (size_t i = base; i < ne.size(); i++) {
            ne[i].off = total;
            total += (size_t)cl_[ne[i].n2].size * cp.size;
        }
        for (int c = cp.child_begin; c < cp.child_end; c++)
            for (int e : cl_[c].out) {
                int q = cl_[ed_[e].n2].parent;
                if (ed_[e].original) ne[slot[q]].original = true;
                pend.push_back({e, slot[q]});
            }
        for (int q : tlist) slot[q] = -1;
        slot[p] = -1;
    }
    double* nb = arena_->alloc_n<double>(total + 1);
    auto evz = fam_begin(F_COPY);
    CK(cudaMemsetAsync(nb, 0, total * sizeof(double), st_));
    fam_end(F_COPY, evz);
    lg.by_merge += 8.0 * total;
    std::vector<int> new_ids(ne.size());
    for (size_t i = 0; i < ne.size(); i++) {
        int ld = std::max(1, cl_[ne[i].n2].size);
        int e = new_edge(ne[i].n1, ne[i].n2, nb + ne[i].off, ld, ne[i].original);
        new_ids[i] = e;
        if (ne[i].n1 == ne[i].n2) cl_[ne[i].n1].out.insert(cl_[ne[i].n1].out.begin(), e);
        else {
            cl_[ne[i].n1].out.push_back(e);
            cl_[ne[i].n2].in.push_back(e);
        }
    }
    const int CHUNK = 4096;
    for (auto& pc : pend) {
        Edge& eo = ed_[pc.edge_old];
        const Edge& en = ed_[new_ids[pc.newidx]];
        int rows = cl_[eo.n2].size, cols = cl_[eo.n1].size;
        eo.alive = false;
        if (rows == 0 || cols == 0) continue;
        double* dst = en.A + pos[eo.n2] + (size_t)pos[eo.n1] * en.ld;
        lg.by_merge += 8.0 * rows * cols;
        if (eo.identity) {
            copies.push_back({nullptr, dst, 0, en.ld, rows, cols});
            continue;
        }
        int cstep = std::max(1, CHUNK / rows);
        for (int c0 = 0; c0 < cols; c0 += cstep) {
            int w = std::min(cstep, cols - c0);
            copies.push_back({eo.A + (size_t)c0 * eo.ld, dst + (size_t)c0 * en.ld, eo.ld, en.ld, rows, w});
        }
    }
    for (int p : parents)
        for (int c = cl_[p].child_begin; c < cl_[p].child_end; c++) {
            cl_[c].out.clear();
            cl_[c].in.clear();
        }
    CopyTask* dc = to_device(copies, scratch_);
    auto evc = fam_begin(F_COPY);
    launch_copy(dc, (int)copies.size(), st_);
    fam_end(F_COPY, evc);
    lg.launches += 2;
}

// ------------------------------------------------------------------------------------------------
// FACTORIZE — src/tree.cpp:1447-1551
// ------------------------------------------------------------------------------------------------
void Tree::factorize() {
    if (symm_kind == SPD && scale_kind != LLT) throw std::runtime_error("SPD requires LLT scaling");
    if (symm_kind == GEN && scale_kind != PLU) throw std::runtime_error("GEN requires PLU scaling (PLUQ is out of scope)");
    if (symm_kind == SYM) throw std::runtime_error("SYM/LDLT is out of scope (SURVEY.md section 2)");
    if (ed_.empty() || factorized_) throw std::runtime_error("factorize: call assemble first");
    ensure_device();
    CK(cudaMemsetAsync(d_err_, 0, sizeof(int), st_));
    for (int f = 0; f < F_COUNT; f++) {
        family_ms[f] = 0;
        family_launches[f] = 0;
    }
    std::vector<cudaEvent_t> ev(nlevels * 5);
    for (auto& e : ev) CK(cudaEventCreate(&e));
    cudaEvent_t ev_begin, ev_end;
    CK(cudaEventCreate(&ev_begin));
    CK(cudaEventCreate(&ev_end));
    CK(cudaEventRecord(ev_begin, st_));
    bool stopped = false;
    int last_level = -1;
    for (ilvl_ = 0; ilvl_ < nlevels && !stopped; ilvl_++) {
        LevelLog& lg = log[ilvl_];
        SolveLevel& sl = solve_[ilvl_];
        last_level = ilvl_;
        double h0 = wtime();
        if (verb) printf("Level %d, %d dofs left\n", ilvl_, ndofs_left());
        CK(cudaEventRecord(ev[ilvl_ * 5 + 0], st_));
        double p0 = wtime();
        phase_eliminate(lg, sl);
        lg.t_plan_elim = wtime() - p0;
        CK(cudaEventRecord(ev[ilvl_ * 5 + 1], st_));
        lg.dofs_left_elim = ndofs_left();
        if (ilvl_ == stop_level && stop_phase == 0) stopped = true;
        if (!stopped && ilvl_ >= skip) {
            p0 = wtime();
            phase_scale(lg, sl);
            lg.t_plan_scale = wtime() - p0;
            CK(cudaEventRecord(ev[ilvl_ * 5 + 2], st_));
            if (ilvl_ == stop_level && stop_phase == 1) stopped = true;
            if (!stopped) {
                phase_sparsify(lg, sl);
                if (ilvl_ == stop_level && stop_phase == 2) stopped = true;
            }
        } else {
            CK(cudaEventRecord(ev[ilvl_ * 5 + 2], st_));
            check_error();
            stager_.reset();
            scratch_->reset();
        }
        CK(cudaEventRecord(ev[ilvl_ * 5 + 3], st_));
        p0 = wtime();
        if (!stopped && ilvl_ < nlevels - 1) phase_merge(lg, sl);
        lg.t_plan_merge = wtime() - p0;
        CK(cudaEventRecord(ev[ilvl_ * 5 + 4], st_));
        lg.dofs_left_spars = ndofs_left();
        lg.fact_nnz = nnz_;
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
    for (auto& e : ev) cudaEventDestroy(e);
    cudaEventDestroy(ev_begin);
    cudaEventDestroy(ev_end);
    stager_.reset();
    scratch_->reset();
    factorized_ = !stopped;
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
    if (!factorized_) throw std::runtime_error("solve: call factorize first");
    double* xleaf = cl_[bottoms_[0][0]].x - cl_[bottoms_[0][0]].start;
    launch_gather(N, d_perm_, x_dev, xleaf, st_);  // b = P^T x
    for (int l = 0; l < nlevels; l++) {
        SolveLevel& s = solve_[l];
        launch_trsv(s.e_trsv, s.n_e_trsv, 0, st_);
        launch_gemv(s.e_gemv_f, s.n_e_gemv_f, s.e_gemv_fc, 0, st_);
        launch_trsv(s.s_trsv, s.n_s_trsv, 0, st_);
        launch_house(s.house, s.n_house, 1, st_);
        launch_xcopy(s.m_fwd, s.n_merge, st_);
    }
    for (int l = nlevels - 1; l >= 0; l--) {
        SolveLevel& s = solve_[l];
        launch_xcopy(s.m_bwd, s.n_merge, st_);
        launch_house(s.house, s.n_house, 0, st_);
        // LLT: x <- L^-T x, x_s -= A[n,s]^T x_n ; PLU: x <- U^-1 x, x_s -= A[s,n] x_n   (operations.cpp bwd)
        const bool plu = scale_kind == PLU;
        launch_trsv(s.s_trsv, s.n_s_trsv, plu ? 2 : 1, st_);
        launch_gemv(s.e_gemv_b, s.n_e_gemv_b, s.e_gemv_bc, plu ? 0 : 1, st_);
        launch_trsv(s.e_trsv, s.n_e_trsv, plu ? 2 : 1, st_);
    }
    launch_scatter(N, d_perm_, xleaf, x_dev, st_);  // x = P b
}

void Tree::solve(double* x_host) {
    if (!factorized_) throw std::runtime_error("solve: call factorize first");
    CK(cudaMemcpyAsync(d_xnat_, x_host, sizeof(double) * N, cudaMemcpyHostToDevice, st_));
    solve_device(d_xnat_);
    CK(cudaMemcpyAsync(x_host, d_xnat_, sizeof(double) * N, cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));
}

// src/is.cpp:39-121 with every vector resident in HBM
int Tree::cg(const SpMat& A, const double* rhs, double* x, int iters, double tol_, bool verbose, double* seconds) {
    if (!factorized_) throw std::runtime_error("cg: call factorize first");
    int n = A.cols;
    // CSR of A = CSC of A^T
    std::vector<Triplet> tt;
    tt.reserve(A.nnz());
    for (int j = 0; j < n; j++)
        for (int k = A.colptr[j]; k < A.colptr[j + 1]; k++) tt.push_back({j, A.rowind[k], A.val[k]});
    SpMat At = from_triplets(n, n, tt);
    int *d_rp, *d_ci;
    double *d_v, *d_x, *d_r, *d_p, *d_z, *d_tmp, *d_b, *d_s;
    CK(cudaMalloc((void**)&d_rp, sizeof(int) * (n + 1)));
    CK(cudaMalloc((void**)&d_ci, sizeof(int) * At.nnz()));
    CK(cudaMalloc((void**)&d_v, sizeof(double) * At.nnz()));
    double** vecs[] = {&d_x, &d_r, &d_p, &d_z, &d_tmp, &d_b};
    for (auto v : vecs) CK(cudaMalloc((void**)v, sizeof(double) * n));
    CK(cudaMalloc((void**)&d_s, sizeof(double) * 4));
    CK(cudaMemcpyAsync(d_rp, At.colptr.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice, st_));
    CK(cudaMemcpyAsync(d_ci, At.rowind.data(), sizeof(int) * At.nnz(), cudaMemcpyHostToDevice, st_));
    CK(cudaMemcpyAsync(d_v, At.val.data(), sizeof(double) * At.nnz(), cudaMemcpyHostToDevice, st_));
    CK(cudaMemcpyAsync(d_b, rhs, sizeof(double) * n, cudaMemcpyHostToDevice, st_));
    CK(cudaMemcpyAsync(d_x, x, sizeof(double) * n, cudaMemcpyHostToDevice, st_));
    auto dot = [&](const double* a, const double* b) {
        double h;
        CK(cudaMemsetAsync(d_s, 0, sizeof(double), st_));
        launch_dot(n, a, b, d_s, st_);
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
    return result;
}

// src/tree.cpp:1730-1763 (permuted ordering; both triangles for symmetric kinds)
SpMat Tree::trailing_mat() {
    std::vector<Triplet> t;
    std::vector<double> hb;
    CK(cudaStreamSynchronize(st_));
    for (int s : bottoms_[current_bottom_]) {
        const Cluster& cs = cl_[s];
        if (cs.eliminated) continue;
        for (int eid : cs.out) {
            const Edge& e = ed_[eid];
            const Cluster& c2 = cl_[e.n2];
            int rows = c2.size, cols = cs.size;
            hb.assign((size_t)rows * cols, 0.0);
            if (e.identity) {
                for (int i = 0; i < rows; i++) hb[i + (size_t)i * rows] = 1.0;
            } else if (rows > 0 && cols > 0) {
                CK(cudaMemcpy2D(hb.data(), sizeof(double) * rows, e.A, sizeof(double) * e.ld, sizeof(double) * rows,
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
    }
    return from_triplets(N, N, t);
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
