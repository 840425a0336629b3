#include "partition.hpp"
#include "parallel.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <limits>
#include <numeric>
#include <atomic>
#include <stdexcept>
#include <thread>
#include <unordered_map>

// METIS as shipped in the CUDA toolkit's static library: 64-bit idx_t (the reference wants
// the 32-bit build, src/partition.cpp:53; arrays are widened at this boundary instead).
extern "C" {
int METIS_SetDefaultOptions(int64_t* options);
int METIS_ComputeVertexSeparator(int64_t* nvtxs, int64_t* xadj, int64_t* adjncy, int64_t* vwgt, int64_t* options,
                                 int64_t* sepsize, int64_t* part);
}

namespace spand {

namespace {

// Graph without self loops (src/partition.cpp:21-39)
void graph_csc(const SpMat& A, std::vector<int>& colptr, std::vector<int>& rowval) {
    int n = A.rows;
    colptr.assign(n + 1, 0);
    rowval.clear();
    rowval.reserve(A.nnz());
    for (int j = 0; j < n; j++) {
        for (int k = A.colptr[j]; k < A.colptr[j + 1]; k++)
            if (A.rowind[k] != j) rowval.push_back(A.rowind[k]);
        colptr[j + 1] = (int)rowval.size();
    }
}

// src/partition.cpp:51-97 (vertex-separator branch only; Tree::partition hard-wires
// use_vertex_sep = true, src/tree.cpp:318)
void separator_metis(const std::vector<int>& colptr, const std::vector<int>& rowval, const std::vector<int>& dofs,
                     std::vector<int>& parts) {
    int size = (int)dofs.size();
    if (size == 0) return;
    std::vector<int64_t> xadj(size + 1, 0), adj;
    for (int i = 0; i < size; i++) {
        int g = dofs[i];
        for (int k = colptr[g]; k < colptr[g + 1]; k++) {
            int nb = rowval[k];
            if (nb == g) continue;
            auto it = std::lower_bound(dofs.begin(), dofs.end(), nb);
            if (it != dofs.end() && *it == nb) adj.push_back((int64_t)(it - dofs.begin()));
        }
        xadj[i + 1] = (int64_t)adj.size();
    }
    int64_t options[64];
    if (METIS_SetDefaultOptions(options) != 1) throw std::runtime_error("METIS_SetDefaultOptions failed");
    int64_t nv = size, sepsize = 0;
    std::vector<int64_t> p64(size, 0);
    if (adj.empty()) adj.push_back(0);
    int rc = METIS_ComputeVertexSeparator(&nv, xadj.data(), adj.data(), nullptr, options, &sepsize, p64.data());
    if (rc != 1) throw std::runtime_error("METIS_ComputeVertexSeparator failed");
    for (int i = 0; i < size; i++) parts[i] = (int)p64[i];
}

// fn(chunk, begin, end): loop [0, n) in contiguous, ordered chunks on `inner` host threads (the first depths of the bisection have fewer sub-domains
// than cores: the loops over the dofs of ONE sub-domain are split instead)
template <class F>
void inner_chunks(int inner, size_t n, F fn) {
    if (inner <= 1 || n < 32768) {
        fn(0, (size_t)0, n);
        return;
    }
    std::vector<std::thread> pool;
    for (int t = 1; t < inner; t++) pool.emplace_back(fn, t, n * t / inner, n * (t + 1) / inner);
    fn(0, (size_t)0, n / inner);
    for (auto& th : pool) th.join();
}

// src/partition.cpp:139-210
void separator_geo(const std::vector<int>& colptr, const std::vector<int>& rowval, const std::vector<int>& dofs,
                   std::vector<int>& parts, const DenseMat& X, std::vector<int>& invp, int inner) {
    int N = (int)dofs.size();
    if (N == 0) return;
    int bestdim = -1;
    double maxrange = -1.0;
    {
        const int D = X.rows;
        const int nch = (inner <= 1 || N < 32768) ? 1 : inner;
        std::vector<double> lo((size_t)nch * D, std::numeric_limits<double>::max());
        std::vector<double> hi((size_t)nch * D, std::numeric_limits<double>::lowest());
        inner_chunks(inner, (size_t)N, [&](int c, size_t b, size_t e) {
            for (int d = 0; d < D; d++) {
                double maxi = std::numeric_limits<double>::lowest(), mini = std::numeric_limits<double>::max();
                for (size_t i = b; i < e; i++) {
                    double x = X(d, dofs[i]);
                    maxi = std::max(maxi, x);
                    mini = std::min(mini, x);
                }
                lo[(size_t)c * D + d] = mini;
                hi[(size_t)c * D + d] = maxi;
            }
        });
        for (int d = 0; d < D; d++) {
            double maxi = std::numeric_limits<double>::lowest(), mini = std::numeric_limits<double>::max();
            for (int c = 0; c < nch; c++) {
                maxi = std::max(maxi, hi[(size_t)c * D + d]);
                mini = std::min(mini, lo[(size_t)c * D + d]);
            }
            if (maxi - mini > maxrange) {
                bestdim = d;
                maxrange = maxi - mini;
            }
        }
    }
    // Median coordinate = value at rank N/2 (the reference sorts the dofs; only this value is used)
    std::vector<double> xs(N);
    inner_chunks(inner, (size_t)N, [&](int, size_t b, size_t e) {
        for (size_t i = b; i < e; i++) xs[i] = X(bestdim, dofs[i]);
    });
    std::nth_element(xs.begin(), xs.begin() + N / 2, xs.end());
    double midv = xs[N / 2];
    for (int g : dofs) {
        invp[g] = -1;
        for (int k = colptr[g]; k < colptr[g + 1]; k++) invp[rowval[k]] = -1;
    }
    inner_chunks(inner, (size_t)N, [&](int, size_t b, size_t e) {
        for (size_t i = b; i < e; i++) {
            int g = dofs[i];
            invp[g] = (int)i;
            parts[i] = (X(bestdim, g) < midv) ? 0 : 1;
        }
    });
    // a dof of side 1 with a neighbour of side 0 goes to the separator (only sides 0 are looked at and they never
    // change, so the chunks are independent; the flags are applied afterwards, nobody reads what another one writes)
    std::vector<char> tosep(N, 0);
    inner_chunks(inner, (size_t)N, [&](int, size_t b, size_t e) {
        for (size_t i = b; i < e; i++) {
            if (parts[i] == 0) continue;
            int g = dofs[i];
            for (int k = colptr[g]; k < colptr[g + 1]; k++) {
                int in = invp[rowval[k]];
                if (in == -1) continue;
                if (parts[in] == 0) {
                    tosep[i] = 1;
                    break;
                }
            }
        }
    });
    inner_chunks(inner, (size_t)N, [&](int, size_t b, size_t e) {
        for (size_t i = b; i < e; i++)
            if (tosep[i]) parts[i] = 2;
    });
}

}  // namespace

std::vector<ClusterID> partition_modifiedND(const SpMat& A, int nlevels, const DenseMat* Xcoo, bool verb) {
    int N = A.rows;
    bool geo = (Xcoo != nullptr);
    if (verb) printf(geo ? "Geometric MND partitioning & ordering\n" : "Algebraic MND partitioning & ordering\n");
    std::vector<int> colptr, rowval;
    graph_csc(A, colptr, rowval);
    std::vector<std::vector<int>> doms(1, std::vector<int>(N));
    std::iota(doms[0].begin(), doms[0].end(), 0);
    SepID top(nlevels - 1, 0);
    std::vector<ClusterID> part(N, ClusterID(top, top, top));
    // The separators of one depth are independent: a dof shared by two sub-domains is touched by both, but each of
    // them only rewrites the ClusterID fields that carry its own separator id, and reads nothing the other one
    // writes. The geometric bisection therefore runs on a pool of host threads (same result whatever the schedule);
    // METIS keeps the reference's sequential call order.
    unsigned hw = std::thread::hardware_concurrency();
    if (const char* e = getenv("SPAND_HOST_THREADS")) hw = (unsigned)std::max(1, atoi(e));
    const int max_threads = geo ? (int)std::min<unsigned>(std::max(1u, hw), 16u) : 1;
    struct Work {
        std::vector<int> parttmp, scratch;
        long sepmin, sepmax, septot;
    };
    std::vector<Work> work(max_threads);
    std::vector<ClusterID> part_prev;
    for (int depth = 0; depth < nlevels - 1; depth++) {
        int level = nlevels - depth - 1;
        int nseps = 1 << depth;
        std::vector<std::vector<int>> newdoms(2 * nseps);
        const int nthreads = (N >= 100000) ? std::min(max_threads, nseps) : 1;
        std::atomic<int> next(0);
        // sub-domains of this depth go to the pool; while there are fewer of them than threads, the loops over the dofs
        // of one sub-domain are split instead (inner threads)
        const int inner = (N >= 100000 && nseps < max_threads) ? std::max(1, max_threads / nseps) : 1;
        part_prev.resize(part.size());
        parallel_chunks(part.size(), [&](int, size_t b, size_t e) {
            std::copy(part.begin() + b, part.begin() + e, part_prev.begin() + b);
        }, (size_t)(max_threads > 1 ? 262144 : ~(size_t)0 >> 1));
        auto run = [&](int tix) {
            Work& w = work[tix];
            if ((int)w.scratch.size() < N + 1) {
                w.parttmp.resize(N);
                w.scratch.resize(N + 1);
            }
            std::vector<int>& parttmp = w.parttmp;
            std::vector<int>& scratch = w.scratch;
            w.sepmin = N;
            w.sepmax = 0;
            w.septot = 0;
            for (int sep = next.fetch_add(1); sep < nseps; sep = next.fetch_add(1)) {
                SepID idself(level, sep);
                std::vector<int>& dofs = doms[sep];
                if (!std::is_sorted(dofs.begin(), dofs.end())) std::sort(dofs.begin(), dofs.end());
                if (geo) separator_geo(colptr, rowval, dofs, parttmp, *Xcoo, scratch, inner);
                else separator_metis(colptr, rowval, dofs, parttmp);
                SepID idleft(level - 1, 2 * sep), idright(level - 1, 2 * sep + 1);
                // A separator dof sits in two sub-domains (its l and r sides), i.e. two threads may hold it: each
                // reads the ClusterID as it was when the depth started (part_prev, immutable) and writes back only the
                // fields that carried its own separator id, which no other sub-domain of this depth can own.
                const int nch = (inner <= 1 || dofs.size() < 32768) ? 1 : inner;
                std::vector<std::vector<int>> lefts(nch), rights(nch);
                std::vector<long> newsep(nch, 0);
                inner_chunks(inner, dofs.size(), [&](int c, size_t b, size_t e) {
                    std::vector<int>& L = lefts[c];
                    std::vector<int>& R = rights[c];
                    long nnew = 0;
                    for (size_t i = b; i < e; i++) {
                        const int g = dofs[i];
                        const ClusterID& p0 = part_prev[g];
                        ClusterID q = p0;
                        int side = parttmp[i];
                        if (side == 0 || side == 1) {
                            const SepID& idside = (side == 0 ? idleft : idright);
                            if (q.self == idself) q.self = idside;
                            if (q.l == idself) q.l = idside;
                            if (q.r == idself) q.r = idside;
                        } else if (side == 2) {
                            if (q.self == idself) {
                                q.l = idleft;
                                q.r = idright;
                            }
                        }
                        ClusterID& p = part[g];
                        if (!(q.self == p0.self)) p.self = q.self;
                        if (!(q.l == p0.l)) p.l = q.l;
                        if (!(q.r == p0.r)) p.r = q.r;
                        if (q.self == idleft || q.l == idleft || q.r == idleft) L.push_back(g);
                        if (q.self == idright || q.l == idright || q.r == idright) R.push_back(g);
                        if (q.self == idself && q.l == idleft && q.r == idright) nnew++;
                    }
                    newsep[c] = nnew;
                });
                long nnewsep = 0;
                for (int c = 0; c < nch; c++) {
                    newdoms[2 * sep].insert(newdoms[2 * sep].end(), lefts[c].begin(), lefts[c].end());
                    newdoms[2 * sep + 1].insert(newdoms[2 * sep + 1].end(), rights[c].begin(), rights[c].end());
                    nnewsep += newsep[c];
                }
                w.sepmin = std::min(w.sepmin, nnewsep);
                w.sepmax = std::max(w.sepmax, nnewsep);
                w.septot += nnewsep;
            }
        };
        if (nthreads <= 1) {
            run(0);
        } else {
            std::vector<std::thread> pool;
            for (int tix = 1; tix < nthreads; tix++) pool.emplace_back(run, tix);
            run(0);
            for (auto& th : pool) th.join();
        }
        long sepmin = N, sepmax = 0, septot = 0;
        for (int tix = 0; tix < std::max(1, nthreads); tix++) {
            sepmin = std::min(sepmin, work[tix].sepmin);
            sepmax = std::max(sepmax, work[tix].sepmax);
            septot += work[tix].septot;
        }
        doms.swap(newdoms);
        if (getenv("SPAND_TIMING_DEPTH")) {
            struct timespec ts;
            clock_gettime(CLOCK_MONOTONIC, &ts);
            static double tprev = 0;
            double tn = ts.tv_sec + 1e-9 * ts.tv_nsec;
            fprintf(stderr, "[spand] MND depth %d: %d sub-domains, %d threads x %d inner, +%.1f ms\n", depth, nseps, nthreads,
                    inner, tprev > 0 ? (tn - tprev) * 1e3 : 0.0);
            tprev = tn;
        }
        if (verb)
            printf("  Depth %2d: (%5d separators, [%5ld %5ld], mean %6.1f)\n", depth + 1, nseps, sepmin, sepmax,
                   (double)septot / nseps);
    }
    return part;
}

Ordering build_ordering(const SpMat& A, int nlevels, const DenseMat* Xcoo, bool verb) {
    if (A.rows != A.cols) throw std::runtime_error("partition: matrix must be square");
    if (nlevels <= 0) throw std::runtime_error("partition: nlevels must be > 0");
    if (Xcoo && Xcoo->cols != A.rows) throw std::runtime_error("partition: coordinates must be dim x N");
    Ordering o;
    int N = A.rows;
    o.N = N;
    o.nlevels = nlevels;
    const bool timing = getenv("SPAND_TIMING") != nullptr;
    auto now = [] {
        struct timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec + 1e-9 * ts.tv_nsec;
    };
    double t0 = now();
    o.part = partition_modifiedND(A, nlevels, Xcoo, verb);
    if (timing) fprintf(stderr, "[spand] partition: modified ND       %8.2f ms\n", (now() - t0) * 1e3);
    t0 = now();
    // Ordering: identity, then L stable sorts by progressively merged ClusterIDs (tree.cpp:344-352). Dofs with the same
    // ClusterID have the same key in every one of those sorts and stay together in index order, so the sorts are done
    // on the distinct ClusterIDs (the future leaf clusters, ~N/10) and the dofs are placed by one counting pass: the
    // same permutation at a fraction of the cost (2 M dofs: 4.9 s -> 0.2 s).
    std::vector<int> gid(N);
    std::vector<ClusterID> gids;
    {
        struct Key {
            long long a, b, c;
            bool operator==(const Key& o) const { return a == o.a && b == o.b && c == o.c; }
        };
        struct KeyHash {
            size_t operator()(const Key& k) const {
                unsigned long long h = (unsigned long long)k.a * 0x9E3779B97F4A7C15ull;
                h ^= (unsigned long long)k.b * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
                h ^= (unsigned long long)k.c * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
                return (size_t)h;
            }
        };
        auto key_of = [](const ClusterID& c) {
            return Key{((long long)c.self.lvl << 32) | (unsigned)c.self.sep, ((long long)c.l.lvl << 32) | (unsigned)c.l.sep,
                       ((long long)c.r.lvl << 32) | (unsigned)c.r.sep};
        };
        std::unordered_map<Key, int, KeyHash> index;
        index.reserve((size_t)N / 4 + 16);
        for (int i = 0; i < N; i++) {
            auto it = index.find(key_of(o.part[i]));
            if (it == index.end()) {
                it = index.emplace(key_of(o.part[i]), (int)gids.size()).first;
                gids.push_back(o.part[i]);
            }
            gid[i] = it->second;
        }
    }
    const int ng = (int)gids.size();
    std::vector<int> gorder(ng);
    std::iota(gorder.begin(), gorder.end(), 0);
    {
        std::vector<ClusterID> merged = gids;
        auto cmp = [&merged](int i, int j) { return merged[i] < merged[j]; };
        // stable sort = stable sorts of contiguous chunks on host threads + stable merges of neighbouring runs
        // (std::inplace_merge keeps the left run first on ties): the same permutation as one std::stable_sort
        auto sort_stable = [&]() {
            const size_t n = gorder.size();
            std::vector<size_t> cut;
            const int nth = parallel_chunks(n, [&](int, size_t b, size_t e) {
                std::stable_sort(gorder.begin() + b, gorder.begin() + e, cmp);
            }, 8192);
            for (int t = 0; t <= nth; t++) cut.push_back(n * t / nth);
            while (cut.size() > 2) {
                std::vector<size_t> next;
                std::vector<std::thread> pool;
                for (size_t k = 0; k + 2 < cut.size(); k += 2) {
                    pool.emplace_back([&, k] {
                        std::inplace_merge(gorder.begin() + cut[k], gorder.begin() + cut[k + 1], gorder.begin() + cut[k + 2], cmp);
                    });
                }
                for (auto& th : pool) th.join();
                for (size_t k = 0; k < cut.size(); k += 2) next.push_back(cut[k]);
                if (next.back() != n) next.push_back(n);
                cut.swap(next);
            }
        };
        sort_stable();
        for (int lvl = 1; lvl < nlevels; lvl++) {
            parallel_chunks(merged.size(), [&](int, size_t b, size_t e) {
                for (size_t i = b; i < e; i++) merged[i] = merge_if(merged[i], lvl);
            }, 8192);
            sort_stable();
        }
    }
    {
        std::vector<int> count(ng, 0), offset(ng, 0);
        for (int i = 0; i < N; i++) count[gid[i]]++;
        int run = 0;
        for (int g : gorder) {
            offset[g] = run;
            run += count[g];
        }
        o.perm.resize(N);
        for (int i = 0; i < N; i++) o.perm[offset[gid[i]]++] = i;
    }
    if (timing) fprintf(stderr, "[spand] partition: ordering (sorts)  %8.2f ms\n", (now() - t0) * 1e3);
    t0 = now();
    // Leaf clusters = maximal runs of equal ClusterID (tree.cpp:360-370)
    o.levels.assign(nlevels, {});
    int order = 0;
    for (int k = 0; k < N;) {
        const ClusterID& id = o.part[o.perm[k]];
        int knext = k + 1;
        while (knext < N && o.part[o.perm[knext]] == id) knext++;
        ClusterNode c;
        c.start = k;
        c.size = knext - k;
        c.level = id.self.lvl;
        c.order = order++;
        c.sparsify = (id.l.lvl == 0 && id.r.lvl == 0);
        c.parent = -1;
        c.child_begin = c.child_end = 0;
        c.id = id;
        o.levels[0].push_back(c);
        k = knext;
    }
    // Hierarchy (tree.cpp:381-415)
    for (int lvl = 1; lvl < nlevels; lvl++) {
        auto& prev = o.levels[lvl - 1];
        size_t begin = 0;
        while (begin < prev.size() && prev[begin].level < lvl) begin++;
        for (size_t k = begin; k < prev.size();) {
            if (prev[k].level < lvl) throw std::runtime_error("partition: hierarchy is not sorted by level");
            ClusterID idp = merge_if(prev[k].id, lvl);
            ClusterNode p;
            p.start = prev[k].start;
            p.size = 0;
            p.child_begin = (int)k;
            while (k < prev.size() && merge_if(prev[k].id, lvl) == idp) {
                p.size += prev[k].size;
                prev[k].parent = (int)o.levels[lvl].size();
                k++;
            }
            p.child_end = (int)k;
            p.level = idp.self.lvl;
            p.order = order++;
            p.sparsify = (idp.l.lvl == lvl && idp.r.lvl == lvl);
            p.parent = -1;
            p.id = idp;
            o.levels[lvl].push_back(p);
        }
    }
    o.norders = order;
    if (timing) fprintf(stderr, "[spand] partition: hierarchy         %8.2f ms\n", (now() - t0) * 1e3);
    if (verb) {
        printf("Hierarchy numbers (# of cluster at each level of the cluster-hierarchy)\n");
        for (int lvl = 0; lvl < nlevels; lvl++) printf("%3d %9zu\n", lvl, o.levels[lvl].size());
    }
    return o;
}

}  // namespace spand
