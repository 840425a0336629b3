// Symbolic replay of Tree::factorize (reference src/tree.cpp:1447-1551) on integers only. See symbolic.hpp.
#include "symbolic.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <stdexcept>

#include "host/parallel.hpp"
#include "kernels.cuh"

namespace spand {

namespace {
template <class T>
size_t vbytes(const std::vector<T>& v) { return v.size() * sizeof(T); }
}  // namespace

size_t SymbolicPlan::bytes() const {
    size_t b = vbytes(en1) + vbytes(en2);
    for (auto& l : lv)
        b += vbytes(l.E) + vbytes(l.e_piv) + vbytes(l.e_out) + vbytes(l.e_in) + vbytes(l.e_gemm) + vbytes(l.e_con) +
             vbytes(l.e_gf) + vbytes(l.e_gb) + vbytes(l.e_gfc) + vbytes(l.e_gbc) + vbytes(l.S) + vbytes(l.s_piv) +
             vbytes(l.s_right) + vbytes(l.s_left) + vbytes(l.q) + vbytes(l.qs) + vbytes(l.m_copy);
    return b;
}

void build_symbolic(const std::vector<SymCluster>& cl, const std::vector<std::vector<int>>& bottoms,
                    const std::vector<int>& leaf_n1, const std::vector<int>& leaf_n2, bool symmetric, bool want_flag,
                    SymbolicPlan& plan) {
    const int nlevels = (int)bottoms.size();
    const int ncl = (int)cl.size();
    plan = SymbolicPlan();
    plan.nlevels = nlevels;
    plan.symmetric = symmetric;
    plan.want_flag = want_flag;
    plan.lv.assign(nlevels, SymLevel());
    std::vector<int>& en1 = plan.en1;
    std::vector<int>& en2 = plan.en2;
    std::vector<std::vector<int>> out(ncl), in(ncl);  // edge ids; out: pivot first (cluster.cpp:113-123)
    std::vector<char> eliminated(ncl, 0);
    auto new_edge = [&](int n1, int n2) {
        en1.push_back(n1);
        en2.push_back(n2);
        return (int)en1.size() - 1;
    };
    auto find_out = [&](int c, int n2) {
        for (int e : out[c])
            if (en2[e] == n2) return e;
        return -1;
    };
    // leaf edges (src/tree.cpp:505-575); the adjacency lists are sized first (one allocation per list instead of the
    // doubling sequence of push_back: 10^6 lists)
    {
        std::vector<int> od(ncl, 0), id(ncl, 0);
        for (size_t i = 0; i < leaf_n1.size(); i++) {
            od[leaf_n1[i]]++;
            if (leaf_n1[i] != leaf_n2[i]) id[leaf_n2[i]]++;
        }
        parallel_chunks((size_t)ncl, [&](int, size_t b, size_t e) {
            for (size_t c = b; c < e; c++) {
                if (od[c]) out[c].reserve(od[c] + 8);  // + fill-in edges of the first levels
                if (id[c]) in[c].reserve(id[c]);
            }
        }, 65536);
    }
    for (size_t i = 0; i < leaf_n1.size(); i++) {
        int e = new_edge(leaf_n1[i], leaf_n2[i]);
        if (leaf_n1[i] == leaf_n2[i]) out[leaf_n1[i]].insert(out[leaf_n1[i]].begin(), e);
        else {
            out[leaf_n1[i]].push_back(e);
            in[leaf_n2[i]].push_back(e);
        }
    }
    plan.nleaf_edges = (int)en1.size();

    std::vector<int> task_of_edge;
    std::vector<int> color;
    std::vector<std::vector<int>> slots(kMaxHostThreads);  // per host thread: parent -> new edge index during a merge
    std::vector<char> edge_dead;
    double tsec[6] = {0, 0, 0, 0, 0, 0};
    auto nowf = [] {
        struct timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec + 1e-9 * ts.tv_nsec;
    };
    double tlast = nowf();
    auto lap = [&](int i) {
        double n = nowf();
        tsec[i] += n - tlast;
        tlast = n;
    };
    for (int l = 0; l < nlevels; l++) {
        SymLevel& L = plan.lv[l];
        const std::vector<int>& bottom = bottoms[l];
        // ---------------- eliminate ----------------
        for (int c : bottom)
            if (cl[c].level == l && !eliminated[c]) L.E.push_back(c);
        for (int s : L.E) {
            if (out[s].empty()) throw std::runtime_error("symbolic: cluster without pivot");
            const int piv = out[s][0];
            L.e_piv.push_back(piv);
            if (symmetric && !in[s].empty()) throw std::runtime_error("eliminate: unexpected in-edges on an SPD interior");
            for (int e : in[s]) L.e_in.push_back({e, piv, en1[e], s});
            for (size_t k = 1; k < out[s].size(); k++) L.e_out.push_back({out[s][k], piv, en2[out[s][k]], s});
        }
        lap(0);
        // Schur complement targets in the reference's loop order (tree.cpp:862-869 / :943-947, gemm_edges :761-772)
        L.fill0 = (int)en1.size();
        {
            struct Triple { int target, e1, e2; };
            std::vector<Triple> triples;
            std::vector<int> targets;
            std::vector<char> fresh;
            auto visit = [&](int e1, int e2, int col_cluster, int row_cluster) {
                int tg = find_out(col_cluster, row_cluster);
                bool is_new = false;
                if (tg < 0) {
                    tg = new_edge(col_cluster, row_cluster);
                    out[col_cluster].push_back(tg);
                    in[row_cluster].push_back(tg);
                    is_new = true;
                }
                if ((int)task_of_edge.size() <= tg) task_of_edge.resize(en1.size() + 1024, -1);
                if (task_of_edge[tg] < 0) {
                    task_of_edge[tg] = (int)targets.size();
                    targets.push_back(tg);
                    fresh.push_back(is_new);
                }
                triples.push_back({tg, e1, e2});
            };
            for (int s : L.E) {
                const std::vector<int>& o = out[s];
                if (symmetric) {
                    for (size_t a = 1; a < o.size(); a++) {
                        int n1 = en2[o[a]];
                        for (size_t b = 1; b < o.size(); b++) {
                            int n2 = en2[o[b]];
                            if (n1 < n2) continue;  // lower blocks only; id == order
                            visit(o[a], o[b], n2, n1);
                        }
                    }
                } else {
                    const std::vector<int>& i_ = in[s];
                    for (size_t a = 1; a < o.size(); a++)
                        for (size_t b = 0; b < i_.size(); b++) visit(o[a], i_[b], en1[i_[b]], en2[o[a]]);
                }
            }
            L.e_gemm.resize(targets.size());
            std::vector<int> count(targets.size(), 0);
            for (auto& t : triples) count[task_of_edge[t.target]]++;
            int off = 0;
            for (size_t i = 0; i < targets.size(); i++) {
                int e = targets[i];
                SymGemm& g = L.e_gemm[i];
                g.target = e;
                g.c0 = off;
                g.nc = 0;
                if (symmetric) g.flags = (en1[e] == en2[e] ? GEMM_LOWER : 0) | (fresh[i] ? GEMM_ZERO_INIT : 0);
                else g.flags = GEMM_NN | (fresh[i] ? GEMM_ZERO_INIT : 0);
                off += count[i];
            }
            L.e_con.resize(triples.size());
            for (auto& t : triples) {
                SymGemm& g = L.e_gemm[task_of_edge[t.target]];
                L.e_con[g.c0 + g.nc++] = {t.e1, t.e2};
            }
            for (int e : targets) task_of_edge[e] = -1;
        }
        L.fill1 = (int)en1.size();
        lap(1);
        // recorded operations (tree.cpp:909-910, :885-892 / GEN :950-956)
        {
            struct F { int target, edge, src; };
            std::vector<F> fwd;
            for (int s : L.E) {
                SymGemv bt{s, (int)L.e_gbc.size(), 0};
                for (size_t k = 1; k < out[s].size(); k++) {
                    int e = out[s][k];
                    fwd.push_back({en2[e], e, s});
                    if (symmetric) {
                        L.e_gbc.push_back({e, en2[e]});
                        bt.nc++;
                    }
                }
                if (!symmetric)
                    for (int e : in[s]) {
                        L.e_gbc.push_back({e, en1[e]});
                        bt.nc++;
                    }
                if (bt.nc > 0) L.e_gb.push_back(bt);
            }
            std::stable_sort(fwd.begin(), fwd.end(), [](const F& a, const F& b) { return a.target < b.target; });
            for (size_t i = 0; i < fwd.size();) {
                size_t j = i;
                SymGemv t{fwd[i].target, (int)L.e_gfc.size(), 0};
                while (j < fwd.size() && fwd[j].target == fwd[i].target) {
                    L.e_gfc.push_back({fwd[j].edge, fwd[j].src});
                    t.nc++;
                    j++;
                }
                L.e_gf.push_back(t);
                i = j;
            }
        }
        lap(2);
        // set_eliminated (cluster.cpp:32-45): the edges of the eliminated clusters leave the lists of their other end
        // (every touched cluster edits its own lists: host threads)
        {
            std::vector<int> touched;
            if (edge_dead.size() < en1.size()) edge_dead.resize(en1.size(), 0);
            for (int s : L.E) {
                for (int e : out[s]) {
                    edge_dead[e] = 1;
                    if (en2[e] != s) touched.push_back(en2[e]);
                }
                for (int e : in[s]) {
                    edge_dead[e] = 1;
                    touched.push_back(en1[e]);
                }
                eliminated[s] = 1;
            }
            auto is_dead = [&](int e) { return edge_dead[e] != 0; };
            for (int s : L.E) {
                out[s].clear();
                in[s].clear();
            }
            std::sort(touched.begin(), touched.end());
            touched.erase(std::unique(touched.begin(), touched.end()), touched.end());
            parallel_chunks(touched.size(), [&](int, size_t b0, size_t e0) {
                for (size_t x = b0; x < e0; x++) {
                    const int n = touched[x];
                    if (eliminated[n]) continue;
                    auto& i_ = in[n];
                    i_.erase(std::remove_if(i_.begin(), i_.end(), is_dead), i_.end());
                    if (!symmetric) {
                        auto& o = out[n];
                        o.erase(std::remove_if(o.begin(), o.end(), is_dead), o.end());
                    }
                }
            }, 4096);
        }
        lap(3);
        // ---------------- scale ----------------
        for (int c : bottom) {
            if (eliminated[c] || cl[c].level <= l) continue;
            L.S.push_back(c);
            L.s_piv.push_back(out[c][0]);
        }
        {
            std::vector<size_t> off(L.S.size() + 1, 0);
            for (size_t i = 0; i < L.S.size(); i++) off[i + 1] = off[i] + out[L.S[i]].size() - 1;
            L.s_right.resize(off.back());
            L.s_left.resize(off.back());
            parallel_chunks(L.S.size(), [&](int, size_t b0, size_t e0) {
                for (size_t i = b0; i < e0; i++) {
                    const int c = L.S[i];
                    const int piv = out[c][0];
                    size_t w = off[i];
                    for (size_t k = 1; k < out[c].size(); k++, w++) {
                        const int e = out[c][k];
                        const int c2 = en2[e];
                        L.s_right[w] = {e, piv, c2, c};
                        L.s_left[w] = {e, out[c2][0], c, c2};
                    }
                }
            }, 4096);
        }
        // ---------------- sparsify: wavefronts of the Gauss-Seidel order (tree.cpp:1523-1527, :1194-1200) ----------------
        {
            int first = bottom.empty() ? 0 : bottom.front();
            int span = bottom.empty() ? 0 : bottom.back() - first + 1;
            color.assign(span, -1);
            // tasks and their source ranges in list order, sources filled by host threads
            size_t nsrc_total = 0;
            for (int s : L.S) {
                bool want = want_flag ? cl[s].sparsify : true;
                if (!want) {
                    L.ignored++;
                    continue;
                }
                SymQr t;
                t.cluster = s;
                t.src0 = (int)nsrc_total;
                t.nsrc = (int)(in[s].size() + out[s].size() - 1);
                t.color = 0;
                nsrc_total += t.nsrc;
                L.q.push_back(t);
            }
            L.qs.resize(nsrc_total);
            parallel_chunks(L.q.size(), [&](int, size_t b0, size_t e0) {
                for (size_t i = b0; i < e0; i++) {
                    const int s = L.q[i].cluster;
                    size_t w = L.q[i].src0;
                    for (int e : in[s]) L.qs[w++] = {e, en1[e], 0};
                    for (size_t k = 1; k < out[s].size(); k++) L.qs[w++] = {out[s][k], en2[out[s][k]], 1};
                }
            }, 4096);
            // colours: a task comes after the neighbours sparsified earlier in list order
            for (SymQr& t : L.q) {
                int col = 0;
                for (int k = t.src0; k < t.src0 + t.nsrc; k++) {
                    const int cn = color[L.qs[k].nbr - first];
                    if (cn >= 0) col = std::max(col, cn + 1);
                }
                t.color = col;
                color[t.cluster - first] = col;
                L.ncolors = std::max(L.ncolors, col + 1);
            }
        }
        lap(4);
        // ---------------- merge (tree.cpp:1106-1184) ----------------
        L.medge0 = L.medge1 = (int)en1.size();
        if (l < nlevels - 1) {
            const std::vector<int>& parents = bottoms[l + 1];
            struct NewEdge { int n1, n2; };
            struct Pending { int edge_old, newidx; };
            // the parents are independent (their new edges are numbered parent by parent): chunks of parents on host
            // threads with private slot tables, stitched together in parent order
            struct Part {
                std::vector<NewEdge> ne;
                std::vector<Pending> pend;  // newidx local to the chunk
            };
            std::vector<Part> parts(kMaxHostThreads);
            const int nth = parallel_chunks(parents.size(), [&](int t, size_t b0, size_t e0) {
                Part& P = parts[t];
                if (slots[t].empty()) slots[t].assign(ncl, -1);
                std::vector<int>& slot = slots[t];
                std::vector<int> tlist;
                for (size_t x = b0; x < e0; x++) {
                    const int p = parents[x];
                    tlist.clear();
                    for (int c = cl[p].child_begin; c < cl[p].child_end; c++)
                        for (int e : out[c]) {
                            int q = cl[en2[e]].parent;
                            if (slot[q] == -1) {
                                slot[q] = -2;
                                tlist.push_back(q);
                            }
                        }
                    std::sort(tlist.begin(), tlist.end());
                    slot[p] = (int)P.ne.size();
                    P.ne.push_back({p, p});  // pivot first, then by increasing order
                    for (int q : tlist)
                        if (q != p) {
                            slot[q] = (int)P.ne.size();
                            P.ne.push_back({p, q});
                        }
                    for (int c = cl[p].child_begin; c < cl[p].child_end; c++)
                        for (int e : out[c]) P.pend.push_back({e, slot[cl[en2[e]].parent]});
                    for (int q : tlist) slot[q] = -1;
                    slot[p] = -1;
                }
            }, 2048);
            size_t npend = 0;
            std::vector<size_t> pend_off(nth + 1, 0), ne_off(nth + 1, 0);
            for (int t = 0; t < nth; t++) {
                ne_off[t + 1] = ne_off[t] + parts[t].ne.size();
                pend_off[t + 1] = pend_off[t] + parts[t].pend.size();
            }
            npend = pend_off[nth];
            {   // sizes of the parents' lists (pivot + one out-edge per neighbour, in-edges counted here)
                for (int t = 0; t < nth; t++) {
                    size_t k = 0;
                    const auto& ne = parts[t].ne;
                    while (k < ne.size()) {
                        size_t k2 = k;
                        while (k2 < ne.size() && ne[k2].n1 == ne[k].n1) k2++;
                        out[ne[k].n1].reserve(k2 - k + 8);
                        k = k2;
                    }
                }
            }
            for (int t = 0; t < nth; t++)
                for (auto& n : parts[t].ne) {
                    int e = new_edge(n.n1, n.n2);
                    if (n.n1 == n.n2) out[n.n1].insert(out[n.n1].begin(), e);
                    else {
                        out[n.n1].push_back(e);
                        in[n.n2].push_back(e);
                    }
                }
            L.medge1 = (int)en1.size();
            L.m_copy.resize(npend);
            parallel_chunks((size_t)nth, [&](int, size_t b0, size_t e0) {
                for (size_t t = b0; t < e0; t++) {
                    size_t w = pend_off[t];
                    for (auto& pc : parts[t].pend)
                        L.m_copy[w++] = {pc.edge_old, L.medge0 + (int)ne_off[t] + pc.newidx, en1[pc.edge_old], en2[pc.edge_old]};
                }
            }, 2);
            parallel_chunks(parents.size(), [&](int, size_t b0, size_t e0) {
                for (size_t x = b0; x < e0; x++) {
                    const int p = parents[x];
                    for (int c = cl[p].child_begin; c < cl[p].child_end; c++) {
                        out[c].clear();
                        in[c].clear();
                    }
                }
            }, 4096);
        }
        lap(5);
    }
    if (getenv("SPAND_TIMING"))
        fprintf(stderr, "[spand] build_symbolic: elim lists %.0f ms, schur %.0f ms, solve lists %.0f ms, set_eliminated %.0f ms, "
                        "scale+sparsify %.0f ms, merge %.0f ms\n", tsec[0] * 1e3, tsec[1] * 1e3, tsec[2] * 1e3, tsec[3] * 1e3,
                tsec[4] * 1e3, tsec[5] * 1e3);
}

}  // namespace spand
