// Symbolic factorization plan of the spaND hot path.
//
// With the spaND sparsification used by every in-scope configuration (pred == true, reference
// src/tree.cpp:1292-1347) the block structure of the trailing matrix never depends on the numerical values:
// which clusters are eliminated at a level, which fill-in edges gemm_edges creates (src/tree.cpp:761-772) and in
// which order, what every merge copies where (src/tree.cpp:1133-1184). Only the block *sizes* (the ranks) are data
// dependent. The plan below is therefore computed once per (ordering, block pattern) by replaying the reference's
// list manipulations on integers only, in the reference's loop order, and every numerical phase becomes a fixed
// sequence of batches over flat id arrays. The arrays are uploaded once; the kernels resolve ids through the
// device tables (edge -> pointer / ld, cluster -> current size) at run time.
#pragma once
#include <cstdint>
#include <vector>

namespace spand {

struct SymTrsm {   // B <- B op(T)^-1 or op(T)^-1 B
    int eB, eT;    // edge of the block, edge of the triangle (a pivot)
    int cm, cn;    // cluster giving the free dimension of B, cluster of the triangle
};
struct SymGemm {   // C = (0|C) - sum_c A_c op(B_c)
    int target, c0, nc, flags;
};
struct SymCon {
    int e1, e2;  // A = block e1, B = block e2; inner dimension = size of cluster n1(e1)
};
struct SymGemv {   // y(cluster) -= sum_c op(A_c) x_c
    int cluster, c0, nc;
};
struct SymGemvCon {
    int edge, xcluster;
};
struct SymQr {
    int cluster, src0, nsrc, color;
};
struct SymQrSrc {
    int edge, nbr, transposed;
};
struct SymCopy {   // child block -> parent block at (pos[c2], pos[c1])
    int eold, enew;
    int c1, c2;    // column / row child cluster of the old block; eold == pivot <=> c1 == c2
};

struct SymLevel {
    // eliminate (src/tree.cpp:895-967)
    std::vector<int> E, e_piv;
    std::vector<SymTrsm> e_out;  // LLT: B L^-T ; PLU: B U^-1
    std::vector<SymTrsm> e_in;   // PLU only: L^-1 P^T B
    int fill0 = 0, fill1 = 0;    // ids of the fill-in edges created at this level
    std::vector<SymGemm> e_gemm;
    std::vector<SymCon> e_con;
    std::vector<SymGemv> e_gf, e_gb;  // recorded Gemm* operations, forward / backward
    std::vector<SymGemvCon> e_gfc, e_gbc;
    // scale (src/tree.cpp:796-856): every remaining cluster and every remaining off-diagonal block
    std::vector<int> S, s_piv;
    std::vector<SymTrsm> s_right, s_left;  // same blocks, same order: column-side then row-side solve
    // sparsify (src/tree.cpp:1417-1433)
    std::vector<SymQr> q;
    std::vector<SymQrSrc> qs;
    int ncolors = 0, ignored = 0;
    // merge (src/tree.cpp:1435-1445)
    int medge0 = 0, medge1 = 0;  // ids of the parent blocks created by this merge
    std::vector<SymCopy> m_copy;
};

struct SymCluster {
    int level;            // ND level at which it is eliminated
    int hlevel;           // hierarchy level at which it lives
    int parent;           // cluster id or -1
    int child_begin, child_end;
    bool sparsify;
};

struct SymbolicPlan {
    int nlevels = 0;
    bool symmetric = true;   // SPD/LLT (lower blocks only) vs GEN/PLU
    bool want_flag = true;   // use_want_sparsify
    std::vector<int> en1, en2;  // edge table: block A[rows of n2, cols of n1], ids in creation order
    int nleaf_edges = 0;
    std::vector<SymLevel> lv;
    size_t bytes() const;
};

// bottoms[h] = cluster ids living at hierarchy level h (increasing); leaf = (n1, n2) of the assembled blocks in
// assembly order (pivot first for every column cluster, then by increasing n2).
void build_symbolic(const std::vector<SymCluster>& cl, const std::vector<std::vector<int>>& bottoms,
                    const std::vector<int>& leaf_n1, const std::vector<int>& leaf_n2, bool symmetric, bool want_flag,
                    SymbolicPlan& plan);

}  // namespace spand
