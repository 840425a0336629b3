// Host-side modified nested dissection + ordering + cluster hierarchy.
//
// Restates (does not copy) the integer logic of the reference that north_star keeps on the
// host "bit-exact" (citations relative to /root/reference):
//   SepID / ClusterID / merge_if          include/partition.h:23-75, src/partition.cpp:8-16
//   partition_modifiedND                  src/partition.cpp:384-478
//   partition_nd_geo (geometric bisection) src/partition.cpp:139-210
//   partition_metis (vertex separator)    src/partition.cpp:51-97
//   ordering by L stable sorts, leaf clusters, hierarchy   src/tree.cpp:344-415
#pragma once
#include <vector>

#include "sparse.hpp"

namespace spand {

struct SepID {
    int lvl = -1, sep = 0;
    SepID() {}
    SepID(int l, int s) : lvl(l), sep(s) {}
    bool operator==(const SepID& o) const { return lvl == o.lvl && sep == o.sep; }
    bool operator<(const SepID& o) const { return lvl < o.lvl || (lvl == o.lvl && sep < o.sep); }
};

struct ClusterID {
    SepID self, l, r;
    ClusterID() {}
    ClusterID(SepID s, SepID l_, SepID r_) : self(s), l(l_), r(r_) {}
    bool operator==(const ClusterID& o) const { return self == o.self && l == o.l && r == o.r; }
    bool operator<(const ClusterID& o) const {
        return (self < o.self) || (self == o.self && l < o.l) || (self == o.self && l == o.l && r < o.r);
    }
};

inline SepID merge_sep(const SepID& s) { return SepID(s.lvl + 1, s.sep / 2); }
inline ClusterID merge_if(const ClusterID& c, int lvl) {
    return ClusterID(c.self, c.l.lvl < lvl ? merge_sep(c.l) : c.l, c.r.lvl < lvl ? merge_sep(c.r) : c.r);
}

// Xcoo == nullptr -> algebraic (METIS vertex separator); else geometric (dim x N, col-major).
std::vector<ClusterID> partition_modifiedND(const SpMat& A, int nlevels, const DenseMat* Xcoo, bool verb);

// One node of the cluster hierarchy as created by Tree::partition.
struct ClusterNode {
    int start;      // first row/col in the permuted matrix
    int size;       // number of dofs (uncompressed)
    int level;      // ND level at which it is eliminated
    int order;      // global unique id: leaves first, then level-1 parents, ...
    bool sparsify;  // l.lvl == r.lvl == hierarchy level  (tree.cpp:365,406)
    int parent;     // index into levels[h+1], or -1
    int child_begin, child_end;  // range into levels[h-1] (children are consecutive), empty for leaves
    ClusterID id;
};

struct Ordering {
    int N = 0, nlevels = 0;
    std::vector<ClusterID> part;  // per dof, natural ordering
    std::vector<int> perm;        // permuted index -> natural index
    std::vector<std::vector<ClusterNode>> levels;  // levels[h] = clusters existing at hierarchy level h
    int norders = 0;
};

Ordering build_ordering(const SpMat& A, int nlevels, const DenseMat* Xcoo, bool verb);

}  // namespace spand
