// Compiles include/spaND_b200.hpp as the reference's drivers would (g++ -std=c++14, no Eigen, no CUDA headers) and runs
// the host-only part of the Tree contract through it: setters -> partition on the 5 x 5 Laplacian of the reference's
// PartitionTest.Square (tests/tests.cpp:378-412). Compute entry points must fail loudly without a GPU.
#include <cstdio>
#include <cstring>
#include <vector>

#include "spaND_b200.hpp"

namespace spaND = spaND_b200;
using namespace spaND;

int main() {
    const int n = 5, N = n * n;
    std::vector<int> colptr(N + 1, 0), rowind;
    std::vector<double> val;
    for (int j = 0; j < N; j++) {
        const int x = j % n, y = j / n;
        auto push = [&](int xx, int yy, double v) {
            if (xx >= 0 && xx < n && yy >= 0 && yy < n) {
                rowind.push_back(xx + n * yy);
                val.push_back(v);
            }
        };
        push(x, y - 1, -1.0);
        push(x - 1, y, -1.0);
        push(x, y, 4.0);
        push(x + 1, y, -1.0);
        push(x, y + 1, -1.0);
        colptr[j + 1] = (int)rowind.size();
    }
    CscView A{N, N, colptr.data(), rowind.data(), val.data()};
    std::vector<double> X(2 * N);
    for (int j = 0; j < N; j++) {
        X[2 * j] = j / n;      // linspace_nd: first coordinate slowest (src/util.cpp:488-517)
        X[2 * j + 1] = j % n;
    }
    Tree t(3);
    t.set_verb(false);
    t.set_symm_kind(SymmKind::SPD);
    t.set_scaling_kind(ScalingKind::LLT);
    t.set_tol(1e-2);
    t.set_skip(0);
    t.set_use_geo(true);
    t.set_Xcoo(2, N, X.data());
    t.set_monitor_flops(true);
    std::vector<ClusterID6> part = t.partition(A);
    if ((int)part.size() != N || t.get_N() != N || t.get_nlevels() != 3) return 1;
    int top = 0;
    for (const auto& c : part) top += c.self_lvl == 2;
    std::vector<int> perm = t.get_assembly_perm();
    std::vector<char> seen(N, 0);
    for (int p : perm) {
        if (p < 0 || p >= N || seen[p]) return 2;
        seen[p] = 1;
    }
    std::printf("FACADE_OK top_separator_dofs=%d log_fields=%zu\n", top, Tree::log_field_names().size());
    // no CPU fallback: without a device assemble() must throw, with one the whole pipeline must run
    try {
        t.assemble(A);
        t.factorize();
        std::vector<double> b(N, 1.0);
        t.solve(b);
        Tree::Csc T = t.get_trailing_mat();
        std::printf("FACADE_GPU nnz=%lld trailing_nnz=%zu\n", t.nnz(), T.val.size());
    } catch (const std::exception& e) {
        std::printf("FACADE_NOGPU %s\n", e.what());
        if (!std::strstr(e.what(), "no CUDA device")) return 3;
    }
    return 0;
}
