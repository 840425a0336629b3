"""CPU tests of the symbolic plan (spand_public_b200/csrc/symbolic.cpp): the block structure it predicts for every
(level, phase) must be the structure the reference algorithm produces. The checker is the oracle's trailing matrix
at the same stop point (reference Tree::get_trailing_mat, src/tree.cpp:1730-1763): in exact configurations
(tol = 0: no cluster shrinks) its non-zero pattern is the union of the live dense blocks. No GPU is needed: the plan
is built on the host only (spand_plan_analyze)."""
import numpy as np
import pytest

import oracle_lib as O
import spand_public_b200 as S


def _blocks_of_trailing(T, starts, sizes):
    """Set of (column cluster, row cluster) index pairs (indices into the level's cluster list) touched by T."""
    T = T.tocoo()
    ends = starts + sizes
    def owner(ix):
        k = np.searchsorted(starts, ix, side="right") - 1
        assert np.all((k >= 0) & (ix < ends[k]))
        return k
    return set(zip(owner(T.col).tolist(), owner(T.row).tolist()))


@pytest.mark.parametrize("n,d,L,gen", [(12, 2, 4, False), (8, 3, 4, False), (10, 3, 5, False), (8, 3, 4, True)])
def test_plan_structure_matches_oracle(n, d, L, gen):
    A = S.neglapl(n, d)
    if gen:
        rng = np.random.RandomState(3)
        A = A.copy()
        A.data = A.data + rng.uniform(-0.1, 0.1, A.nnz)
    X = S.linspace_nd(n, d)
    g = S.Tree(L)
    g.set_use_geo(True)
    g.set_Xcoo(X)
    if gen:
        g.set_symm_kind(S.GEN)
        g.set_scaling_kind(S.PLU)
    g.partition(S.symmetric_graph(A))
    g.plan_analyze(A)
    ids, sizes, _ = g.stats()
    starts, hlev = g.cluster_layout()
    kw = dict(symm_kind=O.GEN, scaling_kind=O.PLU) if gen else {}
    for lvl in range(L):
        for phase in (0, 3):
            if phase == 3 and lvl == L - 1:
                continue
            o = O.OracleTree(L, tol=0.0, **kw)
            o.set_coords(X)
            o.set_stop(lvl, phase)
            o.partition(S.symmetric_graph(A))
            o.assemble(A)
            o.factorize()
            h = lvl + (1 if phase == 3 else 0)          # hierarchy level of the live clusters
            sel = np.nonzero(hlev == h)[0]
            # sizes of merged clusters are only known to the oracle run: take them from its stats
            osz = o.stats()[1][sel].astype(np.int64)
            ref = _blocks_of_trailing(o.trailing_mat(), starts[sel].astype(np.int64), osz)
            n1, n2 = g.plan_live_edges(lvl, phase)
            base = ids[sel][0]
            mine = set()
            for a, b in zip(n1.tolist(), n2.tolist()):
                if osz[a - base] == 0 or osz[b - base] == 0:
                    continue
                mine.add((a - base, b - base))
                if not gen:
                    mine.add((b - base, a - base))      # symmetric kinds store the lower blocks only
            assert mine == ref, (lvl, phase, len(mine), len(ref))


def test_plan_counts_and_wavefronts():
    n, d, L = 10, 3, 5
    A = S.neglapl(n, d)
    g = S.Tree(L)
    g.set_use_geo(True)
    g.set_Xcoo(S.linspace_nd(n, d))
    g.partition(A)
    g.plan_analyze(A)
    ids, sizes, _ = g.stats()
    _, hlev = g.cluster_layout()
    total_elim = 0
    for lvl in range(L):
        c = g.plan_counts(lvl)
        total_elim += c["eliminated"]
        assert c["schur_contribs"] >= c["schur_targets"] >= 0
        assert c["scaled_clusters"] + total_elim == np.count_nonzero(hlev == lvl)   # everything alive is scaled
        assert c["rrqr_tasks"] <= c["scaled_clusters"]
        assert (c["rrqr_wavefronts"] >= 1) == (c["rrqr_tasks"] > 0)
        if lvl < L - 1:
            assert c["merged_blocks"] <= c["merge_copies"]
            # clusters alive at level lvl + 1 = the parents of the survivors
            total_elim_next_base = np.count_nonzero(hlev == lvl + 1)
            assert total_elim_next_base <= c["scaled_clusters"]
            total_elim = 0  # the next hierarchy level lists only live clusters
    assert g.plan_counts(L - 1)["scaled_clusters"] == 0


def test_plan_needs_partition_first():
    g = S.Tree(3)
    with pytest.raises(RuntimeError):
        g.plan_analyze(S.neglapl(5, 2))


def test_csc_input_need_not_be_canonical():
    """spand_partition / spand_plan_analyze take borrowed CSC arrays: rows sorted inside every column take the
    copy-only path, anything else (unsorted rows, duplicates to be summed) is canonicalised first
    (spand_public_b200/csrc/host/sparse.cpp from_csc). Both must give the same partition and the same plan."""
    A = S.neglapl(10, 2).tocsc()
    A.sort_indices()
    N = A.shape[0]
    L = S.lib()
    X = S.linspace_nd(10, 2)
    rng = np.random.RandomState(0)
    indptr = A.indptr.astype(np.int32)
    ind = A.indices.astype(np.int32).copy()
    for j in range(N):  # shuffle the rows of every column
        sl = slice(indptr[j], indptr[j + 1])
        ind[sl] = ind[sl][rng.permutation(indptr[j + 1] - indptr[j])]
    results = []
    for rowind in (A.indices.astype(np.int32), ind):
        h = L.spand_create(3)
        L.spand_set_use_geo(h, 1)
        assert L.spand_set_coords(h, 2, N, np.ascontiguousarray(X.T).ravel()) == 0
        assert L.spand_partition(h, N, indptr, np.ascontiguousarray(rowind)) == 0
        perm = np.zeros(N, dtype=np.int32)
        L.spand_get_perm(h, perm)
        assert L.spand_plan_analyze(h, N, indptr, np.ascontiguousarray(rowind)) == 0
        counts = np.zeros(12, dtype=np.int64)
        per_level = []
        for lvl in range(3):
            assert L.spand_plan_counts(h, lvl, counts) == 0
            per_level.append(counts.copy())
        results.append((perm, np.array(per_level)))
        L.spand_destroy(h)
    assert np.array_equal(results[0][0], results[1][0])
    assert np.array_equal(results[0][1], results[1][1])


def test_partition_is_independent_of_host_threads(monkeypatch):
    """The geometric bisections of one depth run on a thread pool (host/partition.cpp) and the ordering sorts the
    distinct ClusterIDs instead of the dofs: the ClusterID of every dof and the assembly permutation must not depend
    on the number of threads (1 = the sequential order of the reference, src/partition.cpp:384-478)."""
    n, d, L = 48, 3, 9  # 110 592 dofs: above the size where the pool is used
    A = S.symmetric_graph(S.neglapl(n, d))
    X = S.linspace_nd(n, d)
    results = []
    for threads in ("1", "3", "8", "16", "16", "16"):  # repeated: the schedule of the pool differs from run to run
        monkeypatch.setenv("SPAND_HOST_THREADS", threads)
        t = S.Tree(L)
        t.set_use_geo(True)
        t.set_Xcoo(X)
        t.partition(A)
        results.append((t.get_assembly_perm().copy(), np.stack(t.partition_ids())))
    for perm, ids in results[1:]:
        assert np.array_equal(perm, results[0][0])
        assert np.array_equal(ids, results[0][1])
    # the permutation is the one of the reference's ordering, restated literally: identity, then one stable sort per
    # level by the progressively merged ClusterIDs (src/tree.cpp:344-352) = a lexicographic sort whose most
    # significant key is the ClusterID merged up to the top level
    perm, ids = results[0]
    slv, ssp, llv, lsp, rlv, rsp = [ids[i].astype(np.int64) for i in range(6)]
    keys = []  # most significant first
    per_level = []
    for lvl in range(L):
        if lvl > 0:
            ml, mr = llv < lvl, rlv < lvl
            llv, lsp = np.where(ml, llv + 1, llv), np.where(ml, lsp // 2, lsp)
            rlv, rsp = np.where(mr, rlv + 1, rlv), np.where(mr, rsp // 2, rsp)
        per_level.append([slv, ssp, llv.copy(), lsp.copy(), rlv.copy(), rsp.copy()])
    for lvl in reversed(range(L)):
        keys.extend(per_level[lvl])
    ref = np.lexsort(tuple(reversed(keys)))
    assert np.array_equal(ref.astype(perm.dtype), perm)


def test_csc_duplicates_and_empty_columns():
    """from_csc (host/sparse.cpp): duplicate entries are summed, empty columns are kept, the result is what scipy's
    canonical form gives; checked through spand_util round trips on the graph used for partitioning."""
    import scipy.sparse as sp
    rng = np.random.RandomState(5)
    n = 40
    rows = rng.randint(0, n, 300)
    cols = rng.randint(0, n - 5, 300)  # the last five columns stay empty
    vals = rng.uniform(-1, 1, 300)
    A = sp.coo_matrix((vals, (rows, cols)), shape=(n, n))
    C = A.tocsc()  # scipy sums duplicates and sorts
    # hand the raw, non-canonical arrays (sorted by column only) to the library
    order = np.argsort(cols, kind="stable")
    indptr = np.zeros(n + 1, dtype=np.int32)
    np.add.at(indptr, cols + 1, 1)
    indptr = np.cumsum(indptr).astype(np.int32)
    ri = rows[order].astype(np.int32)
    L = S.lib()
    h = L.spand_create(1)
    assert L.spand_partition(h, n, indptr, np.ascontiguousarray(ri)) == 0   # one level: everything in one cluster
    assert L.spand_plan_analyze(h, n, indptr, np.ascontiguousarray(ri)) == 0
    assert L.spand_get_N(h) == n
    L.spand_destroy(h)
    assert C.has_canonical_format and C.nnz <= 300


def test_threaded_geometric_partition_matches_literal_restatement():
    """Modified nested dissection with geometric bisection (src/partition.cpp:139-210, :384-478) restated with numpy,
    sub-domain by sub-domain in the reference's sequential order, against the threaded C++ front end on a problem
    large enough for the thread pool (110 592 dofs, 9 levels, up to 128 sub-domains per depth)."""
    n, d, L = 48, 3, 9
    A = S.symmetric_graph(S.neglapl(n, d)).tocsr()
    X = S.linspace_nd(n, d)
    N = A.shape[0]
    t = S.Tree(L)
    t.set_use_geo(True)
    t.set_Xcoo(X)
    t.partition(A)
    got = np.stack(t.partition_ids()).astype(np.int64)  # self_lvl, self_sep, l_lvl, l_sep, r_lvl, r_sep
    ids = np.zeros((6, N), dtype=np.int64)
    ids[0::2] = L - 1  # (lvl, sep) = (L - 1, 0) for self, l, r
    doms = [np.arange(N)]
    for depth in range(L - 1):
        level = L - depth - 1
        new = []
        for sep, dofs in enumerate(doms):
            if len(dofs) == 0:
                new += [dofs, dofs]
                continue
            Xd = X[:, dofs]
            rng_ = Xd.max(axis=1) - Xd.min(axis=1)
            best = int(np.argmax(rng_))  # first maximal range, like the strict '>' of the reference
            xs = Xd[best]
            midv = np.partition(xs, len(dofs) // 2)[len(dofs) // 2]
            parts = (xs >= midv).astype(np.int64)
            sub = A[dofs][:, dofs]
            touches_left = (sub @ (parts == 0).astype(np.float64)) > 0
            parts[(parts == 1) & touches_left] = 2
            idself, idl, idr = (level, sep), (level - 1, 2 * sep), (level - 1, 2 * sep + 1)
            q = ids[:, dofs].copy()
            for f in (0, 2, 4):  # self, l, r move to the side the dof fell on
                m = (q[f] == idself[0]) & (q[f + 1] == idself[1])
                for side, idside in ((0, idl), (1, idr)):
                    mm = m & (parts == side)
                    q[f, mm], q[f + 1, mm] = idside
            on_sep = (parts == 2) & (ids[0, dofs] == idself[0]) & (ids[1, dofs] == idself[1])
            q[2, on_sep], q[3, on_sep] = idl
            q[4, on_sep], q[5, on_sep] = idr
            ids[:, dofs] = q
            def member(idx):
                return ((q[0] == idx[0]) & (q[1] == idx[1])) | ((q[2] == idx[0]) & (q[3] == idx[1])) | \
                       ((q[4] == idx[0]) & (q[5] == idx[1]))
            new += [dofs[member(idl)], dofs[member(idr)]]
        doms = new
    assert np.array_equal(got, ids)


def _plan_dump(n, d, L, threads):
    """All plan counts and live edge lists of one configuration, built in a fresh process with a given thread cap."""
    import json, os, subprocess, sys
    code = f"""
import json, sys
sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
import numpy as np
import spand_public_b200 as S
A = S.neglapl({n}, {d}); X = S.linspace_nd({n}, {d})
g = S.Tree({L}); g.set_use_geo(True); g.set_Xcoo(X); g.partition(A); g.plan_analyze(A)
out = []
for lvl in range({L}):
    out.append([[k, int(v)] for k, v in sorted(g.plan_counts(lvl).items())])
    for phase in (0, 3):
        if phase == 3 and lvl == {L} - 1: continue
        n1, n2 = g.plan_live_edges(lvl, phase)
        out.append([int(np.asarray(n1, dtype=np.int64).sum()), int(np.asarray(n2, dtype=np.int64).sum()), len(n1),
                    int((np.asarray(n1, dtype=np.int64) * (np.arange(len(n1)) % 1009)).sum()),
                    int((np.asarray(n2, dtype=np.int64) * (np.arange(len(n2)) % 1013)).sum())])
print(json.dumps(out))
"""
    env = dict(os.environ)
    if threads is None:
        env.pop("SPAND_HOST_THREADS", None)
    else:
        env["SPAND_HOST_THREADS"] = str(threads)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_threaded_planner_equals_serial():
    """Leaf structure, value map and the per-level lists of build_symbolic are built by chunks on host threads
    (host/parallel.hpp); the plan must not depend on the thread count. 48^3 / 12 levels is large enough for every
    threaded section to split (10^4-10^5 clusters per level)."""
    serial = _plan_dump(48, 3, 12, 1)
    assert serial == _plan_dump(48, 3, 12, None)
    assert serial == _plan_dump(48, 3, 12, 3)
