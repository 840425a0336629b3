"""CPU, world_size 2 over gloo: the host logic of the sub-tree sharding (spand_mg_owner_map, Tree::owner_map):
every rank derives the same ownership from the same partition, merges stay local (a parent has the owner of its
children), every rank owns interiors, and the shared top separators go to distinct ranks."""
import os

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ownership(nranks, n=12, d=3, L=6):
    import spand_public_b200 as S
    A = S.neglapl(n, d)
    t = S.Tree(L)
    t.set_use_geo(True)
    t.set_Xcoo(S.linspace_nd(n, d))
    t.partition(A)
    t.plan_analyze(A)
    ids, sizes, _ = t.stats()
    starts, hlev = t.cluster_layout()
    own = t.mg_owner_map(nranks)
    return t, ids, sizes, starts, hlev, own


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _, ids, sizes, starts, hlev, own = _ownership(world)
    gathered = [None] * world
    dist.all_gather_object(gathered, own.tobytes())
    same = all(g == gathered[0] for g in gathered)
    mine = int(sizes[(own == rank) & (hlev == 0)].sum())
    q.put((rank, same, mine, int(sizes[hlev == 0].sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_all_ranks_derive_the_same_ownership():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, 29541, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = sorted(q.get(timeout=180) for _ in range(2))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][1] and got[1][1]
    total = got[0][3]
    assert got[0][2] + got[1][2] == total            # every leaf dof has exactly one owner
    assert min(got[0][2], got[1][2]) > 0.35 * total  # the two sub-trees are balanced


def test_ownership_properties():
    for nranks in (2, 4, 8):
        t, ids, sizes, starts, hlev, own = _ownership(nranks)
        assert own.min() == 0 and own.max() == nranks - 1 and len(set(own.tolist())) == nranks
        # the owner is a function of the separator the cluster belongs to (SepID self: level, index), hence a
        # parent has the owner of its children: merges and solution-segment copies are local
        perm = t.get_assembly_perm()
        self_lvl, self_sep = t.partition_ids()[0], t.partition_ids()[1]
        L = int(hlev.max()) + 1
        g = int(np.log2(nranks))
        Ls = L - 1 - g
        for c in range(len(ids)):
            dof = perm[starts[c]]
            lvl, sep = int(self_lvl[dof]), int(self_sep[dof])
            want = sep >> (Ls - lvl) if lvl <= Ls else (sep << (lvl - Ls)) + (1 << (lvl - Ls - 1))
            assert own[c] == want
        top = own[hlev == L - 1]
        assert len(top) == 1 and top[0] == nranks // 2
        # the blocks alive after the last merge before the top belong to few ranks, those of level 0 to all
        n1, _ = t.plan_live_edges(-1, 0)
        assert len(set(own[n1].tolist())) == nranks
    # nranks = 1: everything on rank 0
    assert _ownership(1)[5].max() == 0
