"""torchrun worker of tests/test_mg_gpu.py: factorize one matrix sharded over the ranks and compare, on every rank,
with the same factorization done by this rank alone (single-GPU path). The sharded path runs the same kernels on
the same tasks in the same order, only on different GPUs: ranks, nnz and the solve must be bit-identical."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spand_public_b200 as S  # noqa: E402


def run(n, d, L, tol, sharded, local_rank, gen=False):
    A = S.aniso_convdiff(n) if gen else S.neglapl(n, d)
    t = S.Tree(L)
    t.set_device(local_rank)
    if gen:
        t.set_symm_kind(S.GEN)
        t.set_scaling_kind(S.PLU)
    if sharded:
        t.mg_init(dist, device=local_rank, arena_gb=float(os.environ.get("SPAND_MG_ARENA_GB", "8")))
    t.set_tol(tol)
    t.set_use_geo(True)
    t.set_Xcoo(S.linspace_nd(n, d))
    t.partition(S.symmetric_graph(A))
    t.assemble(A)
    t.factorize()
    b = S.random(A.shape[0], 2019)
    x = t.solve(b)
    x2 = t.solve(b)
    it, xc = t.gmres(A, b, 200, 100, 1e-12) if gen else t.cg(A, b, 200, 1e-12)
    return dict(A=A, b=b, x=x, x2=x2, ranks=t.stats()[2].copy(), nnz=t.nnz(), it=it, xc=xc,
                tfact=t.factorize_seconds(), tree=t)


def main():
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    out = []
    # the last one (64^3, 13 levels) is large enough for the streaming RRQR shapes (wavefronts of >= 48 panels > 300 KB)
    # negative L marks the GEN / PLU family (anisotropic convection-diffusion, GMRES): config C5's path
    for (n, d, L, tol) in [(16, 3, 6, 1e-2), (32, 2, 6, 0.0), (24, 3, 8, 1e-2), (64, 3, 13, 1e-2), (16, 3, -6, 1e-2),
                           (24, 3, -8, 1e-2)]:
        gen = L < 0
        L = abs(L)
        sh = run(n, d, L, tol, True, local_rank, gen)
        one = run(n, d, L, tol, False, local_rank, gen)
        res = float(np.linalg.norm(sh["A"] @ sh["x"] - sh["b"]) / np.linalg.norm(sh["b"]))
        dr = sh["ranks"].astype(int) - one["ranks"].astype(int)
        rec = dict(cfg=[n, d, L, tol], gen=gen, rank=rank, world=world, same_ranks=bool(np.array_equal(sh["ranks"], one["ranks"])),
                   ranks_differ=int((dr != 0).sum()), ranks_maxdiff=int(abs(dr).max()), nclusters=int(len(dr)),
                   nnz_rel=float(abs(sh["nnz"] - one["nnz"]) / max(1, one["nnz"])), it_sharded=int(sh["it"]),
                   it_single=int(one["it"]),
                   cg_x_rel=float(np.linalg.norm(sh["xc"] - one["xc"]) / np.linalg.norm(one["xc"])),
                   same_nnz=bool(sh["nnz"] == one["nnz"]), same_x=bool(np.array_equal(sh["x"], one["x"])),
                   repeat_x=bool(np.array_equal(sh["x"], sh["x2"])), same_cg=bool(sh["it"] == one["it"]),
                   same_cg_x=bool(np.array_equal(sh["xc"], one["xc"])), residual=res,
                   maxdiff=float(np.abs(sh["x"] - one["x"]).max()), t_sharded=sh["tfact"], t_single=one["tfact"])
        out.append(rec)
        del sh, one
        dist.barrier()
    # error propagation: an indefinite matrix (non-SPD pivot on whichever rank owns the offending cluster) must make
    # EVERY rank fail at the same point instead of leaving the others in a peer barrier
    import scipy.sparse as sp
    A = (S.neglapl(16, 3) - 5.5 * sp.identity(16**3, format="csc")).tocsc()
    t = S.Tree(6)
    t.set_device(local_rank)
    t.mg_init(dist, device=local_rank, arena_gb=float(os.environ.get("SPAND_MG_ARENA_GB", "8")))
    t.set_tol(1e-2)
    t.set_use_geo(True)
    t.set_Xcoo(S.linspace_nd(16, 3))
    t.partition(S.symmetric_graph(A))
    t.assemble(A)
    msg = ""
    try:
        t.factorize()
    except RuntimeError as ex:
        msg = str(ex)
    out.append(dict(cfg=[16, 3, 6, -1.0], error_case=True, raised="Non-SPD" in msg, message=msg.strip()))
    del t
    dist.barrier()
    allrec = [None] * world
    dist.all_gather_object(allrec, out)
    if rank == 0:
        print("MG_RESULT " + json.dumps(allrec))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
