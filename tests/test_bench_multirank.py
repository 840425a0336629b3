"""CPU, world_size 2 over gloo: the N>1 bookkeeping of bench.py (max-over-ranks timing, whole-job throughput,
rank-0-only printing) and the reference arm's contract under torchrun."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # each rank "factorizes" its own replica: time differs per rank, the job time is the max
    t_local = torch.tensor([1.0 + rank, 2.0 + 3 * rank], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
    units = 1000 * 3 * world
    q.put((rank, units / float(t_local[0]), float(t_local[1])))
    dist.barrier()
    dist.destroy_process_group()


def test_max_over_ranks_and_aggregate_units():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29531
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    # both ranks agree on the job-level numbers: 6000 units / max(1,2) s ; e2e max = 5 s
    assert got[0][1] == got[1][1] == 3000.0
    assert got[0][2] == got[1][2] == 5.0


def test_reference_arm_prints_one_line_on_rank0_only():
    env = dict(os.environ, OPENBLAS_NUM_THREADS="2")
    outs = []
    for rank in (0, 1):
        e = dict(env, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                            "--steps", "1", "--warmup", "0", "--config", "c1"], capture_output=True, text=True, env=e,
                           timeout=600)
        assert r.returncode == 0, r.stderr
        outs.append(r.stdout.strip())
    assert outs[1] == ""
    line = json.loads(outs[0].splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "factorize_throughput" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
