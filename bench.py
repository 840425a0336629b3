#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 spaND factorization path.

One "step" = one numerical factorization (Tree::factorize, reference src/tree.cpp:1447-1551) of the
configuration BASELINE.json quotes its metric on: the synthetic 3-D 7-point Laplacian 128^3
(2 097 152 dofs), tol 1e-2, 16 levels, geometric modified ND (config C4). Blocks are already assembled
in HBM when the timed region starts (`value`); `e2e` times assemble (host CSC -> dense blocks -> HBM)
+ factorize + one solve with host vectors through the C ABI.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c4|c3|c2|c1|c5|c5s|n,d,L,tol]

N > 1 (torchrun): the ranks factorize ONE matrix together, sharded by nested-dissection sub-trees; blocks of
shared separators are read / written through peer-mapped memory over NVLink between peer barriers
(DESIGN.md section 6). `value` = dofs of that matrix / max-over-ranks device time ("scaling": "strong").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CONFIGS = {
    # name: (n, d, nlevels, tol, description)
    "c1": (32, 2, 5, 1e-2, "2D Laplacian 32^2 (mats/neglapl_2_32.mm), tol 1e-2, 5 levels"),
    "c2": (30, 3, 8, 1e-2, "3D Laplacian 30^3 (mats/neglapl_3_30.mm), tol 1e-2, 8 levels"),
    "c3": (1024, 2, 14, 1e-3, "2D Laplacian 1024^2, tol 1e-3, 14 levels"),
    "c4": (128, 3, 16, 1e-2, "3D Laplacian 128^3, tol 1e-2, 16 levels"),
    "s96": (96, 3, 15, 1e-2, "3D Laplacian 96^3, tol 1e-2, 15 levels"),
    "s80": (80, 3, 14, 1e-2, "3D Laplacian 80^3, tol 1e-2, 14 levels"),
    "s64": (64, 3, 13, 1e-2, "3D Laplacian 64^3, tol 1e-2, 13 levels"),
    "s48": (48, 3, 12, 1e-2, "3D Laplacian 48^3, tol 1e-2, 12 levels"),
    # config C5: non-symmetric, GEN + PLU, GMRES(100); 16 = round(log2(N / 64)) (tests/spaND.cpp:137-140)
    "c5": (160, 3, 16, 1e-2, "3D anisotropic variable-coefficient convection-diffusion 160^3 (non-symmetric, GEN+PLU), "
                             "tol 1e-2, 16 levels"),
    "c5s": (96, 3, 14, 1e-2, "3D anisotropic variable-coefficient convection-diffusion 96^3 (non-symmetric, GEN+PLU), "
                             "tol 1e-2, 14 levels"),
    "a48": (48, 3, 11, 1e-2, "3D anisotropic variable-coefficient convection-diffusion 48^3 (non-symmetric, GEN+PLU), "
                             "tol 1e-2, 11 levels"),
}
ANISO = ("c5", "c5s", "a48")  # matrix family of these configurations: S.aniso_convdiff, GEN / PLU / GMRES


# ---- inputs of the CPU (reference) arm, generated with numpy / scipy only: that process maps oracle/_build/liboracle.so
# and nothing of the product (the generators are checked against spand_util_* in tests/test_bench_inputs.py) ----
def np_neglapl(n, d):
    """mats/neglapl_d_n.mm: Dirichlet stencil Laplacian, diagonal 2 d, dof index x + n y (+ n^2 z)."""
    import scipy.sparse as sp
    T = sp.diags([-np.ones(n - 1), np.zeros(n), -np.ones(n - 1)], [-1, 0, 1], format="csc")
    I = sp.identity(n, format="csc")
    A = None
    for k in range(d):
        term = None
        for j in range(d):
            M = T if (d - 1 - j) == k else I
            term = M if term is None else sp.kron(term, M, format="csc")
        A = term if A is None else A + term
    A = (A + 2 * d * sp.identity(n**d, format="csc")).tocsc()
    A.sort_indices()
    return A


def np_linspace_nd(n, d):
    """src/util.cpp:488-517: first coordinate slowest."""
    X = np.zeros((d, n**d))
    r = np.arange(n**d)
    for k in range(d - 1, -1, -1):
        X[k] = r % n
        r = r // n
    return X


def np_random(size, seed):
    """src/util.cpp:549-558: std::mt19937(seed) + uniform_real_distribution<double>(-1, 1) (two 32-bit draws each)."""
    raw = np.random.RandomState(seed).randint(0, 2**32, size=2 * size, dtype=np.uint64)
    u = (raw[0::2].astype(np.float64) + raw[1::2].astype(np.float64) * 4294967296.0) / 18446744073709551616.0
    return -1.0 + 2.0 * u


def np_symmetric_graph(A):
    """src/util.cpp:47-63: |A| + |A^T| + I."""
    import scipy.sparse as sp
    G = (abs(A) + abs(A).T + sp.identity(A.shape[0], format="csc")).tocsc()
    G.sort_indices()
    return G


def np_aniso_convdiff(n):
    """Config C5: 7-point finite volumes of -div(kappa diag(1, .1, .01) grad u) + (1,1,1).grad u,
    kappa = 10^(sin 2 pi x sin 2 pi y sin 2 pi z) at cell centres, harmonic face averages, first-order upwind."""
    import scipy.sparse as sp
    h = 1.0 / n
    c = (np.arange(n) + 0.5) * h
    sx = np.sin(2 * np.pi * c)
    kap = 10.0 ** (sx[None, None, :] * sx[None, :, None] * sx[:, None, None])  # [k, j, i]
    eps = (1.0, 1e-1, 1e-2)
    N = n**3
    idx = np.arange(N).reshape(n, n, n)  # row = i + n j + n^2 k
    diag = np.zeros((n, n, n))
    rows, cols, vals = [], [], []
    for d, ax in ((0, 2), (1, 1), (2, 0)):  # direction d of the grid = axis `ax` of the [k, j, i] arrays
        for sgn in (-1, 1):
            sl_c = [slice(None)] * 3
            sl_n = [slice(None)] * 3
            sl_c[ax] = slice(1, None) if sgn < 0 else slice(None, -1)
            sl_n[ax] = slice(None, -1) if sgn < 0 else slice(1, None)
            sl_c, sl_n = tuple(sl_c), tuple(sl_n)
            ka, kb = kap[sl_c], kap[sl_n]
            kf = eps[d] * 2.0 * ka * kb / (ka + kb)
            diag[sl_c] += kf
            off = -kf - (1.0 if sgn < 0 else 0.0)
            rows.append(idx[sl_c].ravel())
            cols.append(idx[sl_n].ravel())
            vals.append(off.ravel())
            bd = [slice(None)] * 3
            bd[ax] = slice(0, 1) if sgn < 0 else slice(n - 1, n)
            bd = tuple(bd)
            diag[bd] += eps[d] * kap[bd]  # Dirichlet ghost cell
        diag += 1.0  # upwind convection, one per direction
    rows.append(idx.ravel())
    cols.append(idx.ravel())
    vals.append(diag.ravel())
    A = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N))
    A.sort_indices()
    return A


def np_matrix_of(cfg):
    return np_aniso_convdiff(cfg[0]) if is_aniso(cfg) else np_neglapl(cfg[0], cfg[1])


def golden_of(name):
    """One full run of the CPU oracle on a BASELINE-size configuration, committed with its generating script."""
    import gzip
    fn = os.path.join(ROOT, "tests", "golden", f"{name}_oracle.json.gz")
    if not os.path.exists(fn):
        return None
    with gzip.open(fn, "rt") as f:
        g = json.load(f)
    return {k: g[k] for k in ("config", "N", "iterations", "solver", "residual_one_solve", "nnz", "gflop", "host")}


def is_aniso(cfg):
    return "convection-diffusion" in cfg[4]


def matrix_of(S, cfg):
    return S.aniso_convdiff(cfg[0]) if is_aniso(cfg) else S.neglapl(cfg[0], cfg[1])


def parse_config(name):
    if name in CONFIGS:
        return CONFIGS[name]
    n, d, L, tol = name.split(",")
    return (int(n), int(d), int(L), float(tol), f"{d}D Laplacian {n}^{d}, tol {tol}, {L} levels")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def measure_fp64_peak(torch, dev):
    """DGEMM ceiling (cuBLAS 8192^3, best of 5): MEASURED_PEAKS.json holds no FP64 figure."""
    try:
        a = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
        b = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
        best = 1e9
        for _ in range(6):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            torch.matmul(a, b)
            e.record()
            torch.cuda.synchronize()
            best = min(best, s.elapsed_time(e) * 1e-3)
        del a, b
        torch.cuda.empty_cache()
        return 2 * 8192**3 / best / 1e12
    except Exception:
        return None


def flops_of(log):
    return float(sum(log[k].sum() for k in ("fl_pivot", "fl_panel", "fl_schur", "fl_rrqr_rank")))


def run_oracle(cfg, steps, warmup, threads):
    """CPU arm: the oracle's factorize() timed exactly where the reference driver times it
    (tests/spaND.cpp:274-292 -> <<<<tfact)."""
    import oracle_lib as O
    n, d, L, tol, desc = cfg
    O.lib().orc_set_threads(threads)
    A = np_matrix_of(cfg)
    X = np_linspace_nd(n, d)
    G = np_symmetric_graph(A)
    gen = is_aniso(cfg)
    times = []
    info = {}
    for it in range(warmup + steps):
        t = O.OracleTree(L, tol=tol, symm_kind=O.GEN, scaling_kind=O.PLU) if gen else O.OracleTree(L, tol=tol)
        t.set_coords(X)
        t.partition(G)
        t.assemble(A)
        t0 = time.perf_counter()
        t.factorize()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        if it == warmup + steps - 1:
            b = np_random(A.shape[0], 2019)
            ts = time.perf_counter()
            x = t.solve(b)
            info["tsolve_s"] = time.perf_counter() - ts
            info["residual"] = float(np.linalg.norm(A @ x - b) / np.linalg.norm(b))
            if gen:
                info["gmres_iterations"] = int(t.gmres(A, b, 500, 100, 1e-12)[0])
            else:
                info["cg_iterations"] = int(t.cg(A, b, 500, 1e-12)[0])
            info["gflop"] = flops_of(t.log()) / 1e9
            info["nnz_fact"] = int(t.nnz())
        del t
    return A.shape[0], times, info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c4")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cg", action="store_true")
    ap.add_argument("--algebraic", action="store_true",
                    help="algebraic modified ND (METIS vertex separators, src/partition.cpp:51-97) instead of the "
                         "geometric one: the reference's default when no coordinates are given (SURVEY 8d, config C2)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = parse_config(args.config)
    n, d, L, tol, desc = cfg
    metric, unit = "factorize_throughput", "Mdof/s"

    if args.impl == "reference":
        # The reference's own CPU path = the oracle port (the reference cannot be compiled here).
        if rank != 0:
            return 0
        ncpu = os.cpu_count() or 1
        # The reference is sequential C++ over BLAS/LAPACK (tests/Makefile links mkl_sequential); extra BLAS threads
        # can only help inside the few large blocks and hurt on the thousands of tiny ones. Probe both settings on
        # config C2 and time the sample with the faster one, so that the CPU arm gets its best case.
        probe = {}
        for th in sorted({1, ncpu}):
            _, tt, _ = run_oracle(CONFIGS["c2"], 1, 0, th)
            probe[th] = tt[0]
        cores = min(probe, key=probe.get)
        # Like for like: the workload itself whenever (steps + warmup) repetitions of it fit the time budget of this
        # arm (SPAND_REF_BUDGET_S, default 900 s), otherwise the largest member of the same family that does. Cost
        # of one factorize() relative to C2, measured with the oracle on two hosts (tests/golden/*_oracle.json.gz:
        # 64^3 14.9 s, 128^3 176 s where C2 takes 0.6 s; the GPU boxes run the same ratios 2.3x faster).
        rel = {"c1": 0.02, "c2": 1.0, "c3": 8.0, "s48": 8.0, "s64": 25.0, "s80": 55.0, "s96": 105.0, "c4": 300.0,
               "a48": 14.0, "c5s": 230.0, "c5": 1500.0}
        budget = float(os.environ.get("SPAND_REF_BUDGET_S", "900"))
        reps = max(1, args.steps + args.warmup)
        family = ["c5", "c5s", "a48"] if args.config in ANISO else (
            [args.config] if args.config in ("c1", "c2", "c3") else ["c4", "s96", "s80", "s64", "s48"])
        if args.config in family:
            family = family[family.index(args.config):]
        elif args.config not in CONFIGS:
            family = [None] + family  # ad-hoc configuration: itself first
        sample_name = family[-1]
        for name in family:
            cost = rel.get(name, 300.0) * probe[cores] * reps
            if name is None:
                cost = 300.0 * probe[cores] * reps * (cfg[0] ** cfg[1] / 128.0**3) ** 1.35
            if cost <= budget:
                sample_name = name
                break
        scfg = cfg if sample_name is None or sample_name == args.config else CONFIGS[sample_name]
        N, times, info = run_oracle(scfg, args.steps, args.warmup, cores)
        info["thread_probe_c2_seconds"] = {str(k): v for k, v in probe.items()}
        info["host_cpus"] = ncpu
        tmean = float(np.mean(times))
        val = N / tmean / 1e6
        out = {"impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": args.gpus,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": tmean * 1e3, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": desc, "sample": scfg[4], "same_as_workload": scfg[4] == desc},
               "cpu_baseline": {"value": val, "unit": unit, "cores": cores, "kind": "port",
                                "sample": f"full factorize() of {scfg[4]}"
                                          + (" (the workload itself)" if scfg[4] == desc else
                                             " (largest member of the workload's family whose repetitions fit the time "
                                             "budget of this arm; throughput in dofs/s is size-normalised and falls "
                                             "with size on the CPU)")
                                          + f", OpenBLAS threads={cores} (the faster of 1 and {ncpu} on a probe)",
                                "full_workload_golden": golden_of(args.config)},
               "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "factorize_time_s": tmean, **info}
        print(json.dumps(out))
        return 0

    import torch
    import spand_public_b200 as S
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the spaND B200 path has no CPU fallback")
    dist = None
    stdout_fd = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        # NCCL prints its version banner on fd 1 at the first collective: keep stdout for the one JSON line
        sys.stdout.flush()
        stdout_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    fp64_peak = measure_fp64_peak(torch, dev) if rank == 0 else None
    A = matrix_of(S, cfg)
    N = A.shape[0]
    X = S.linspace_nd(n, d)
    G = S.symmetric_graph(A)
    b = S.random(N, 2019)
    gen = is_aniso(cfg)
    t = S.Tree(L)
    t.set_device(local_rank)
    if gen:
        t.set_symm_kind(S.GEN)
        t.set_scaling_kind(S.PLU)
    if world > 1:
        # one factorization sharded by ND sub-trees over the ranks (blocks of shared separators reached through
        # peer memory over NVLink); every rank makes the same calls
        t.mg_init(dist, device=local_rank)
    t.set_tol(tol)
    if args.algebraic:
        t.set_use_geo(False)
    else:
        t.set_use_geo(True)
        t.set_Xcoo(X)
    tp0 = time.perf_counter()
    t.partition(G)
    tpart = time.perf_counter() - tp0
    h2d = A.data.nbytes + A.indices.nbytes + A.indptr.nbytes + b.nbytes
    d2h = b.nbytes

    tcold0 = time.perf_counter()

    def step():
        t0 = time.perf_counter()
        t.assemble(A)
        t1 = time.perf_counter()
        t.factorize()
        t2 = time.perf_counter()
        x = t.solve(b)
        t3 = time.perf_counter()
        return t.factorize_seconds(), t3 - t0, t1 - t0, t3 - t2, x

    # cold end-to-end time: partition + symbolic analysis (inside the first assemble) + assemble + factorize + solve
    step()
    torch.cuda.synchronize()
    e2e_cold = tpart + (time.perf_counter() - tcold0)
    for _ in range(max(0, args.warmup - 1)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    dev_times, e2e_times, tassm, tsolve = [], [], [], []
    launches = 0
    for _ in range(args.steps):
        ft, et, ta, ts, x = step()
        dev_times.append(ft)
        e2e_times.append(et)
        tassm.append(ta)
        tsolve.append(ts)
        launches += t.kernel_launches()
    barrier()
    clocks = sampler.stop()
    tdev = float(np.sum(dev_times))
    te2e = float(np.sum(e2e_times))
    if dist is not None:
        tt = torch.tensor([tdev, te2e], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        tdev, te2e = float(tt[0]), float(tt[1])
    lg = t.log()
    res = float(np.linalg.norm(A @ x - b) / np.linalg.norm(b))
    # one more factorization, outside the timed region, with CUDA events around every launch of every kernel family
    # (on the factorization stream): per-family kernel time for the roofline object
    t.set_profile(True)
    t.assemble(A)
    t.factorize()
    fam = t.family_stats()
    t.set_profile(False)
    arena_b = float(t.arena_bytes())
    if dist is not None:
        # sharded: a family's time is the slowest rank's (every rank works on its share of the same batches); the
        # bytes / flops of the roofline are those of the whole matrix
        names = sorted(fam)
        ft = torch.tensor([fam[k][0] for k in names] + [-arena_b], dtype=torch.float64, device=dev)
        dist.all_reduce(ft, op=dist.ReduceOp.MAX)
        fam = {k: (float(ft[i]), fam[k][1]) for i, k in enumerate(names)}
        ab = torch.tensor([arena_b], dtype=torch.float64, device=dev)
        dist.all_reduce(ab, op=dist.ReduceOp.SUM)
        arena_b = float(ab[0])
    cg_it, t_cg = None, None
    if not args.no_cg:  # collective when sharded: every rank runs the same PCG around the distributed solve
        if gen:  # tests/spaND.cpp:312-322: GMRES for non-symmetric problems
            cg_it, _ = t.gmres(A, b, 500, 100, 1e-12)
            t_cg = t.t_gmres
        else:
            cg_it, _ = t.cg(A, b, 500, 1e-12)
            t_cg = t.t_cg
    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # the bench line carries its own correctness check (at every GPU count): one-solve residual within the reference's
    # ApproxTest bound (tests/tests.cpp:799-856) and the Krylov iteration count against the committed oracle golden
    gold = None if args.algebraic else golden_of(args.config)  # the goldens are geometric-partition runs
    check = {"residual_one_solve": res, "residual_bound": 200 * tol, "iterations": cg_it,
             "oracle_iterations": gold["iterations"] if gold else None,
             "oracle_residual_one_solve": gold["residual_one_solve"] if gold else None}
    check["ok"] = bool(res <= 200 * tol and (gold is None or cg_it is None or abs(cg_it - gold["iterations"]) <= 1))
    if not check["ok"]:
        raise SystemExit("bench.py: correctness check failed: " + json.dumps(check))
    units = N * args.steps  # sharded: the ranks factorize ONE matrix together (strong scaling)
    value = units / tdev / 1e6
    e2e_val = units / te2e / 1e6
    flops = flops_of(lg)
    pk = peaks()
    hbm_peak = pk.get("hbm_gbs", 6650.0)
    hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in pk else "fallback (B200_PROFILING.md)"
    fp64_peak_1gpu = fp64_peak
    if world > 1:
        # sharded: the bytes / flops below are those of the whole matrix and the family times the slowest rank's, so
        # `achieved` is an aggregate over the ranks: it is compared with the aggregate peak of the N GPUs
        hbm_peak *= world
        hbm_src += " x %d GPUs (aggregate)" % world
        if fp64_peak:
            fp64_peak *= world
    # Dominant kernel family and its roofline. Algorithmic bytes / flops are SURVEY.md 8(d)'s per-unit figures summed
    # over the factorization; the duration is the family's kernel time from CUDA events around each of its launches.
    fam_s = {k: v[0] * 1e-3 for k, v in fam.items()}
    dom = max(fam_s, key=fam_s.get)
    by = {"rrqr": lg["by_rrqr"].sum(), "trsm": lg["by_scale"].sum(), "copy": lg["by_merge"].sum()}
    kern = {"rrqr": "rrqr_blocked_kernel / rrqr_hc_kernel / rrqr_hc2_kernel (gather + truncated QRCP + scatter)",
            "trsm": "scale_sym/trsm_strip kernels (two-sided scaling + panels)", "copy": "copy_sym_kernel + memset",
            "gemm": "gemm_tiled/gemm_sym kernels (Schur updates)", "potrf": "potrf kernels"}
    if dom in ("gemm", "potrf"):
        fl = float(lg["fl_schur"].sum() if dom == "gemm" else lg["fl_pivot"].sum())
        roof = {"bound": "tensor", "kernel": kern[dom], "achieved": fl / fam_s[dom] / 1e12, "peak": fp64_peak or 40.0,
                "unit": "TFLOP/s",
                "peak_source": ("cuBLAS DGEMM 8192^3 measured in this run (no FP64 figure in MEASURED_PEAKS.json)"
                                + (" x %d GPUs (aggregate)" % world if world > 1 else "")
                                if fp64_peak else "nominal 40 TFLOP/s (vendor)")}
    else:
        roof = {"bound": "hbm", "kernel": kern[dom], "achieved": float(by[dom]) / fam_s[dom] / 1e9, "peak": hbm_peak,
                "unit": "GB/s", "peak_source": hbm_src}
    roof["frac"] = roof["achieved"] / roof["peak"] if roof["peak"] else None
    roof["traffic"] = None
    try:
        # DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch of the dominant kernel family, from the
        # ncu capture of this same command committed under profiles/ (scripts/ncu_traffic.py); averaged over the
        # launches of the family like `achieved`
        cand = [os.path.join(ROOT, "profiles", f) for f in ("r2c_ncu_traffic.json", "r2_ncu_traffic.json")]
        tr = json.load(open([f for f in cand if os.path.exists(f)][0]))
        if dom in tr.get("families", {}):
            roof["traffic"] = tr["families"][dom]["dram_bytes_per_launch"]
            roof["traffic_source"] = tr.get("source")
            roof["algorithmic_bytes_per_launch"] = float(by.get(dom, 0.0)) / max(1, fam[dom][1])
    except Exception:
        pass
    roof["launches"] = int(fam[dom][1])
    roof["avg_launch_ms"] = fam[dom][0] / max(1, fam[dom][1])
    roof["family_kernel_seconds"] = fam_s
    if dom == "rrqr":
        # what the kernel actually streams: every Householder step reads the active part of the panel once
        # (BLAS-2 bound, 2 bytes per flop); the compulsory-traffic model above assumes the panel stays on chip
        roof["streamed_model"] = {"bytes": 2.0 * float(lg["fl_rrqr_rank"].sum()),
                                  "achieved_gbs": 2.0 * float(lg["fl_rrqr_rank"].sum()) / fam_s[dom] / 1e9,
                                  "frac_of_hbm": 2.0 * float(lg["fl_rrqr_rank"].sum()) / fam_s[dom] / 1e9 / hbm_peak}
    gemm_fl = float(lg["fl_schur"].sum())
    roof["gemm_family"] = {"bound": "tensor", "achieved": gemm_fl / max(fam_s["gemm"], 1e-9) / 1e12,
                           "peak": fp64_peak, "unit": "TFLOP/s",
                           "frac": (gemm_fl / max(fam_s["gemm"], 1e-9) / 1e12 / fp64_peak) if fp64_peak else None}
    # every family against its own bound (algorithmic work of SURVEY.md 8(d) / kernel time of the family)
    roof["families"] = {
        "rrqr": {"bound": "hbm", "achieved_gbs": float(by["rrqr"]) / max(fam_s["rrqr"], 1e-9) / 1e9,
                 "frac": float(by["rrqr"]) / max(fam_s["rrqr"], 1e-9) / 1e9 / hbm_peak},
        "trsm": {"bound": "hbm", "achieved_gbs": float(by["trsm"]) / max(fam_s["trsm"], 1e-9) / 1e9,
                 "frac": float(by["trsm"]) / max(fam_s["trsm"], 1e-9) / 1e9 / hbm_peak,
                 "tflops": float(lg["fl_panel"].sum()) / max(fam_s["trsm"], 1e-9) / 1e12},
        "copy": {"bound": "hbm", "achieved_gbs": float(by["copy"]) / max(fam_s["copy"], 1e-9) / 1e9,
                 "frac": float(by["copy"]) / max(fam_s["copy"], 1e-9) / 1e9 / hbm_peak},
        "gemm": {"bound": "tensor", "achieved_tflops": gemm_fl / max(fam_s["gemm"], 1e-9) / 1e12,
                 "frac": (gemm_fl / max(fam_s["gemm"], 1e-9) / 1e12 / fp64_peak) if fp64_peak else None},
        "potrf": {"bound": "tensor", "achieved_tflops": float(lg["fl_pivot"].sum()) / max(fam_s["potrf"], 1e-9) / 1e12,
                  "frac": (float(lg["fl_pivot"].sum()) / max(fam_s["potrf"], 1e-9) / 1e12 / fp64_peak) if fp64_peak else None},
    }
    phases = {"eliminate": lg["t_elim"].sum(), "scale": lg["t_scale"].sum(), "sparsify": lg["t_spars"].sum(),
              "merge": lg["t_merge"].sum()}
    roof["phase_seconds"] = {k: float(v) for k, v in phases.items()}

    cpu = None
    if world == 1 and not args.no_cpu_baseline and not args.algebraic:
        # bounded live sample on this box's host cores (10-30 s of CPU work); the one-off run of the oracle on the
        # full workload is the committed golden (tests/golden/, measured on the build container's host)
        scfg = CONFIGS["s80"] if args.config == "c4" else (CONFIGS["a48"] if args.config in ANISO else cfg)
        Nc, times, info = run_oracle(scfg, 1, 0, 1)
        cpu = {"value": Nc / times[0] / 1e6, "unit": unit, "cores": 1, "kind": "port",
               "sample": f"one full factorize() of {scfg[4]} by the oracle port (OpenBLAS 1 thread, as the reference's "
                         f"mkl_sequential build); {times[0]:.2f} s on this box",
               "factorize_time_s": times[0], "residual": info["residual"],
               **{k: info[k] for k in ("cg_iterations", "gmres_iterations") if k in info},
               "full_workload_golden": golden_of(args.config)}

    out = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": tdev / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak" if world == 1 else "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "N": N, "parallelism": "single GPU" if world == 1 else
                   f"one factorization sharded by ND sub-trees over {world} GPUs, peer memory over NVLink",
                   "l2": "inputs (assembled blocks and factors, ~%.1f GB over all ranks) are larger than L2" % (arena_b / 1e9),
                   "partition": ("algebraic modified ND, METIS vertex separators (host, untimed: %.2f s)" if args.algebraic
                                 else "geometric modified ND on linspace_nd coordinates (host, untimed: %.2f s)") % tpart,
                   "symbolic": "block structure of all levels analysed once per pattern in the first assemble() "
                               "(untimed: %.2f s) and reused by every later assemble()/factorize()" % t.analyze_seconds()},
        "factorize_time_s": tdev / args.steps, "fp64_tflops": flops / (tdev / args.steps) / 1e12,
        "gflop_per_factorization": flops / 1e9, "fp64_dgemm_peak_tflops": fp64_peak_1gpu,
        "e2e": {"value": e2e_val, "unit": unit, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "seconds_per_step": te2e / args.steps, "assemble_s": float(np.mean(tassm)),
                "solve_s": float(np.mean(tsolve))},
        "e2e_cold": {"seconds": e2e_cold, "value": N / e2e_cold / 1e6, "unit": unit,
                     "what": "first call on a new matrix: partition + symbolic analysis + assemble + factorize + one solve"},
        "correctness": check,
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        "residual_one_solve": res, ("gmres_iterations" if gen else "cg_iterations"): cg_it,
        ("gmres_seconds" if gen else "cg_seconds"): t_cg, "nnz_fact": int(t.nnz()),
        "arena_gb": arena_b / 1e9,
        "per_level": {k: [float(v) for v in lg[k]] for k in ("t_elim", "t_scale", "t_spars", "t_merge", "t_host",
                                                               "launches", "wavefronts", "dofs_left_spars")},
    }
    if stdout_fd is not None:
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
    print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
