"""In-process sweep of the RRQR launch heuristics (environment hooks read by Tree::phase_sparsify at every
factorize()): the tree is partitioned / analysed once, every variant costs two factorizations.
usage: python scripts/rrqr_tune.py <config> [--dump file] VAR=val,VAR=val ...      ("base" = no overrides)
Prints factorize / sparsify device time per variant and whether the ranks equal those of the first variant."""
import gzip, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import spand_public_b200 as S

args = sys.argv[1:]
cfg = bench.parse_config(args.pop(0))
dump = None
if args and args[0] == "--dump":
    args.pop(0)
    dump = args.pop(0)
n, d, L, tol, desc = cfg
A = S.neglapl(n, d); X = S.linspace_nd(n, d)
t = S.Tree(L); t.set_tol(tol); t.set_use_geo(True); t.set_Xcoo(X); t.partition(S.symmetric_graph(A))
if dump:
    os.environ["SPAND_DUMP_QR"] = dump
    if os.path.exists(dump):
        os.remove(dump)
t.assemble(A); t.factorize()
if dump:
    os.environ.pop("SPAND_DUMP_QR")
    with open(dump, "rb") as f, gzip.open(dump + ".gz", "wb") as g:
        g.write(f.read())
    os.remove(dump)
ref_ranks = None
b = S.random(A.shape[0], 2019)
for spec in args or ["base"]:
    keys = []
    if spec != "base":
        for kv in spec.split(","):
            k, v = kv.split("=")
            name = k if k.startswith("SPAND_") else "SPAND_RRQR_" + k
            os.environ[name] = v
            keys.append(name)
    try:
        best = None
        for rep in range(2):
            t.assemble(A); t.factorize()
            lg = t.log()
            cur = (t.factorize_seconds(), float(lg["t_spars"].sum()), [round(float(x) * 1e3, 1) for x in lg["t_spars"]],
                   [round(float(x) * 1e3, 1) for x in lg["t_scale"]], [round(float(x) * 1e3, 1) for x in lg["t_elim"]])
            best = cur if best is None or cur[0] < best[0] else best
        ranks = t.stats()[2].copy()
        if ref_ranks is None:
            ref_ranks = ranks
        x = t.solve(b)
        res = float(np.linalg.norm(A @ x - b) / np.linalg.norm(b))
        print(json.dumps({"variant": spec, "factorize_ms": round(best[0] * 1e3, 1), "sparsify_ms": round(best[1] * 1e3, 1),
                          "ranks_differ": int((ranks != ref_ranks).sum()), "residual": res, "spars_per_level_ms": best[2],
                          "scale_per_level_ms": best[3], "elim_per_level_ms": best[4]}),
              flush=True)
    except Exception as e:  # a variant that does not fit (shared memory, scratch) must not end the sweep
        print(json.dumps({"variant": spec, "error": str(e)[:300]}), flush=True)
    for k in keys:
        os.environ.pop(k, None)
