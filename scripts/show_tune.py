"""Pretty-print a log of scripts/rrqr_tune.py: python scripts/show_tune.py gpurun_out/tune.log"""
import json, sys
for l in open(sys.argv[1]):
    try:
        d = json.loads(l)
    except Exception:
        print(l.rstrip()[:200]); continue
    if "error" in d:
        print(d["variant"], "ERROR", d["error"]); continue
    print(f"{d['variant']:36s} fact {d['factorize_ms']:7.1f} spars {d['sparsify_ms']:7.1f} ranks_differ {d['ranks_differ']} res {d['residual']:.3e}")
    for k in ("spars_per_level_ms", "scale_per_level_ms", "elim_per_level_ms"):
        if k in d:
            print("    ", k[:5], d[k][:14])
