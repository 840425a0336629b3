"""Sweep of the RRQR launch heuristics (environment hooks of Tree::phase_sparsify) on one configuration.
usage: python scripts/rrqr_sweep.py <config> TMIN:L2MB[:CTAS[:SMEM1KB]] ..."""
import os, sys, json, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg = sys.argv[1]
for spec in sys.argv[2:]:
    f = spec.split(":")
    env = dict(os.environ, SPAND_RRQR_TMIN=f[0], SPAND_RRQR_L2MB=f[1])
    if len(f) > 2:
        env["SPAND_RRQR_CTAS"] = f[2]
    if len(f) > 3:
        env["SPAND_RRQR_SMEM1KB"] = f[3]
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "family_times.py"), cfg], env=env,
                         capture_output=True, text=True).stdout
    try:
        j = json.loads(out.strip().splitlines()[-1])
        print(spec, "factorize %.3f s  rrqr %.1f ms  spars per level:" % (j["factorize_s"], j["families"]["rrqr"][0]),
              [round(x * 1e3, 1) for x in j["dev"]["t_spars"]], flush=True)
    except Exception as e:
        print(spec, "failed", e, out[-500:], flush=True)
