#!/bin/bash
# A/B of environment variants on the C4 bench: scripts/gpu_ab.sh tag "VAR=1 VAR2=3" "..." 
T=$1; shift
i=0
for v in "$@"; do
  env $v timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-cg > gpurun_out/ab_${T}_$i.json 2> gpurun_out/ab_${T}_$i.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_${T}_$i.json").read().strip().splitlines()[-1])
    print("[$v]", round(d["ms_per_step"],1), {k:round(x*1e3,1) for k,x in d["roofline"]["phase_seconds"].items()}, d["residual_one_solve"])
    print("   t_spars", [round(x*1e3,1) for x in d["per_level"]["t_spars"]])
except Exception as e:
    print("[$v] failed", e)
PY
  i=$((i+1))
done
