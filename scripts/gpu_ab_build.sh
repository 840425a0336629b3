#!/bin/bash
# A/B of two builds on the C4 and C5s bench lines, then the GPU suite with the default build.
# usage: bash scripts/gpu_ab_build.sh <tag> <other build dir, e.g. _build_old>
mkdir -p gpurun_out
T=$1; OTHER=$2
for B in $OTHER _build; do
  for C in c4 c5s; do
    SPAND_B200_BUILD=$B timeout 600 python bench.py --steps 3 --warmup 3 --config $C --no-cpu-baseline --no-cg > gpurun_out/ab_${T}_${B}_$C.json 2> gpurun_out/ab_${T}_${B}_$C.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_${T}_${B}_$C.json").read().strip().splitlines()[-1])
    print("$B $C", round(d["ms_per_step"],2), "ms", {k:round(v*1e3,1) for k,v in d["roofline"]["phase_seconds"].items()}, {k:round(v*1e3,1) for k,v in d["roofline"]["family_kernel_seconds"].items()}, d["residual_one_solve"])
except Exception as e:
    print("$B $C failed", e)
PY
  done
done
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$T.log 2>&1
echo "gpu suite (default build) rc=$?"; tail -3 gpurun_out/pytest_gpu_$T.log
