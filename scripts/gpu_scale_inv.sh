#!/bin/bash
# One GPU call for the scaling kernels: parity suites with the defaults, then an A/B of the kernel switches on C4.
mkdir -p gpurun_out
T=${1:-s1}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q > gpurun_out/pytest_scale_$T.log 2>&1
echo "parity + golden rc=$?"
grep "rank report" gpurun_out/pytest_scale_$T.log | sort | uniq | head
tail -4 gpurun_out/pytest_scale_$T.log
shift
bash scripts/gpu_ab.sh scale_$T "$@"
N=$#
python - <<PY
import json
for i in range($N):
    try:
        d=json.loads(open("gpurun_out/ab_scale_${T}_%d.json" % i).read().strip().splitlines()[-1])
        print(i, "t_scale", [round(x*1e3,1) for x in d["per_level"]["t_scale"]])
        print(i, "t_elim ", [round(x*1e3,1) for x in d["per_level"]["t_elim"]])
        print(i, "families", {k:round(v*1e3,1) for k,v in d["roofline"]["family_kernel_seconds"].items()}, "cold", d["e2e_cold"]["seconds"], d["config"]["partition"][-20:], d["config"]["symbolic"][55:90])
    except Exception as e:
        print(i, "failed", e)
PY
