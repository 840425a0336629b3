#!/bin/bash
# One GPU call: kernel-level RRQR tests, (optionally) the GPU suite, bench records, phase statistics.
mkdir -p gpurun_out
T=${1:-r2a}
SUITE=${2:-1}
timeout 600 python -m pytest tests/test_gpu_rrqr.py -x -q > gpurun_out/pytest_rrqr_$T.log 2>&1
echo "rrqr tests rc=$?"
tail -5 gpurun_out/pytest_rrqr_$T.log
if [ "$SUITE" = "1" ]; then
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_rrqr.py > gpurun_out/pytest_gpu_$T.log 2>&1
echo "gpu suite rc=$?"
tail -8 gpurun_out/pytest_gpu_$T.log
fi
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c4_$T.json 2> gpurun_out/bench_c4_$T.err
echo "bench rc=$?"
tail -3 gpurun_out/bench_c4_$T.err
python - <<PY
import json
for f in ("gpurun_out/bench_c4_$T.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["roofline"]["phase_seconds"], d.get("cg_iterations"), d["residual_one_solve"])
        print(" t_spars per level", [round(x*1e3,1) for x in d["per_level"]["t_spars"]])
    except Exception as e:
        print(f, "failed", e)
PY
SPAND_B200_BUILD=_build_timing timeout 300 python scripts/rrqr_phases.py c4 > gpurun_out/phases_$T.json 2>gpurun_out/phases_$T.err
python - <<PY
import json
d=json.load(open("gpurun_out/phases_$T.json"))
print(json.dumps({k:v for k,v in d.items() if k in ("factorize_ms","hot-set kernel")}, indent=1))
PY
