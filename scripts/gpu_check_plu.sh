#!/bin/bash
# PLU checks: parity suite + goldens, then C5s with the explicit-inverse scaling of the large blocks on / off.
mkdir -p gpurun_out
T=${1:-plu}
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -q > gpurun_out/pytest_plu_$T.log 2>&1
echo "parity + golden rc=$?"
grep "rank report" gpurun_out/pytest_plu_$T.log | sort | uniq | head -8
tail -4 gpurun_out/pytest_plu_$T.log
for V in 0 1; do
  SPAND_SCALE_INV=$V timeout 600 python bench.py --steps 3 --warmup 3 --config c5s --no-cpu-baseline > gpurun_out/bench_c5s_${T}_inv$V.json 2> gpurun_out/bench_c5s_${T}_inv$V.err
  echo "bench c5s inv=$V rc=$?"
done
python - <<PY
import json
for V in (0, 1):
    try:
        d=json.loads(open("gpurun_out/bench_c5s_${T}_inv%d.json" % V).read().strip().splitlines()[-1])
        print("inv", V, round(d["ms_per_step"],2), "ms", d["correctness"]["iterations"], d["correctness"]["ok"], d["residual_one_solve"], {k:round(v*1e3,1) for k,v in d["roofline"]["phase_seconds"].items()},
              {k:round(v*1e3,1) for k,v in d["roofline"]["family_kernel_seconds"].items()})
        print("   t_scale", [round(x*1e3,1) for x in d["per_level"]["t_scale"]])
    except Exception as e:
        print(V, "failed", e)
PY
