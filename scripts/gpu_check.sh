#!/bin/bash
# GPU suite + C4 / C5s bench lines (no CPU baseline). usage: bash scripts/gpu_check.sh <tag>
mkdir -p gpurun_out
T=${1:-chk}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$T.log 2>&1
echo "gpu suite rc=$?"
grep "rank report" gpurun_out/pytest_gpu_$T.log | sort | uniq | head -8
tail -4 gpurun_out/pytest_gpu_$T.log
for C in c4 c5s; do
  timeout 600 python bench.py --steps 3 --warmup 3 --config $C --no-cpu-baseline > gpurun_out/bench_${C}_$T.json 2> gpurun_out/bench_${C}_$T.err
  echo "bench $C rc=$?"
done
python - <<PY
import json
for f in ("bench_c4_$T", "bench_c5s_$T"):
    try:
        d=json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],2), "ms e2e s/step", round(d["e2e"]["seconds_per_step"],4), "assemble", round(d["e2e"]["assemble_s"],4), "solve", round(d["e2e"]["solve_s"],4),
              "krylov s", d.get("cg_seconds", d.get("gmres_seconds")), d["correctness"]["iterations"], d["correctness"]["ok"], {k:round(v*1e3,1) for k,v in d["roofline"]["phase_seconds"].items()},
              {k:round(v*1e3,1) for k,v in d["roofline"]["family_kernel_seconds"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
