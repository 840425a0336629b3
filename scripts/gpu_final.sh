#!/bin/bash
# Round-end records on ONE GPU: the whole GPU test suite, the bench lines of the configurations, the ncu launch list
# with DRAM traffic of one C4 factorization. usage: bash scripts/gpu_final.sh <tag>
mkdir -p gpurun_out
T=${1:-r3}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$T.log 2>&1
echo "gpu suite rc=$?"
grep "rank report" gpurun_out/pytest_gpu_$T.log | sort | uniq | head -8
tail -6 gpurun_out/pytest_gpu_$T.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c4_$T.json 2> gpurun_out/bench_c4_$T.err
echo "bench c4 rc=$?"; tail -2 gpurun_out/bench_c4_$T.err
for C in c3 c5s; do
  timeout 600 python bench.py --steps 3 --warmup 3 --config $C --no-cpu-baseline > gpurun_out/bench_${C}_$T.json 2> gpurun_out/bench_${C}_$T.err
  echo "bench $C rc=$?"
done
timeout 600 python bench.py --steps 2 --warmup 2 --config c5 --no-cpu-baseline > gpurun_out/bench_c5_$T.json 2> gpurun_out/bench_c5_$T.err
echo "bench c5 rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --config c2 --algebraic > gpurun_out/bench_c2_algebraic_$T.json 2> gpurun_out/bench_c2_algebraic_$T.err
echo "bench c2 algebraic rc=$?"
python - <<PY
import json
for f in ("bench_c4_$T", "bench_c3_$T", "bench_c5s_$T", "bench_c5_$T", "bench_c2_algebraic_$T"):
    try:
        d=json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],2), "ms", round(d["value"],2), "Mdof/s e2e", round(d["e2e"]["value"],2), "cold", round(d["e2e_cold"]["seconds"],2),
              d["correctness"], {k:round(v*1e3,1) for k,v in d["roofline"]["phase_seconds"].items()},
              {k:round(v*1e3,1) for k,v in d["roofline"]["family_kernel_seconds"].items()}, "frac", round(d["roofline"]["frac"],4),
              d["config"]["partition"][-22:], d["config"]["symbolic"][55:90])
    except Exception as e:
        print(f, "failed", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_$T.csv python scripts/ncu_launchlist.py c4 > gpurun_out/ncu_$T.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_$T.log
python scripts/ncu_traffic.py gpurun_out/launches_$T.csv gpurun_out/${T}_ncu_traffic "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none python scripts/ncu_launchlist.py c4" | head -45
