#!/bin/bash
# Scaling records on one node: bash scripts/gpu_scale.sh <tag> <config> <N...>
T=$1; C=$2; shift; shift
for N in "$@"; do
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --steps 3 --warmup 2 --config $C --no-cpu-baseline > gpurun_out/scale_${T}_${C}_$N.json 2> gpurun_out/scale_${T}_${C}_$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 3 --warmup 2 --config $C > gpurun_out/scale_${T}_${C}_$N.json 2> gpurun_out/scale_${T}_${C}_$N.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_${T}_${C}_$N.json").read().strip().splitlines()[-1])
    print("$C N=$N", round(d["ms_per_step"],1), "ms", round(d["value"],2), "Mdof/s e2e", round(d["e2e"]["value"],2), d["correctness"]["iterations"], d["correctness"]["ok"], {k:round(v*1e3,1) for k,v in d["roofline"]["phase_seconds"].items()}, "cold", round(d["e2e_cold"]["seconds"],2))
except Exception as e:
    print("$C N=$N failed", e); print(open("gpurun_out/scale_${T}_${C}_$N.err").read()[-800:])
PY
done
