#!/bin/bash
# One GPU call: parity suites with the compiled defaults, A/B of the scaling-kernel switches on the C4 bench, then the
# parity suites and the round-end records (scripts/gpu_final.sh) under the fastest variant whose suites are green.
mkdir -p gpurun_out
T=${1:-r2b}
VARS=("SPAND_MID256=0 SPAND_SCALE_INV=0 SPAND_FASTDIV=0" "SPAND_MID256=1 SPAND_SCALE_INV=0 SPAND_FASTDIV=0" "SPAND_MID256=1 SPAND_SCALE_INV=1 SPAND_FASTDIV=0" "SPAND_MID256=1 SPAND_SCALE_INV=1 SPAND_FASTDIV=1" "SPAND_MID256=0 SPAND_SCALE_INV=1 SPAND_FASTDIV=1")
bash scripts/gpu_ab.sh $T "${VARS[@]}"
BEST=$(python - <<PY
import json
best, bi = 1e9, 0
for i in range(${#VARS[@]}):
    try:
        d = json.loads(open("gpurun_out/ab_${T}_%d.json" % i).read().strip().splitlines()[-1])
        print(i, "t_scale", [round(x*1e3,1) for x in d["per_level"]["t_scale"]], "t_elim", [round(x*1e3,1) for x in d["per_level"]["t_elim"]], file=__import__("sys").stderr)
        if d["ms_per_step"] < best:
            best, bi = d["ms_per_step"], i
    except Exception as e:
        print(i, "failed", e, file=__import__("sys").stderr)
print(bi)
PY
)
echo "fastest variant: $BEST = ${VARS[$BEST]}"
export ${VARS[$BEST]}
echo "${VARS[$BEST]}" > gpurun_out/best_variant_$T.txt
bash scripts/gpu_final.sh $T
