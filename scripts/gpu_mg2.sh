#!/bin/bash
# Two-GPU call: sharded-vs-single parity test, C4 on 1 and 2 GPUs, C5s on 2.
mkdir -p gpurun_out
T=${1:-r2b}
timeout 900 python -m pytest tests/test_mg_gpu.py -q > gpurun_out/pytest_mg2_$T.log 2>&1
echo "mg test rc=$?"; tail -3 gpurun_out/pytest_mg2_$T.log
bash scripts/gpu_scale.sh $T c4 1 2
bash scripts/gpu_scale.sh $T c5s 2
