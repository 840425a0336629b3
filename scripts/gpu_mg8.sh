#!/bin/bash
# Eight-GPU records: C4 and C5 sharded over the node.
mkdir -p gpurun_out
T=${1:-r2b}
bash scripts/gpu_scale.sh $T c4 8
bash scripts/gpu_scale.sh $T c5 8
