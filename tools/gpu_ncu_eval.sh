#!/bin/bash
# ncu --set full capture of the hot kernel inside one batched bench step.  Usage: bash tools/gpu_ncu_eval.sh tag
TAG=${1:-ncu}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ndt_eval_kernel -s 20 -c 2 -o $OUT/ndt_eval \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ls -la $OUT
