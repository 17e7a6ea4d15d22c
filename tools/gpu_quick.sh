#!/bin/bash
# Quick GPU check: parity tests + bench.  Usage: bash tools/gpu_quick.sh tag [extra bench args]
TAG=${1:-q}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 600 python bench.py "$@" > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?" >> $OUT/bench.err
cat $OUT/bench.json; tail -3 $OUT/bench.err
timeout 300 python tools/step_breakdown.py > $OUT/breakdown.log 2>&1; cat $OUT/breakdown.log
