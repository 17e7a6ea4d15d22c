#!/bin/bash
# round-2 first visit: parity tests (all), bench exact / fast / pca
OUT=gpurun_out/r2a; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
timeout 600 python -m pytest tests/test_ndt_fast_gpu.py -m gpu -q -s > $OUT/pytest_fast.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_fast.log
timeout 300 python bench.py --no-cpu-baseline > $OUT/bench_exact.json 2> $OUT/bench_exact.err
timeout 300 python bench.py --no-cpu-baseline --accumulation fast > $OUT/bench_fast.json 2> $OUT/bench_fast.err
timeout 300 python bench.py --no-cpu-baseline --variant pca > $OUT/bench_pca_exact.json 2> $OUT/bench_pca_exact.err
timeout 300 python bench.py --no-cpu-baseline --variant pca --accumulation fast > $OUT/bench_pca_fast.json 2> $OUT/bench_pca_fast.err
tail -5 $OUT/pytest_gpu.log; tail -15 $OUT/pytest_fast.log; cat $OUT/bench_*.json; tail -2 $OUT/*.err
