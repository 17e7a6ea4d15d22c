#!/bin/bash
# Round-2 GPU visit (one GPU): parity tests, smoke, bench (both arms), ncu launch list of the bench command, full captures of the four
# evaluation kernels of the bench's own 64-pair launches, solver launch lists.  Usage (under gpurun): bash tools/gpu_round2.sh [tag]
TAG=${1:-r02_final}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt; lscpu | head -20 >> $OUT/nproc.txt
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log
( time timeout 900 python bench.py ) > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?" >> $OUT/bench.err
( time timeout 900 python bench.py --impl reference --steps 10 --warmup 3 ) > $OUT/bench_ref.json 2> $OUT/bench_ref.err
# launch list of the bench command (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $OUT/launches_bench.log 2>&1
# launches 10.. of `bench.py --steps 1 --warmup 1` are the timed resident step's working 64-pair launches
for spec in "ndt_eval:ndt_eval_kernel:10:" "ndt_eval_fast:ndt_eval_fast_kernel:10:--accumulation fast" "ndt_eval_pca:ndt_eval_kernel:11:--variant pca" "ndt_eval_fast_pca:ndt_eval_fast_kernel:11:--variant pca --accumulation fast"; do
  IFS=: read name kern skip flags <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^${kern}" -s $skip -c 2 -o $OUT/$name \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras $flags > $OUT/ncu_$name.log 2>&1
done
timeout 300 python tools/tail_timing.py > $OUT/tail_timing.log 2>&1
timeout 300 python tools/aux_perf.py > $OUT/aux_perf.log 2>&1
ls -la $OUT; du -sh gpurun_out; tail -3 $OUT/pytest_gpu.log; tail -3 $OUT/smoke.log; tail -4 $OUT/bench.err; tail -4 $OUT/bench_ref.err
