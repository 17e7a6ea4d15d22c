#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list and a full capture of the hot kernel.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt; lscpu | head -20 >> $OUT/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?" >> $OUT/bench.err
timeout 600 python bench.py --variant pca > $OUT/bench_pca.json 2> $OUT/bench_pca.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
timeout 300 python tools/gpu_first.py > $OUT/gpu_first.log 2>&1
timeout 400 python tools/pgo_perf.py --no-cpu --big > $OUT/pgo_perf.log 2>&1     # GPU legs of configs 4 and 5 (the CPU legs: tools/pgo_perf.py, ~25 s)
timeout 300 python tools/aux_perf.py > $OUT/aux_perf.log 2>&1
timeout 300 python tools/chol_profile.py 100 50 5 > $OUT/chol_solve.log 2>&1
# launch list of the bench command (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline > $OUT/launches_bench.log 2>&1
# per-launch lists of the two secondary paths: one direct pose-graph solve, three getFitnessScore calls
timeout 200 ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none --csv --log-file $OUT/chol_launches.csv \
    python tools/chol_profile.py 100 50 1 > $OUT/chol_ncu.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/fitness_launches.csv \
    python tools/fitness_profile.py > $OUT/fitness_ncu.log 2>&1
# full capture of the hot kernel inside the batched bench step
# (the first align of an object queues 6 evaluation launches, later ones as many as the previous align needed + 1 = 4: launches 0-5
#  warm-up resident, 6-9 warm-up host path, 10-13 the timed RESIDENT step - 10, 11, 12 are the working 64-pair launches bench.py's
#  roofline is about)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ndt_eval_kernel -s 10 -c 3 -o $OUT/ndt_eval \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ls -la $OUT
tail -3 $OUT/pytest_gpu.log; cat $OUT/smoke.log | tail -3; cat $OUT/bench.json; tail -2 $OUT/bench.err
