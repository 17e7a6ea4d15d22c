#!/bin/bash
# round-2: the full bench line (both arms) + ncu captures of the four evaluation kernels of the bench's own 64-pair launches
OUT=gpurun_out/r2c; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt; lscpu | head -20 >> $OUT/nproc.txt
( time timeout 900 python bench.py ) > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?" >> $OUT/bench.err
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 1 ) > $OUT/bench_ref.json 2> $OUT/bench_ref.err
# launches 10.. of `bench.py --steps 1 --warmup 1` are the timed resident step's working 64-pair launches (0-5 first align incl. idle ones, 6-9 warm-up e2e)
for spec in "ndt_eval:ndt_eval_kernel:10:" "ndt_eval_fast:ndt_eval_fast_kernel:10:--accumulation fast" "ndt_eval_pca:ndt_eval_kernel:11:--variant pca" "ndt_eval_fast_pca:ndt_eval_fast_kernel:11:--variant pca --accumulation fast"; do
  IFS=: read name kern skip flags <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^${kern}" -s $skip -c 2 -o $OUT/$name \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras $flags > $OUT/ncu_$name.log 2>&1
done
ls -la $OUT; du -sh gpurun_out
