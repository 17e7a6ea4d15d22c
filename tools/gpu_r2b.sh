#!/bin/bash
# round-2: GPU tests + ncu captures of the evaluation kernels (exact after the expf change, tolerance mode DIRECT7 and pclpca/DIRECT1)
OUT=gpurun_out/r2b; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ndt_eval_kernel -s 10 -c 1 -o $OUT/ndt_eval \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_exact.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ndt_eval_fast_kernel -s 10 -c 1 -o $OUT/ndt_eval_fast \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --accumulation fast > $OUT/ncu_fast.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ndt_eval_fast_kernel -s 10 -c 1 -o $OUT/ndt_eval_fast_pca \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --accumulation fast --variant pca > $OUT/ncu_fast_pca.log 2>&1
tail -8 $OUT/pytest_gpu.log; ls -la $OUT
