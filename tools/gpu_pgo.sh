#!/bin/bash
OUT=gpurun_out/${1:-pgo}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_pgo_gpu.py -m gpu -q -x > $OUT/pytest_pgo.log 2>&1; tail -5 $OUT/pytest_pgo.log
timeout 300 python tools/chol_profile.py 100 50 10 > $OUT/chol_solve.log 2>&1; cat $OUT/chol_solve.log
timeout 400 python tools/pgo_perf.py --no-cpu --big > $OUT/pgo_perf.log 2>&1; cat $OUT/pgo_perf.log
