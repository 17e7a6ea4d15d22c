#!/bin/bash
OUT=gpurun_out/${1:-pgo}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_pgo_gpu.py -m gpu -q -x > $OUT/pytest_pgo.log 2>&1; tail -3 $OUT/pytest_pgo.log
timeout 300 python tools/chol_profile.py 100 50 10 > $OUT/chol_solve.log 2>&1; grep direct $OUT/chol_solve.log
timeout 300 python tools/chol_profile.py 250 200 3 > $OUT/chol_solve50k.log 2>&1; grep direct $OUT/chol_solve50k.log
timeout 300 ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none --cache-control none --csv --log-file $OUT/chol_launches.csv python tools/chol_profile.py 100 50 1 > $OUT/ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none --cache-control none --csv --log-file $OUT/chol_launches50k.csv python tools/chol_profile.py 250 200 1 > $OUT/ncu50k.log 2>&1
python tools/chol_launch_list.py $OUT/chol_launches50k.csv > $OUT/chol_launch_list50k.txt; tail -1 $OUT/chol_launch_list50k.txt
python tools/chol_launch_list.py $OUT/chol_launches.csv | tee $OUT/chol_launch_list.txt
