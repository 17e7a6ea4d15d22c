#!/bin/bash
# Second GPU visit of round 2: everything of gpu_round2.sh plus the direct solver's launch lists, LM timings and the microbenchmarks behind DESIGN.md §5.
TAG=${1:-r02_final2}
bash tools/gpu_round2.sh $TAG
OUT=gpurun_out/$TAG
timeout 300 python tools/chol_profile.py 100 50 20 > $OUT/chol_solve.log 2>&1; grep direct $OUT/chol_solve.log
timeout 300 python tools/chol_profile.py 250 200 3 > $OUT/chol_solve50k.log 2>&1; grep direct $OUT/chol_solve50k.log
timeout 300 ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none --cache-control none --csv --log-file $OUT/chol_launches.csv python tools/chol_profile.py 100 50 1 > $OUT/ncu_chol.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none --cache-control none --csv --log-file $OUT/chol_launches50k.csv python tools/chol_profile.py 250 200 1 > $OUT/ncu_chol50k.log 2>&1
python tools/chol_launch_list.py $OUT/chol_launches50k.csv > $OUT/chol_launch_list50k.txt; tail -1 $OUT/chol_launch_list50k.txt
python tools/chol_launch_list.py $OUT/chol_launches.csv > $OUT/chol_launch_list.txt; tail -1 $OUT/chol_launch_list.txt
LVS_DEBUG_TIMING=1 python tools/chol_profile.py 100 50 1 2>&1 | grep "chol dbg" | tail -2 > $OUT/chol_phase_clocks.log; cat $OUT/chol_phase_clocks.log
timeout 600 python tools/pgo_perf.py --no-cpu --big > $OUT/pgo_perf.log 2>&1; grep "GPU" $OUT/pgo_perf.log
./tools/ubench/chain_lat > $OUT/ubench_chain_lat.log 2>&1; ./tools/ubench/micro_chol > $OUT/ubench_micro_chol.log 2>&1; ./tools/ubench/dmma > $OUT/ubench_dmma.log 2>&1
cat $OUT/ubench_micro_chol.log
