#!/bin/bash
OUT=gpurun_out/${1:-lat}; mkdir -p $OUT
timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_active.max,sm__cycles_elapsed.max,launch__grid_size,smsp__inst_executed.sum --clock-control none --cache-control none --csv --log-file $OUT/lat.csv python tools/ncu_target.py > $OUT/lat.log 2>&1
python - <<'PY'
import csv,sys
rows=[r for r in csv.reader(open('gpurun_out/lat2/lat.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID')
cur={}
for r in rows[1:]:
    if 'ndt_eval' not in r[ki]: continue
    cur.setdefault((r[ii], r[ki][:40]), {})[r[mi]]=r[vi]
for k,v in list(cur.items())[:14]:
    print(k, v)
PY
