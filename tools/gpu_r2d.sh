#!/bin/bash
OUT=gpurun_out/${1:-r2d}; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -6 $OUT/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?" >> $OUT/bench.err
tail -3 $OUT/bench.err
python - $OUT/bench.json <<'PY'
import json,sys
b=json.load(open(sys.argv[1]))
print("top value %.0f e2e %.0f launch %.3f ms" % (b['value'], b['e2e']['value'], b['roofline']['avg_launch_ms']))
t=b['modes']['tolerance']; print("tol value %.0f e2e %.0f launch %.3f ms" % (t['value'], t['e2e']['value'], t['roofline']['avg_launch_ms']))
for k,v in b['configs']['pca_direct1'].items():
    if 'value' in v: print("pca", k, "value %.0f e2e %.0f launch %.3f ms" % (v['value'], v['e2e']['value'], v['roofline']['avg_launch_ms']))
print("beam128", b['configs']['beam128']['value'], b['configs']['beam128']['ms_per_call'])
pl=b['configs']['pair_latency']; print("pair exact", pl['exact']); print("pair tol", pl['tolerance'])
p=b['configs']['pgo']; print("pgo lm_direct %.1f ms, pcg %.1f ms, gn %.1f ms" % (p['lm_direct']['optimize_ms'], p['lm_pcg']['optimize_ms'], p['gn_direct']['optimize_ms']))
PY
