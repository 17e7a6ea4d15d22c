#!/bin/bash
# 2-GPU check of the driver's launch line at HEAD (frame-sharded top level, point-sharded beam128, replay), nothing else.
OUT=gpurun_out/r02_n2_final; mkdir -p $OUT
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > $OUT/bench_n2.json 2> $OUT/bench_n2.err
tail -2 $OUT/bench_n2.err
python - $OUT/bench_n2.json <<'PY'
import json,sys
b=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
print("N=%d value %.0f e2e %.0f (ms %.2f / %.2f)" % (b['n_gpus'], b['value'], b['e2e']['value'], b['ms_per_step'], b['e2e']['ms_per_step']))
print("beam128:", {k:v for k,v in b['configs']['beam128'].items() if k in ('n_gpus','value','ms_per_call','device_ms_per_call','ranks_bit_identical','iterations')})
print("replay:", {k:v for k,v in b['configs']['replay'].items() if k != 'workload'})
PY
