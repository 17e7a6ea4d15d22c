#!/bin/bash
# N-GPU visit: H2D bandwidth table at 1..N ranks, bench at every power of two up to N (top level + modes + beam128 point-sharded)
N=${1:-2}; OUT=gpurun_out/multi_n$N; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1; lscpu | grep -i -E "numa|socket|model name|^cpu\(s\)" >> $OUT/topo.txt
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  for b in "" "--bind"; do
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 tools/h2d_bw.py $b 2>/dev/null | grep -E "ranks|rank " >> $OUT/h2d.txt
  done
done
cat $OUT/h2d.txt
for n in 2 4 8; do
  [ $n -le $N ] || continue
  [ $n -ge ${2:-2} ] || continue
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 20 --warmup 3 > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err
  tail -2 $OUT/bench_n$n.err
  python - $OUT/bench_n$n.json <<'PY'
import json,sys
b=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
print("N=%d value %.0f e2e %.0f (ms %.2f / %.2f)" % (b['n_gpus'], b['value'], b['e2e']['value'], b['ms_per_step'], b['e2e']['ms_per_step']))
print("tol value %.0f e2e %.0f" % (b['modes']['tolerance']['value'], b['modes']['tolerance']['e2e']['value']))
print("beam128:", {k:v for k,v in b['configs']['beam128'].items() if k in ('n_gpus','value','ms_per_call','device_ms_per_call','ranks_bit_identical','iterations')})
PY
done
