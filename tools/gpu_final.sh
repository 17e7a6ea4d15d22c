#!/bin/bash
# What the driver runs at round end, in one GPU visit: the GPU test suite, smoke(), and both bench arms.  Usage: tools/gpu_final.sh <tag>
tag=${1:-final}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests/ -m gpu -q 2>&1 | tail -15 | tee $out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $out/smoke.log
timeout 600 python bench.py --impl reference > $out/bench_ref.json 2> $out/bench_ref.err
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err
python - <<PY
import json
b = json.load(open("$out/bench.json")); r = json.load(open("$out/bench_ref.json"))
print(b["value"], b["e2e"]["value"], b["roofline"]["frac"], b["gpu_launches"], b["clocks"])
print("ref", r["value"])
print("ground", json.dumps(b["configs"].get("ground_s2k")))
print("pair", b["configs"]["pair_latency"]["exact"]["align_ms"], b["configs"]["pair_latency"]["tolerance"]["align_ms"])
p = b["configs"]["pgo"]; print("pgo", json.dumps({k: p[k] for k in p if k.startswith("lm") or k.startswith("set_graph") or "solve" in k})[:900])
PY
