#!/bin/bash
# One short GPU visit after the Leaf-identity parity fix: every NDT-side GPU test, smoke(), a short top-level bench.
out=gpurun_out/r02_final8; mkdir -p $out
timeout 100 python -m pytest tests/test_ndt_gpu.py tests/test_ndt_fast_gpu.py tests/test_ndt_ground_gpu.py tests/test_pipeline_gpu.py tests/test_shim_exec_gpu.py tests/test_zz_golden_gpu.py -m gpu -q 2>&1 | tail -25 | tee $out/pytest_gpu_ndt.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $out/smoke.log
timeout 60 python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 3 > $out/bench_top.json 2> $out/bench_top.err
python - <<PY
import json
b = json.load(open("$out/bench_top.json"))
print(b["value"], b["e2e"]["value"], b["ms_per_step"], b["roofline"]["frac"], b["config"])
PY
