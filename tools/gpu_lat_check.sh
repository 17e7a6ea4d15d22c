#!/bin/bash
timeout 1200 python -m pytest tests/test_ndt_gpu.py tests/test_ndt_fast_gpu.py tests/test_zz_golden_gpu.py -m gpu -q -x 2>&1 | tail -6
LVS_DEBUG_TIMING=1 python tools/tail_timing.py 2>&1 | tail -24
python - <<'PY'
import time, numpy as np, lv_slam_b200 as L
from lv_slam_b200 import synth
tgt, src, guess, truth = synth.config1_pair()
for acc in (0, 1):
    n = L.NormalDistributionsTransform(variant=0)
    n.setTransformationEpsilon(0.01); n.setMaximumIterations(64); n.setNeighborhoodSearchMethod(2); n.setAccumulation(acc)
    n.setInputTarget(tgt); n.setInputSource(src)
    for _ in range(5): n.align(guess)
    t0 = time.perf_counter()
    for _ in range(20): n.align(guess)
    dt = (time.perf_counter() - t0) / 20
    r = n.result()
    print("acc %d: align %.3f ms, %d iterations, %.1f us per Newton iteration" % (acc, dt * 1e3, r["iterations"], dt * 1e6 / (r["iterations"] + 1)))
PY
