"""Diagnostics: where the last CTA of a single-pair evaluation spends its time (LVS_DEBUG_TIMING=1 clock64 stamps).
   LVS_DEBUG_TIMING=1 python tools/tail_timing.py"""
import ctypes, os, sys, time
os.environ["LVS_DEBUG_TIMING"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import lv_slam_b200 as L
from lv_slam_b200 import synth, _capi as C
tgt, src, guess, truth = synth.config1_pair()
lib = C.lib()
lib.lvs_ndt_batch_debug_stamps.restype = ctypes.c_int
lib.lvs_ndt_batch_debug_stamps.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
names = ["entry->body end", "ticket", "state copy + partial loads", "row sum (+exchange)", "deposit", "warp LU", "state advance", "copy back"]
for acc, variant, search, iters in ((0, 0, 2, 5), (1, 0, 2, 5), (0, 1, 3, 5)):
    n = L.NormalDistributionsTransform(variant=variant)
    n.setTransformationEpsilon(0.01); n.setMaximumIterations(iters); n.setNeighborhoodSearchMethod(search); n.setAccumulation(acc)
    n.setInputTarget(tgt); n.setInputSource(src)
    for _ in range(3): n.align(guess)
    t0 = time.perf_counter(); n.align(guess); wall = time.perf_counter() - t0
    st = (ctypes.c_longlong * 8)()
    C.check(lib.lvs_ndt_batch_debug_stamps(n.batch_handle(), st))
    s = list(st)
    print("acc %d variant %d search %d: align %.1f us for %d evaluations" % (acc, variant, search, wall * 1e6, n.result()["n_eval"]))
    if acc == 0:
        for i in range(1, 8):
            print("   %-28s %7d clk  %6.2f us" % (names[i - 1], s[i] - s[i - 1], (s[i] - s[i - 1]) / 1965.0))
    else:
        for i in range(2, 8):
            print("   %-28s %7d clk  %6.2f us" % (names[i - 1], s[i] - s[i - 1], (s[i] - s[i - 1]) / 1965.0))
