"""First GPU contact: quick parity numbers + timings of the single-pair align path."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import oracle_ndt as O
import lv_slam_b200 as L
from lv_slam_b200 import synth

tgt, src, guess, truth = synth.config1_pair()
for variant, search in ((0, 2), (1, 3)):
    n = L.NormalDistributionsTransform(variant=variant)
    n.setTransformationEpsilon(0.01); n.setMaximumIterations(64); n.setNeighborhoodSearchMethod(search)
    o = O.OracleNDT(variant=variant, trans_eps=0.01, max_iter=64, search=search, num_threads=os.cpu_count())
    t = time.time(); n.setInputTarget(tgt); t1 = time.time() - t
    t = time.time(); n.setInputTarget(tgt); t2 = time.time() - t
    t = time.time(); o.set_target(tgt); t3 = time.time() - t
    print("set_target gpu first %.2f ms, second %.2f ms, oracle %.2f ms" % (t1 * 1e3, t2 * 1e3, t3 * 1e3))
    n.setInputSource(src); o.set_source(src)
    p = O.se3_log_from_matrix4f(guess)
    gs, gg, gH = n.eval_derivatives(p, guess, True)
    os_, og, oH = o.eval_derivatives(p, guess, True)
    print("score", gs, os_, "g rel", np.max(np.abs(gg - og)) / np.max(np.abs(og)), "H rel", np.max(np.abs(gH - oH)) / np.max(np.abs(oH)))
    for rep in range(3):
        t = time.time(); n.align(guess); ta = time.time() - t
        r = n.result()
    import ctypes
    from lv_slam_b200 import _capi as C
    b = n.batch_handle()
    ms, dms = ctypes.c_double(0), ctypes.c_double(0); nl, dl = ctypes.c_int(0), ctypes.c_int(0)
    C.lib().lvs_ndt_batch_set_profiling(b, 1)
    n.align(guess)
    C.lib().lvs_ndt_batch_last_stats(b, ctypes.byref(ms), ctypes.byref(nl), ctypes.byref(dms), ctypes.byref(dl))
    t = time.time(); ro = o.align(guess); to = time.time() - t
    print("variant", variant, "search", search, "gpu align wall %.3f ms (device %.3f ms, %d launches, eval kernels %.3f ms over %d), iters %d n_eval %d | oracle %.1f ms iters %d (threads %d)" % (
        ta * 1e3, ms.value, nl.value, dms.value, dl.value, r["iterations"], r["n_eval"], to * 1e3, ro["iterations"], os.cpu_count()))
    print("final diff", np.max(np.abs(r["final"] - ro["final"])))
