"""Diagnostic: repeated tolerance-mode aligns of the config-1 pair must be bit-identical run to run."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import lv_slam_b200 as L
from lv_slam_b200 import synth
tgt, src, guess, truth = synth.config1_pair()
for acc in (0, 1):
    n = L.NormalDistributionsTransform(variant=0)
    n.setTransformationEpsilon(0.01); n.setMaximumIterations(64); n.setNeighborhoodSearchMethod(L.LVS_DIRECT7)
    n.setAccumulation(acc)
    n.setInputTarget(tgt); n.setInputSource(src)
    first = None
    for k in range(8):
        if k == 5: n.setInputSource(src)
        n.align(guess); r = n.result()
        p = np.zeros(6); s, g, H = n.eval_derivatives(r["trace"][0, 14:20], None, True)
        if first is None: first = (r["final"].copy(), s, g.copy(), H.copy())
        print("acc", acc, "align", k, "iters", r["iterations"], "n_eval", r["n_eval"], "final t", r["final"][:3, 3], "same as first:", np.array_equal(first[0], r["final"]),
              "tap same:", s == first[1] and np.array_equal(g, first[2]) and np.array_equal(H, first[3]))
