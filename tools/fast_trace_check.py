"""Diagnostic: trajectories of the exact and the tolerance mode on the config-1 pair, and the accuracy of the tolerance-mode sums along
ITS OWN trajectory (against the CPU oracle at the same states)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import lv_slam_b200 as L
import oracle_ndt as O
from lv_slam_b200 import synth
tgt, src, guess, truth = synth.config1_pair()
tr = {}
objs = {}
for acc in (0, 1):
    n = L.NormalDistributionsTransform(variant=0)
    n.setTransformationEpsilon(0.01); n.setMaximumIterations(64); n.setNeighborhoodSearchMethod(L.LVS_DIRECT7)
    n.setAccumulation(acc)
    n.setInputTarget(tgt); n.setInputSource(src)
    n.align(guess); r = n.result()
    tr[acc] = r["trace"]; objs[acc] = n
o = O.OracleNDT(variant=0, trans_eps=0.01, max_iter=64, search=O.DIRECT7, num_threads=16)
o.set_target(tgt); o.set_source(src)
print("iter | exact: step score p_after[:3] | fast: step score p_after[:3] | |dp| | fast sums vs oracle at the fast state: score g H (rel)")
for k in range(max(len(tr[0]), len(tr[1]))):
    a = tr[0][k] if k < len(tr[0]) else None
    b = tr[1][k] if k < len(tr[1]) else None
    s = "%2d | " % k
    s += ("%.4f %.3f %s" % (a[12], a[13], np.round(a[14:17], 4)) if a is not None else "-") + " | "
    if b is not None:
        s += "%.4f %.3f %s" % (b[12], b[13], np.round(b[14:17], 4))
        if a is not None: s += " | %.2e" % np.abs(a[14:20] - b[14:20]).max()
        if k < 40:
            p = b[14:20]
            gs, gg, gH = objs[1].eval_derivatives(p, None, True)
            os_, og, oH = o.eval_derivatives(p, None, True)
            dn_f, dn_o = np.linalg.solve(gH, -gg), np.linalg.solve(oH, -og)
            s += " | %.1e %.1e %.1e newton |d| %.4f diff %.1e cond %.1e" % (abs(gs - os_) / abs(os_), np.abs(gg - og).max() / np.abs(og).max(), np.abs(gH - oH).max() / np.abs(oH).max(),
                                                             np.linalg.norm(dn_o), np.abs(dn_f - dn_o).max(), np.linalg.cond(oH))
    print(s)
