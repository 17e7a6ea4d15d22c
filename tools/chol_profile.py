"""One direct solve of the pose-graph normal equations (sparse block Cholesky), timed; run under ncu for the per-launch list:
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/chol_launches.csv python tools/chol_profile.py 100 50 1"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import lv_slam_b200 as L
from lv_slam_b200.synth import posegraph as G
npl, laps, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
g = G.sphere(npl, laps, seed=7)
pg = L.PoseGraph(0)
pg.set_graph(g["poses7"], g["ij"], g["meas7"], g["info21"], g["huber"])
print(pg.chol_info())
lin = pg.linearize()
lam = 1e-5 * np.max(np.abs(np.einsum("nii->ni", lin["Hd"])))
x, _ = pg.solve(lam, 0.0)
t = time.time()
for _ in range(reps): x, _ = pg.solve(lam, 0.0)
print("direct solve %.3f ms (wall, incl. result copy)" % ((time.time() - t) / max(reps, 1) * 1e3))
pg2 = L.PoseGraph(0); pg2.set_graph(g["poses7"], g["ij"], g["meas7"], g["info21"], g["huber"]); pg2.linearize()
xp, it = pg2.solve(lam, 1e-24)   # tolerance > 0 is honoured only by the iterative path when an override is set
print("max |x|", float(np.abs(x).max()))
