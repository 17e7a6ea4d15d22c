import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np
import lv_slam_b200 as L, oracle_pgo as P, oracle_ndt as O
from lv_slam_b200.synth import posegraph as G
from lv_slam_b200 import synth
g = G.sphere(50, 20, seed=11)
pg = L.PoseGraph(0); pg.set_graph(g["poses7"], g["ij"], g["meas7"], g["info21"], g["huber"])
o = P.OraclePGO(); o.set_graph(g["poses7"], g["ij"], g["meas7"], g["info21"], g["huber"])
ol = o.linearize(); pg.linearize()
lam = 1e-5 * np.max(np.abs(np.einsum("nii->ni", ol["Hd"])))
ok, xo, _ = o.solve(lam, P.SOLVER_CSPARSE)
xg, _ = pg.solve(lam, 0.0)
print("direct solve vs CSparse rel", np.max(np.abs(xg - xo)) / np.max(np.abs(xo)))
tgt, src, guess, truth = synth.config1_pair(n_beams=16, n_az=600, seed=5)
n = L.NormalDistributionsTransform(); n.setInputTarget(tgt); n.setInputSource(src)
oo = O.OracleNDT(num_threads=8); oo.set_target(tgt); oo.set_source(src)
for T in (truth, guess):
    a = n.getFitnessScore(1e300, T=T); b = oo.fitness_score(T, 1e300)[0]
    print("fitness rel", abs(a - b) / b)
