"""Writes tests/golden/*.json from the CPU oracle on the seeded small inputs.

These are REGRESSION pins of the oracle, not reference outputs: the reference (and PCL / Eigen / Sophus / g2o) cannot be built
in this image, and it ships no golden vectors of its own."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_ndt as O  # noqa: E402
import oracle_pgo as P  # noqa: E402
from lv_slam_b200 import synth  # noqa: E402
from lv_slam_b200.synth import posegraph as G  # noqa: E402

tgt, src, guess, truth = synth.config1_pair(n_beams=16, n_az=600, seed=5)
o = O.OracleNDT(trans_eps=0.01, max_iter=30, search=O.DIRECT7, num_threads=1)
o.set_target(tgt); o.set_source(src)
lv = o.leaves()
s, g, H = o.eval_derivatives(O.se3_log_from_matrix4f(guess), guess)
r = o.align(guess)
json.dump(dict(n_cells=len(lv["keys"]), n_usable=int((lv["nr_points"] >= 6).sum()), key_sum=int(lv["keys"].astype(np.int64).sum()), score=s,
               gradient=g.tolist(), hessian=H.tolist(), iterations=r["iterations"], final=r["final"].astype(float).tolist()),
          open(os.path.join(HERE, "ndt_small_pair.json"), "w"), indent=1)

# pclomp_ground (ndt_ground_impl.hpp) on the same pair, configured like ground_s2k (scan_matching_odom_nodelet.cpp:121-126)
og = O.OracleNDT(variant=O.VAR_GROUND, resolution=10.0, trans_eps=0.01, max_iter=64, search=O.DIRECT1, num_threads=1)
og.set_target(tgt); og.set_source(src)
ang = og.leaf_angles()
sg, gg, Hg = og.eval_derivatives(O.se3_log_from_matrix4f(guess), guess)
rg = og.align(guess)
json.dump(dict(n_cells=len(ang), n_with_normal=int((ang >= 0).sum()), n_horizontal=int(((ang >= 0) & (ang < 10)).sum()),
               horizontal_key_sum=int(og.leaves()["keys"][(ang >= 0) & (ang < 10)].astype(np.int64).sum()), score=sg, gradient=gg.tolist(),
               hessian=Hg.tolist(), iterations=rg["iterations"], final=rg["final"].astype(float).tolist()),
          open(os.path.join(HERE, "ndt_ground_small_pair.json"), "w"), indent=1)

# the stages either side of the path on the same pair
big = float(np.finfo(np.float64).max)
fit = {}
for name, T, mr in (("truth", truth, big), ("guess", guess, big), ("guess_capped", guess, 0.25)):
    sc, cnt = o.fitness_score(T, mr)
    fit[name] = dict(score=sc, correspondences=cnt)
cloud = np.concatenate([tgt[:, :3], np.random.default_rng(3).random((len(tgt), 1), dtype=np.float32)], axis=1).astype(np.float32)
pf, _ = O.prefilter(cloud, 0.5, 100.0, True, 0.1)
json.dump(dict(fitness=fit, prefilter=dict(n_in=len(cloud), n_out=len(pf), column_sums=pf.astype(np.float64).sum(axis=0).tolist(),
                                           first=pf[0].astype(float).tolist(), last=pf[-1].astype(float).tolist())),
          open(os.path.join(HERE, "aux_small_pair.json"), "w"), indent=1)

gr = G.sphere(20, 10, seed=7)
p = P.OraclePGO()
p.set_graph(gr["poses7"], gr["ij"], gr["meas7"], gr["info21"], gr["huber"])
e, c, tot = p.errors()
st = p.optimize(100, P.ALG_LM, P.SOLVER_DENSE)
json.dump(dict(n_vertices=len(gr["poses7"]), n_edges=len(gr["ij"]), robust_chi2_initial=tot, chi2_initial=float(c.sum()), chi2_final=st["chi2_after"],
               first_lambdas=st["trace"][:5, 1].tolist(), first_chi2=st["trace"][:5, 0].tolist()),
          open(os.path.join(HERE, "pgo_sphere_200.json"), "w"), indent=1)
# the same sphere with the global graph's unary constraints (GPS / IMU priors, the floor constraint) on every fifth vertex
sys.path.insert(0, os.path.dirname(HERE))
from test_oracle_pgo import _priors_on, FLOOR  # noqa: E402
ij, meas, info, hub, ty = _priors_on(gr, np.random.default_rng(5), 5)
p = P.OraclePGO()
p.set_graph(gr["poses7"], ij, meas, info, hub, None, ty, FLOOR)
e, c, tot = p.errors()
st = p.optimize(100, P.ALG_LM, P.SOLVER_DENSE)
json.dump(dict(n_vertices=len(gr["poses7"]), n_edges=len(ij), n_unary=int((ty != 0).sum()), kinds=sorted(set(int(t) for t in ty)), floor_plane=FLOOR.tolist(),
               robust_chi2_initial=tot, chi2_initial=float(c.sum()), unary_error_abs_sum=float(np.abs(e[ty != 0]).sum()), chi2_final=st["chi2_after"],
               pose_0=p.poses()[0].tolist(), pose_last=p.poses()[-1].tolist()),
          open(os.path.join(HERE, "pgo_sphere_200_unary.json"), "w"), indent=1)
print("golden files written")
