"""CPU-oracle stand-ins with the method names lv_slam_b200.pipeline.replay drives (test infrastructure)."""
import numpy as np

import oracle_ndt as O
import oracle_pgo as P
from lv_slam_b200 import graph_slam as GS
from lv_slam_b200.information_matrix import InformationMatrixCalculator
from lv_slam_b200.synth import posegraph as G

DBL_MAX = float(np.finfo(np.float64).max)


class OracleRegistration:
    def __init__(self, variant, search, threads=8):
        self.o = O.OracleNDT(variant=variant, trans_eps=0.01, max_iter=64, search=search, num_threads=threads)
        self.res = None

    def setInputTarget(self, c): self.o.set_target(c)
    def setInputSource(self, c): self.o.set_source(c)
    def align(self, guess): self.res = self.o.align(guess)
    def getFinalTransformation(self): return self.res["final"]
    def hasConverged(self): return self.res["converged"]
    def getFitnessScore(self, max_range=DBL_MAX): return self.o.fitness_score(self.res["final"], max_range)[0]


class OracleInformation(InformationMatrixCalculator):
    def calc_fitness_score(self, cloud1, cloud2, relpose, max_range=DBL_MAX):
        o = O.OracleNDT(num_threads=8)
        o.set_target(cloud1); o.set_source(cloud2)
        return o.fitness_score(np.asarray(relpose, dtype=np.float64).astype(np.float32), max_range)[0]


class OracleGraphSLAM(GS.GraphSLAM):
    """The host containers of the mirror, optimised by the CPU restatement of g2o's LM + the reference's CSparse."""
    def optimize(self, num_iterations):
        if len(self._edges) < 1:
            return -1
        poses, fixed, ij, meas, info, hub, _types = self._arrays()
        o = P.OraclePGO()
        o.set_graph(poses, ij, meas, info, hub, fixed)
        r = o.optimize(num_iterations, P.ALG_LM, P.SOLVER_CSPARSE if P.have_csparse() else P.SOLVER_DENSE)
        for v, p in zip(self._vertices, o.poses()):
            v._T = G.matrix(p)
        self.last_stats = r
        return r["iterations"]


class OraclePrefilter:
    def __init__(self, near=0.5, far=100.0, leaf=0.1): self.a = (near, far, True, leaf)
    def filter(self, cloud): return O.prefilter(cloud, *self.a)[0]


def out_and_back(n_out=50, n_turn=24, n_back=50, n_beams=32, n_az=900, seed=79):
    """A drive down the synthetic street and back in reverse gear (same heading), so the return leg revisits the outbound
    keyframes: the loop detector has something to find.  Returns (scans, ground-truth poses).
    The sparse 16-beam street barely constrains the motion ALONG the street, so the odometry coasts on its constant-velocity guess
    and a sharp reversal can leave it stuck (on the CPU restatement as much as on the device); the turn is therefore gentle
    (0.1 m/frame^2) and the seed is one on which the chain tracks with margin.  32 beams x 900 azimuths (28 k points): the reference's
    voxel covariances carry + I (n - 1) / n^2 (its Leaf starts cov_ at the identity), which blurs cells with few points so much that a
    16-beam scan no longer tracks this drive - on the reference's own arithmetic, hence on both sides of the comparison."""
    from lv_slam_b200 import synth
    v = [1.2] * n_out + [1.2 - 2.4 * (k + 1) / n_turn for k in range(n_turn)] + [-1.2] * n_back
    x, scans, poses = 0.0, [], []
    for f, vel in enumerate(v):
        pose6 = np.array([x, 0.4 * np.sin(0.05 * f), 0.0, 0.03 * np.sin(0.08 * f), 0.0, 0.0])
        scans.append(synth.scan(seed, f, pose6, n_beams, n_az))
        poses.append(synth.pose_matrix(pose6))
        x += vel
    return scans, poses
