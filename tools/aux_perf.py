"""Timing of the stages either side of the hot path on full-size scans: prefilter (distance filter + 0.1 m VoxelGrid) and
getFitnessScore, device vs the CPU restatement."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import torch
import lv_slam_b200 as L
import oracle_ndt as O
from lv_slam_b200 import synth

tgt, src, guess, truth = synth.config1_pair()
cloud = np.concatenate([tgt[:, :3], np.random.default_rng(3).random((len(tgt), 1), dtype=np.float32)], axis=1).astype(np.float32)


def timed(f, reps):
    f(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps * 1e3


pf = L.Prefilter(distance_near_thresh=0.5, distance_far_thresh=100.0, downsample_resolution=0.1)
out = pf.filter(cloud)
t_host = timed(lambda: pf.filter(cloud), 20)
dcloud = torch.from_numpy(cloud).cuda()
t_dev = timed(lambda: pf.filter(dcloud), 20)
t = time.perf_counter(); exp, _ = O.prefilter(cloud, 0.5, 100.0, True, 0.1); t_cpu = (time.perf_counter() - t) * 1e3
print("prefilter %d -> %d points | GPU %.3f ms from host buffers, %.3f ms resident | CPU restatement %.1f ms | identical %s" % (
    len(cloud), len(out), t_host, t_dev, t_cpu, bool(np.array_equal(out, exp))))

n = L.NormalDistributionsTransform()
n.setInputTarget(tgt); n.setInputSource(src)
big = float(np.finfo(np.float64).max)
s, c = n.getFitnessScore(big, T=truth, with_count=True)
t_fit = timed(lambda: n.getFitnessScore(big, T=truth), 50)
t_fit2 = timed(lambda: n.getFitnessScore(big, T=guess), 50)
o = O.OracleNDT(num_threads=os.cpu_count() or 1)
sub = src[:: max(1, len(src) // 4000)]
o.set_target(tgt); o.set_source(sub)
t = time.perf_counter(); so, co = o.fitness_score(truth, big); t_cpu = (time.perf_counter() - t) * 1e3
n.setInputSource(sub)
sg, cg = n.getFitnessScore(big, T=truth, with_count=True)
print("getFitnessScore %d x %d points: GPU %.3f ms (truth pose), %.3f ms (first guess), score %.6f over %d correspondences | CPU exhaustive scan of a %d-point sample %.1f ms "
      "(x%.0f for the full scan, %d threads), same sample on the GPU: |diff| %.1e, counts %d / %d" % (
          len(src), len(tgt), t_fit, t_fit2, s, c, len(sub), t_cpu, len(src) / len(sub), os.cpu_count() or 1, abs(sg - so), cg, co))
