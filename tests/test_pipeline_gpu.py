"""The pieces of the path TOGETHER (BASELINE configs[4], the dlo_lfa_ggo chain, at a scale the CPU oracle replays in seconds):
prefilter -> scan-to-keyframe NDT odometry (pclpca / DIRECT1) -> keyframe graph -> loop-closure validation (pclomp / DIRECT7 +
getFitnessScore) -> information matrices -> LM with the direct solver.  The same host driver (lv_slam_b200/pipeline.py) runs on
the CUDA path through the C-ABI and on the CPU restatement; every decision (keyframes, loop edges) must be identical and the
trajectories must agree to the per-iteration tolerance of the north star (1e-4 m / 1e-5 rad)."""
import numpy as np
import pytest

import oracle_ndt as O
import pipeline_backends as B
from lv_slam_b200 import pipeline as PL

pytestmark = pytest.mark.gpu


def _angle(Ra, Rb):
    d = Ra[:3, :3].T @ Rb[:3, :3]
    v = 0.5 * np.array([d[2, 1] - d[1, 2], d[0, 2] - d[2, 0], d[1, 0] - d[0, 1]])
    return float(np.arcsin(min(1.0, np.linalg.norm(v))))


def test_replay_matches_the_cpu_chain():
    import lv_slam_b200 as L
    scans, truth = B.out_and_back()
    odo = L.NormalDistributionsTransform(variant=L.LVS_NDT_PCA)
    odo.setNeighborhoodSearchMethod(L.LVS_DIRECT1); odo.setTransformationEpsilon(0.01); odo.setMaximumIterations(64)
    loop = L.NormalDistributionsTransform(variant=L.LVS_NDT_OMP)
    loop.setNeighborhoodSearchMethod(L.LVS_DIRECT7); loop.setTransformationEpsilon(0.01); loop.setMaximumIterations(64)
    g = PL.replay(scans, odo, loop, L.GraphSLAM("lm_var_cholmod"), L.InformationMatrixCalculator(fitness_score_thresh=2.0),
                  prefilter=L.Prefilter(distance_near_thresh=0.5, distance_far_thresh=100.0, downsample_resolution=0.1))
    c = PL.replay(scans, B.OracleRegistration(O.VAR_PCA, O.DIRECT1), B.OracleRegistration(O.VAR_OMP, O.DIRECT7), B.OracleGraphSLAM("lm_var_cholmod"),
                  B.OracleInformation(fitness_score_thresh=2.0), prefilter=B.OraclePrefilter())
    # every decision of the chain
    assert g["keyframe_frames"] == c["keyframe_frames"] and len(g["keyframe_frames"]) >= 10
    assert [(a, b) for a, b, _ in g["loops"]] == [(a, b) for a, b, _ in c["loops"]] and len(g["loops"]) >= 1
    assert g["odom_aligns"] == c["odom_aligns"] == len(scans) and g["loop_aligns"] == c["loop_aligns"] >= 1
    for (_, _, sg), (_, _, sc) in zip(g["loops"], c["loops"]):
        assert abs(sg - sc) <= 1e-9 * sc and sg <= 2.0
    # odometry of every frame and the optimised keyframe poses
    for Tg, Tc in zip(g["odom"], c["odom"]):
        assert np.abs(Tg[:3, 3] - Tc[:3, 3]).max() <= 1e-4 and _angle(Tg, Tc) <= 1e-5
    assert g["iterations"] > 0 and c["iterations"] > 0
    for Tg, Tc in zip(g["optimized"], c["optimized"]):
        assert np.abs(Tg[:3, 3] - Tc[:3, 3]).max() <= 1e-4 and _angle(Tg, Tc) <= 1e-5
    # and the chain does its job: the drive is recovered to well under a voxel
    T0 = np.linalg.inv(truth[0])
    err = max(np.linalg.norm((T0 @ truth[f])[:3, 3] - T[:3, 3]) for f, T in zip(g["keyframe_frames"], g["optimized"]))
    assert err < 1.0
