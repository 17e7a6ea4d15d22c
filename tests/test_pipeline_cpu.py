"""Host logic of the replay driver (lv_slam_b200/pipeline.py) on the CPU restatement: keyframe gate, constant-velocity guess,
loop candidates and validation, graph construction.  No device needed."""
import numpy as np

import oracle_ndt as O
import pipeline_backends as B
from lv_slam_b200 import pipeline as PL


def test_replay_driver_on_the_cpu_chain():
    scans, truth = B.out_and_back()
    r = PL.replay(scans, B.OracleRegistration(O.VAR_PCA, O.DIRECT1), B.OracleRegistration(O.VAR_OMP, O.DIRECT7), B.OracleGraphSLAM("lm_var_cholmod"),
                  B.OracleInformation(), prefilter=B.OraclePrefilter())
    kf = r["keyframe_frames"]
    assert kf[0] == 0 and all(b > a for a, b in zip(kf, kf[1:])) and len(kf) >= 10
    # 1.2 m per frame: a keyframe once 10 m are exceeded (launch/dlo_lfa_ggo_kitti.launch:51), i.e. every 9th frame on the straight legs
    assert kf[1:5] == [9, 18, 27, 36]
    # one align per frame after the first, the first pair twice (scan_matching_odom_nodelet.cpp:222-226)
    assert r["odom_aligns"] == len(scans)
    # the return leg closes a loop onto an outbound keyframe: > 100 m travelled apart, < 20 m apart, fitness below 2.0
    assert len(r["loops"]) >= 1
    new, old, score = r["loops"][0]
    assert new > old and score <= 2.0
    T0 = np.linalg.inv(truth[0])
    assert np.linalg.norm((T0 @ truth[new])[:3, 3] - (T0 @ truth[old])[:3, 3]) < 20.0
    assert r["iterations"] > 0
    err = max(np.linalg.norm((T0 @ truth[f])[:3, 3] - T[:3, 3]) for f, T in zip(kf, r["optimized"]))
    assert err < 1.0
    # the rotation gate of matching_s2k: 2 acos(w) of the float quaternion
    c, s = np.cos(0.2), np.sin(0.2)
    assert abs(PL._rot_angle_f32(np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])) - 0.2) < 1e-3
