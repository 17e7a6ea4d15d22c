"""Host logic of the replay driver (lv_slam_b200/pipeline.py) on the CPU restatement: keyframe gate, constant-velocity guess,
loop candidates and validation, graph construction.  No device needed."""
import numpy as np

import oracle_ndt as O
import pipeline_backends as B
from lv_slam_b200 import pipeline as PL


def test_replay_driver_on_the_cpu_chain(tmp_path):
    scans, truth = B.out_and_back()
    r = PL.replay(scans, B.OracleRegistration(O.VAR_PCA, O.DIRECT1), B.OracleRegistration(O.VAR_OMP, O.DIRECT7), B.OracleGraphSLAM("lm_var_cholmod"),
                  B.OracleInformation(fitness_score_thresh=2.0), prefilter=B.OraclePrefilter(), dump_directory=str(tmp_path / "dump"))
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
    # the dump is the reference's artefact set and reads back as the same graph and keyframes
    from lv_slam_b200 import keyframe_io as K
    from lv_slam_b200.graph_slam import load_kitti_poses
    g2 = B.OracleGraphSLAM("lm_var_cholmod")
    back = K.load_dump(str(tmp_path / "dump"), g2)
    assert len(back) == len(kf) and g2.num_edges() == len(kf) - 1 + len(r["loops"])
    assert all(e.kernel == ("Huber", 1.0) for e in g2._edges)
    for b, T in zip(back, r["optimized"]):
        # KeyFrame::load overwrites the vertex with the 6-significant-digit estimate of `data` (keyframe.cpp:191-194)
        np.testing.assert_allclose(b["node"].estimate(), T, rtol=1e-5, atol=1e-5)
    wf = load_kitti_poses(str(tmp_path / "dump" / "ggo_wf_odom.txt"))
    assert len(wf) == len(scans)
    e_wf = max(np.linalg.norm((T0 @ truth[f])[:3, 3] - (np.linalg.inv(wf[0]) @ wf[f])[:3, 3]) for f in range(len(scans)))
    assert e_wf < 1.5
