"""Replay of the dlo_lfa_ggo chain on recorded scans (BASELINE configs[4] at any scale): prefilter -> scan-to-keyframe NDT
odometry -> keyframe graph with odometry and loop-closure edges -> pose-graph optimisation.

This is host control flow only — the nodelets' bookkeeping restated without ROS — over objects that carry the reference's method
names, so the same driver runs on the CUDA path (lv_slam_b200 classes) and on the CPU oracle (tests/): the replay is how the
pieces of the hot path are tested TOGETHER.  Restated, with the reference lines they follow:

  ScanMatchingOdometry.feed      ScanMatchingOdomNodelet::matching_s2k      src/lidar_odometry/scan_matching_odom_nodelet.cpp:192-262
  LoopDetector.find_candidates   LoopDetector::find_candidates              include/global_graph/loop_detector.hpp:104-139
  LoopDetector.matching          LoopDetector::matching                     include/global_graph/loop_detector.hpp:147-214
  GraphBuilder                   GlobalGraphNodelet::flush_keyframe_queue   src/global_graph/global_graph_nodelet.cpp:262-321
                                 + optimization_timer_callback              :660-742 (loop edges, Huber kernels, optimize)

Simplifications, on purpose: every odometry keyframe becomes a graph keyframe (the nodelet's KeyframeUpdater applies the same
10 m / 0.17 rad gate a second time), loop candidates are ranked by geometry only (the DBoW3 image ranking of matching_and_bow is
outside the path), and timing is frame-count based (10 Hz stamps).
"""
import numpy as np

DBL_MAX = float(np.finfo(np.float64).max)


def _rot_angle_f32(R):
    """2 * acos(Eigen::Quaternionf(R.cast<float>()).w()) as matching_s2k computes it (:237); w from the trace branch of Eigen's
    rotation-matrix-to-quaternion conversion."""
    Rf = np.asarray(R, dtype=np.float32)
    t = np.float32(Rf[0, 0] + Rf[1, 1] + Rf[2, 2])
    if t > 0:
        w = np.float32(0.5) * np.sqrt(np.float32(t + np.float32(1.0)))
    else:                                   # not reached for inter-keyframe rotations; generic fallback
        w = np.float32(np.cos(0.5 * np.arccos(np.clip((float(t) - 1.0) / 2.0, -1.0, 1.0))))
    return 2.0 * float(np.arccos(np.clip(np.float64(w), -1.0, 1.0)))


class ScanMatchingOdometry:
    """matching_s2k: scan-to-keyframe registration with a constant-velocity guess and the 10 m / 0.17 rad / 1 s keyframe gate
    (launch/dlo_lfa_ggo_kitti.launch:51-53).  `registration` needs setInputTarget / setInputSource / align(guess) /
    getFinalTransformation."""

    def __init__(self, registration, keyframe_delta_trans=10.0, keyframe_delta_angle=0.17, keyframe_delta_time=1.0):
        self.reg = registration
        self.dt, self.da, self.dtime = keyframe_delta_trans, keyframe_delta_angle, keyframe_delta_time
        self.scan_count = 0
        self.keyframes = []          # (frame index, odom pose 4x4 of the keyframe)
        self.aligns = 0

    def feed(self, stamp, cloud):
        if self.scan_count == 0:
            self.reg.setInputTarget(cloud)
            self.key_id = 0
            self.guess = np.eye(4)
            self.guess[0, 3] = 1.5                               # :199-200
            self.pre_tf_s2k = np.eye(4)
            self.key_pose = np.eye(4)
            self.keyframe_stamp = stamp
            self.keyframes.append((0, np.eye(4)))
            self.scan_count = 1
            return np.eye(4), True
        self.reg.setInputSource(cloud)
        self.reg.align(self.guess.astype(np.float32))
        self.aligns += 1
        tf_s2k = np.asarray(self.reg.getFinalTransformation(), dtype=np.float64)
        if self.scan_count == 1:                                 # the first pair is aligned twice (:222-226)
            self.reg.align(tf_s2k.astype(np.float32))
            self.aligns += 1
            tf_s2k = np.asarray(self.reg.getFinalTransformation(), dtype=np.float64)
        tf_s2s = np.linalg.inv(self.pre_tf_s2k) @ tf_s2k
        odom = self.key_pose @ tf_s2k
        dx = float(np.linalg.norm(tf_s2k[:3, 3]))
        da = _rot_angle_f32(tf_s2k[:3, :3])
        is_key = dx > self.dt or da > self.da or (stamp - self.keyframe_stamp) > self.dtime
        if is_key:
            self.reg.setInputTarget(cloud)
            self.key_id = self.scan_count
            tf_s2k = np.eye(4)
            self.key_pose = odom
            self.keyframe_stamp = stamp
            self.keyframes.append((self.scan_count, odom.copy()))
        self.pre_tf_s2k = tf_s2k
        self.guess = self.pre_tf_s2k @ tf_s2s
        self.scan_count += 1
        return odom, is_key


class LoopDetector:
    """Loop candidates by travelled and Euclidean distance, validation by registration + fitness score
    (loop_detector.hpp:104-214; parameters launch/dlo_lfa_ggo_kitti.launch:104-107)."""

    def __init__(self, registration, distance_thresh=20.0, accum_distance_thresh=100.0, min_edge_interval=50.0, fitness_score_thresh=2.0,
                 fitness_score_max_range=DBL_MAX):
        self.reg = registration
        self.distance_thresh, self.accum_distance_thresh = distance_thresh, accum_distance_thresh
        self.distance_from_last_edge_thresh, self.fitness_score_thresh = min_edge_interval, fitness_score_thresh
        self.fitness_score_max_range = fitness_score_max_range
        self.last_edge_accum_distance = 0.0
        self.aligns = 0

    def find_candidates(self, keyframes, new_kf):
        if new_kf["accum_distance"] - self.last_edge_accum_distance < self.distance_from_last_edge_thresh:
            return []
        out = []
        for k in keyframes:
            if new_kf["accum_distance"] - k["accum_distance"] < self.accum_distance_thresh:
                continue
            if np.linalg.norm(k["estimate"][:2, 3] - new_kf["estimate"][:2, 3]) > self.distance_thresh:
                continue
            out.append(k)
        return out

    def matching(self, candidates, new_kf):
        if not candidates:
            return None
        self.reg.setInputTarget(new_kf["cloud"])
        best_score, best, rel = DBL_MAX, None, None
        for c in candidates:
            self.reg.setInputSource(c["cloud"])
            guess = (np.linalg.inv(new_kf["estimate"]) @ c["estimate"]).astype(np.float32)
            guess[2, 3] = 0.0
            self.reg.align(guess)
            self.aligns += 1
            score = self.reg.getFitnessScore(self.fitness_score_max_range)
            if not self.reg.hasConverged() or score > best_score:
                continue
            best_score, best, rel = score, c, np.asarray(self.reg.getFinalTransformation(), dtype=np.float64)
        if best is None or best_score > self.fitness_score_thresh:
            return None
        self.last_edge_accum_distance = new_kf["accum_distance"]
        return dict(key1=new_kf, key2=best, relative_pose=rel, score=best_score)


def replay(scans, odom_registration, loop_registration, graph_slam, info_calc, prefilter=None, loop_params=None, optimize_iterations=512,
           stamp_step=0.1, dump_directory=None, tf_velo2cam=None):
    """Runs the chain over `scans` (list of float32 [n, >=3]).  Returns a dict with the odometry poses, the keyframe list, the loop
    edges and the optimised keyframe poses.  With `dump_directory` the reference's dump_service artefacts are written at the end
    (keyframe_io.dump: graph.g2o, keyframe directories, ggo_kf_odom.txt / ggo_wf_odom.txt)."""
    odo = ScanMatchingOdometry(odom_registration)
    det = LoopDetector(loop_registration, **(loop_params or {}))
    odom_poses, keyframes, loops = [], [], []
    accum = 0.0
    for f, raw in enumerate(scans):
        cloud = prefilter.filter(raw) if prefilter is not None else raw
        cloud = np.ascontiguousarray(np.asarray(cloud)[:, :3], dtype=np.float32)
        odom, is_key = odo.feed(f * stamp_step, cloud)
        odom_poses.append(odom)
        if not is_key:
            continue
        # flush_keyframe_queue: a node per keyframe, an odometry edge to the previous one (global_graph_nodelet.cpp:275-304)
        node = graph_slam.add_se3_node(odom)
        if keyframes:
            prev = keyframes[-1]
            accum += float(np.linalg.norm(odom[:3, 3] - prev["odom"][:3, 3]))
        kf = dict(frame=f, odom=odom.copy(), estimate=odom.copy(), cloud=cloud, node=node, accum_distance=accum)
        if keyframes:
            prev = keyframes[-1]
            rel = np.linalg.inv(kf["odom"]) @ prev["odom"]                       # (new, prev, new.odom^-1 * prev.odom)  :297-299
            info = info_calc.calc_information_matrix(prev["cloud"], kf["cloud"], rel)
            e = graph_slam.add_se3_edge(kf["node"], prev["node"], rel, info)
            graph_slam.add_robust_kernel(e, "Huber", 1.0)
        # optimization_timer_callback: loop detection for the new keyframe, then the edge with its own information matrix (:672-705)
        loop = det.matching(det.find_candidates(keyframes, kf), kf)
        keyframes.append(kf)
        if loop:
            info = info_calc.calc_information_matrix(loop["key1"]["cloud"], loop["key2"]["cloud"], loop["relative_pose"])
            e = graph_slam.add_se3_edge(loop["key1"]["node"], loop["key2"]["node"], loop["relative_pose"], info)
            graph_slam.add_robust_kernel(e, "Huber", 1.0)
            loops.append((loop["key1"]["frame"], loop["key2"]["frame"], loop["score"]))
    iters = graph_slam.optimize(optimize_iterations) if len(keyframes) > 1 else -1
    est = [np.asarray(k["node"].estimate(), dtype=np.float64) for k in keyframes]
    if dump_directory is not None:
        from . import keyframe_io
        recs = [dict(stamp=divmod(int(round(k["frame"] * stamp_step * 1e9)), 10 ** 9), seq=k["frame"], estimate=e, odom=k["odom"],
                     accum_distance=k["accum_distance"], id=k["node"].id(), cloud=k["cloud"]) for k, e in zip(keyframes, est)]
        keyframe_io.dump(dump_directory, graph_slam, recs, dict(enumerate(odom_poses)), tf_velo2cam)
    return dict(odom=odom_poses, keyframe_frames=[k["frame"] for k in keyframes], loops=loops, optimized=est, iterations=iters,
                odom_aligns=odo.aligns, loop_aligns=det.aligns)
