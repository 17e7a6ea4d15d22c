"""Synthetic inputs: Velodyne-like scans ray-cast by ``synth.c`` and sphere pose graphs.

Input manufacture only — the same arrays feed the CUDA path, the oracle and the CPU baseline.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsynth.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "synth.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["/usr/bin/gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-o", _SO, src, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.synth_scan.restype = ctypes.c_int
        _lib.synth_scan.argtypes = [ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_double, ctypes.c_void_p, ctypes.c_int]
        _lib.synth_traj_pose.restype = None
        _lib.synth_traj_pose.argtypes = [ctypes.c_int, ctypes.c_void_p]
    return _lib


def scan(seed, frame, pose6, n_beams=64, n_az=2000, noise_sigma=0.02, nthreads=None):
    """Ray-cast one scan from ``pose6 = (x, y, z, yaw, pitch, roll)``; returns float32 [n, 3] in the sensor frame."""
    lib = _load()
    pose = np.ascontiguousarray(pose6, dtype=np.float64)
    out = np.empty((n_beams * n_az, 3), dtype=np.float32)
    nthreads = nthreads or min(16, os.cpu_count() or 1)
    n = lib.synth_scan(int(seed), int(frame), pose.ctypes.data, n_beams, n_az, float(noise_sigma), out.ctypes.data, nthreads)
    return np.ascontiguousarray(out[:n])


def traj_pose(frame):
    lib = _load()
    pose = np.zeros(6, dtype=np.float64)
    lib.synth_traj_pose(int(frame), pose.ctypes.data)
    return pose


def pose_matrix(pose6):
    """4x4 world<-sensor matrix of a pose6, R = Rz(yaw) Ry(pitch) Rx(roll)."""
    x, y, z, yaw, pitch, roll = [float(v) for v in pose6]
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    T = np.eye(4)
    T[:3, :3] = [[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                 [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                 [-sp, cp * sr, cp * cr]]
    T[:3, 3] = [x, y, z]
    return T


def config1_pair(n_beams=64, n_az=2000, seed=42):
    """BASELINE config 1 stand-in (no KITTI data here): target at identity, source displaced by
    (1.20, 0.05, 0.01) m, yaw 1.0 deg, pitch 0.2 deg; the reference's first-frame guess x = 1.5 m
    (src/lidar_odometry/scan_matching_odom_nodelet.cpp:199-200).  Returns (target, source, guess4x4, truth4x4)."""
    p0 = np.zeros(6)
    p1 = np.array([1.20, 0.05, 0.01, np.deg2rad(1.0), np.deg2rad(0.2), 0.0])
    tgt = scan(seed, 0, p0, n_beams, n_az)
    src = scan(seed, 1, p1, n_beams, n_az)
    guess = np.eye(4, dtype=np.float32)
    guess[0, 3] = 1.5
    truth = np.linalg.inv(pose_matrix(p0)) @ pose_matrix(p1)
    return tgt, src, guess, truth


def stream(n_frames, seed=1000, n_beams=64, n_az=2000, start=0):
    """Frames ``start .. start+n_frames-1`` of the synthetic drive (configs 2 / 5): returns (scans, poses4x4)."""
    scans, poses = [], []
    for f in range(start, start + n_frames):
        p = traj_pose(f)
        scans.append(scan(seed, f, p, n_beams, n_az))
        poses.append(pose_matrix(p))
    return scans, poses


def keyframe_plan(poses, delta_trans=10.0, delta_angle=0.17, delta_frames=10):
    """Scan-to-keyframe schedule of ``matching_s2k`` (src/lidar_odometry/scan_matching_odom_nodelet.cpp:236-250) replayed on
    known poses: frame 0 is the first keyframe; a frame whose motion from the current keyframe exceeds 10 m / 0.17 rad /
    1 s (10 frames at 10 Hz; launch/dlo_lfa_ggo_kitti.launch:51-53) becomes the next keyframe after being matched.
    Returns a list of (frame, key_frame, guess4x4) for frames >= 2; the guess is the constant-velocity prediction
    key^-1 * prev * (prevprev^-1 * prev)."""
    plan = []
    key = 0
    for f in range(1, len(poses)):
        if f >= 2:
            vel = np.linalg.inv(poses[f - 2]) @ poses[f - 1]
            guess = np.linalg.inv(poses[key]) @ poses[f - 1] @ vel
            plan.append((f, key, guess.astype(np.float32)))
        rel = np.linalg.inv(poses[key]) @ poses[f]
        dx = np.linalg.norm(rel[:3, 3])
        da = np.arccos(np.clip((np.trace(rel[:3, :3]) - 1) / 2, -1, 1))
        if dx > delta_trans or da > delta_angle or (f - key) > delta_frames:
            key = f
    return plan
