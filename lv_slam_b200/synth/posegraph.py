"""Synthetic SE(3) pose graphs: a restatement of g2o's sphere generator
(3rdtools/g2o-a48ff8c.zip!g2o/g2o/examples/sphere/create_sphere.cpp:45-215) with a seeded numpy RNG in place of g2o's sampler.

Input manufacture only — the same arrays feed the CUDA path, the oracle and the CPU baseline.  Poses and measurements are
7-vectors ``x y z qx qy qz qw``; information matrices are the 21 upper-triangular entries, row-major, as in g2o files.
"""
import numpy as np


def _quat_from_R(R):
    """Eigen::Quaterniond(Matrix3d)."""
    t = R[0, 0] + R[1, 1] + R[2, 2]
    if t > 0:
        t = np.sqrt(t + 1.0)
        w = 0.5 * t
        t = 0.5 / t
        return np.array([(R[2, 1] - R[1, 2]) * t, (R[0, 2] - R[2, 0]) * t, (R[1, 0] - R[0, 1]) * t, w])
    i = 0
    if R[1, 1] > R[0, 0]:
        i = 1
    if R[2, 2] > R[i, i]:
        i = 2
    j, k = (i + 1) % 3, (i + 2) % 3
    t = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
    q = np.zeros(4)
    q[i] = 0.5 * t
    t = 0.5 / t
    q[3] = (R[k, j] - R[j, k]) * t
    q[j] = (R[j, i] + R[i, j]) * t
    q[k] = (R[k, i] + R[i, k]) * t
    return q


def _R_from_quat(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz, aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz])


def pose7(T):
    q = _quat_from_R(T[:3, :3])
    q /= np.linalg.norm(q)
    return np.concatenate([T[:3, 3], q])


def matrix(p7):
    T = np.eye(4)
    q = np.asarray(p7[3:7], dtype=np.float64)
    T[:3, :3] = _R_from_quat(q / np.linalg.norm(q))
    T[:3, 3] = p7[:3]
    return T


def info21_diag(d6):
    M = np.diag(np.asarray(d6, dtype=np.float64))
    return np.array([M[r, c] for r in range(6) for c in range(r, 6)])


def sphere(nodes_per_level=100, laps=50, radius=100.0, sigma_t=0.01, sigma_r=0.005, seed=7, information=(2, 2, 2, 10, 10, 10), huber=1.0):
    """Returns dict(poses7 initial estimate by odometry chaining, truth7, ij int32 [E,2], meas7, info21, huber)."""
    rng = np.random.default_rng(seed)
    n = nodes_per_level * laps
    truth = []
    vid = 0
    for f in range(laps):
        for k in range(nodes_per_level):
            vid += 1
            az = -np.pi + 2 * k * np.pi / nodes_per_level
            ay = -0.5 * np.pi + vid * np.pi / (laps * nodes_per_level)
            Rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
            Ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
            T = np.eye(4)
            T[:3, :3] = Rz @ Ry
            T[:3, 3] = T[:3, :3] @ np.array([radius, 0, 0])
            truth.append(T)
    ij = [(i - 1, i) for i in range(1, n)]
    for f in range(1, laps):
        for nn in range(nodes_per_level):
            frm = (f - 1) * nodes_per_level + nn
            for d in (-1, 0, 1):
                if f == laps - 1 and d == 1:
                    continue
                ij.append((frm, f * nodes_per_level + nn + d))
    ij = np.array(ij, dtype=np.int32)
    meas = np.zeros((len(ij), 7))
    for e, (a, b) in enumerate(ij):
        T = np.linalg.inv(truth[a]) @ truth[b]
        gq = _quat_from_R(T[:3, :3])
        qxyz = rng.normal(0.0, sigma_r, 3)
        qw = max(0.0, 1.0 - np.linalg.norm(qxyz))
        rot = np.array([qxyz[0], qxyz[1], qxyz[2], qw])
        rot /= np.linalg.norm(rot)
        q = _quat_mul(gq, rot)
        meas[e, :3] = T[:3, 3] + rng.normal(0.0, sigma_t, 3)
        meas[e, 3:] = q / np.linalg.norm(q)
    est = [truth[0]]
    for e in range(n - 1):           # odometry chaining: EdgeSE3::initialEstimate
        est.append(est[-1] @ matrix(meas[e]))
    poses7 = np.array([pose7(T) for T in est])
    truth7 = np.array([pose7(T) for T in truth])
    info = np.tile(info21_diag(information), (len(ij), 1))
    hub = np.full(len(ij), float(huber))
    return dict(poses7=poses7, truth7=truth7, ij=ij, meas7=meas, info21=info, huber=hub)
