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


def pose7_batch(T):
    """pose7 of a stack of matrices [n, 4, 4] -> [n, 7]; the same branches and operations as _quat_from_R, taken per matrix."""
    T = np.asarray(T, dtype=np.float64).reshape(-1, 4, 4)
    R = T[:, :3, :3]
    n = len(T)
    q = np.zeros((n, 4))
    tr = R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]
    pos = tr > 0
    if pos.any():
        Rp = R[pos]
        t = np.sqrt(tr[pos] + 1.0)
        w = 0.5 * t
        t = 0.5 / t
        q[pos] = np.stack([(Rp[:, 2, 1] - Rp[:, 1, 2]) * t, (Rp[:, 0, 2] - Rp[:, 2, 0]) * t, (Rp[:, 1, 0] - Rp[:, 0, 1]) * t, w], axis=1)
    if not pos.all():
        idx = np.zeros(n, dtype=np.int64)
        idx[R[:, 1, 1] > R[:, 0, 0]] = 1
        d = R[np.arange(n), idx, idx]
        idx[R[:, 2, 2] > d] = 2
        for i in range(3):
            m = (~pos) & (idx == i)
            if not m.any():
                continue
            j, k = (i + 1) % 3, (i + 2) % 3
            Rm = R[m]
            t = np.sqrt(Rm[:, i, i] - Rm[:, j, j] - Rm[:, k, k] + 1.0)
            qm = np.zeros((len(Rm), 4))
            qm[:, i] = 0.5 * t
            t = 0.5 / t
            qm[:, 3] = (Rm[:, k, j] - Rm[:, j, k]) * t
            qm[:, j] = (Rm[:, j, i] + Rm[:, i, j]) * t
            qm[:, k] = (Rm[:, k, i] + Rm[:, i, k]) * t
            q[m] = qm
    q /= np.sqrt((q * q).sum(axis=1))[:, None]
    return np.concatenate([T[:, :3, 3], q], axis=1)


def matrix(p7):
    T = np.eye(4)
    q = np.asarray(p7[3:7], dtype=np.float64)
    T[:3, :3] = _R_from_quat(q / np.linalg.norm(q))
    T[:3, 3] = p7[:3]
    return T


def matrix_batch(p7):
    """matrix() of a stack of 7-vectors [n, 7] -> [n, 4, 4]."""
    p7 = np.asarray(p7, dtype=np.float64).reshape(-1, 7)
    q = p7[:, 3:7] / np.sqrt((p7[:, 3:7] * p7[:, 3:7]).sum(axis=1))[:, None]
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    T = np.zeros((len(p7), 4, 4))
    T[:, 0, 0] = 1 - 2 * (y * y + z * z); T[:, 0, 1] = 2 * (x * y - z * w); T[:, 0, 2] = 2 * (x * z + y * w)
    T[:, 1, 0] = 2 * (x * y + z * w); T[:, 1, 1] = 1 - 2 * (x * x + z * z); T[:, 1, 2] = 2 * (y * z - x * w)
    T[:, 2, 0] = 2 * (x * z - y * w); T[:, 2, 1] = 2 * (y * z + x * w); T[:, 2, 2] = 1 - 2 * (x * x + y * y)
    T[:, :3, 3] = p7[:, :3]
    T[:, 3, 3] = 1.0
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
