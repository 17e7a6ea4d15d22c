"""On-disk artefacts of the global graph, so a replay can consume and produce the reference's own files.

Host-side text/binary formats only; nothing here touches the device.

  save_keyframe / load_keyframe   KeyFrame::save / KeyFrame::load            src/global_graph/keyframe.cpp:48-92, 94-200
  save_pcd_binary / load_pcd      pcl::io::savePCDFileBinary (PointXYZI)     keyframe.cpp:91; PCD v0.7, DATA binary
  load_calib                      the `Tr:` line of a KITTI calib.txt        global_graph_nodelet.cpp:1081-1088
  dump                            GlobalGraphNodelet::dump_service           global_graph_nodelet.cpp:979-1027
  save_pose                       GlobalGraphNodelet::save_pose              global_graph_nodelet.cpp:1077-1147
"""
import os

import numpy as np

from .graph_slam import save_kitti_poses
from .synth import posegraph as _pg


def eigen_matrix_text(M):
    """`ofs << M` with Eigen's default IOFormat on a default stream: 6 significant digits, every coefficient right-aligned to the
    widest one, one space between columns, rows on their own lines (no trailing newline)."""
    M = np.atleast_2d(np.asarray(M, dtype=np.float64))
    s = [["%g" % x for x in row] for row in M]
    w = max(len(x) for row in s for x in row)
    return "\n".join(" ".join(x.rjust(w) for x in row) for row in s)


def save_pcd_binary(path, cloud):
    """cloud: float32 [n, 3] (x y z) or [n, 4] (x y z intensity).  The writer packs the declared fields only (no SSE padding)."""
    c = np.ascontiguousarray(np.asarray(cloud, dtype=np.float32))
    if c.ndim != 2 or c.shape[1] not in (3, 4):
        raise ValueError("cloud must be [n, 3] or [n, 4]")
    n, k = c.shape
    fields = "x y z intensity" if k == 4 else "x y z"
    ones = " ".join(["1"] * k)
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS %s\nSIZE %s\nTYPE %s\nCOUNT %s\nWIDTH %d\nHEIGHT 1\n"
           "VIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary\n") % (fields, " ".join(["4"] * k), " ".join(["F"] * k), ones, n, n)
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii"))
        f.write(c.tobytes())


def load_pcd(path):
    """Reads `DATA binary` and `DATA ascii` files whose fields are 4-byte floats; returns float32 [n, 4] (x y z intensity, the
    intensity 0 when the file has none).  Fields other than x y z intensity are skipped."""
    with open(path, "rb") as f:
        raw = f.read()
    pos, meta = 0, {}
    while True:
        end = raw.index(b"\n", pos)
        line = raw[pos:end].decode("ascii", "replace").strip()
        pos = end + 1
        if not line or line.startswith("#"):
            continue
        key, _, rest = line.partition(" ")
        meta[key] = rest.split()
        if key == "DATA":
            break
    names = meta["FIELDS"]
    sizes = [int(s) for s in meta["SIZE"]]
    types = meta["TYPE"]
    counts = [int(c) for c in meta.get("COUNT", ["1"] * len(names))]
    n = int(meta["POINTS"][0]) if "POINTS" in meta else int(meta["WIDTH"][0]) * int(meta.get("HEIGHT", ["1"])[0])
    out = np.zeros((n, 4), np.float32)
    want = {"x": 0, "y": 1, "z": 2, "intensity": 3}
    kind = meta["DATA"][0]
    if kind == "binary":
        dt, off = [], 0
        for nm, sz, ty, ct in zip(names, sizes, types, counts):
            code = {("F", 4): "<f4", ("F", 8): "<f8", ("U", 1): "u1", ("U", 2): "<u2", ("U", 4): "<u4", ("I", 1): "i1", ("I", 2): "<i2",
                    ("I", 4): "<i4"}[(ty, sz)]
            dt.append((nm if nm != "_" else "_pad%d" % off, code, (ct,)) if ct != 1 else (nm if nm != "_" else "_pad%d" % off, code))
            off += sz * ct
        rec = np.frombuffer(raw, dtype=np.dtype(dt), count=n, offset=pos)
        for nm, col in want.items():
            if nm in names:
                out[:, col] = rec[nm].astype(np.float32)
    elif kind == "ascii":
        tab = np.array(raw[pos:].split(), dtype=np.float64).reshape(n, sum(counts))
        col0 = np.cumsum([0] + counts[:-1])
        for nm, col in want.items():
            if nm in names:
                out[:, col] = tab[:, col0[names.index(nm)]].astype(np.float32)
    else:
        raise ValueError("unsupported PCD DATA kind: " + kind)
    return out


def save_keyframe(directory, kf):
    """kf: dict with stamp (sec, nsec), estimate 4x4, odom 4x4, accum_distance, id, cloud; optional floor_coeffs (4), utm_coord
    (3), acceleration (3), orientation (w x y z).  Field order and spelling as KeyFrame::save writes them."""
    os.makedirs(directory, exist_ok=True)
    with open(os.path.join(directory, "data"), "w") as f:
        sec, nsec = kf.get("stamp", (0, 0))
        f.write("stamp %d %d\n" % (sec, nsec))
        f.write("estimate\n%s\n" % eigen_matrix_text(kf["estimate"]))
        f.write("odom\n%s\n" % eigen_matrix_text(kf["odom"]))
        f.write("accum_distance %g\n" % kf["accum_distance"])
        for key in ("floor_coeffs", "utm_coord", "acceleration"):
            if kf.get(key) is not None:
                f.write("%s %s\n" % (key, eigen_matrix_text(np.asarray(kf[key], dtype=np.float64).reshape(1, -1))))
        if kf.get("orientation") is not None:
            f.write("orientation %s\n" % " ".join("%g" % x for x in kf["orientation"]))
        if kf.get("id") is not None:
            f.write("id %d\n" % kf["id"])
    save_pcd_binary(os.path.join(directory, "cloud.pcd"), kf["cloud"])


def load_keyframe(directory):
    """Token reader like KeyFrame::load; returns None when `data` is missing or holds no node id (the reference's `return false`)."""
    try:
        with open(os.path.join(directory, "data")) as f:
            tok = f.read().split()
    except OSError:
        return None
    kf = dict(stamp=(0, 0), estimate=None, odom=np.eye(4), accum_distance=-1.0, id=None)
    i = 0

    def take(n):
        nonlocal i
        v = np.array(tok[i:i + n], dtype=np.float64)
        i += n
        return v

    while i < len(tok):
        t = tok[i]
        i += 1
        if t == "stamp":
            kf["stamp"] = (int(tok[i]), int(tok[i + 1]))
            i += 2
        elif t in ("estimate", "odom"):
            M = take(16).reshape(4, 4)
            T = np.eye(4)
            T[:3, :4] = M[:3, :4]
            kf[t] = T
        elif t == "accum_distance":
            kf[t] = float(take(1)[0])
        elif t == "floor_coeffs":
            kf[t] = take(4)
        elif t in ("utm_coord", "acceleration"):
            kf[t] = take(3)
        elif t == "orientation":
            kf[t] = take(4)
        elif t == "id":
            kf["id"] = int(tok[i])
            i += 1
    if kf["id"] is None or kf["id"] < 0:
        return None
    kf["cloud"] = load_pcd(os.path.join(directory, "cloud.pcd"))
    return kf


def load_calib(calib_file):
    """tf_velo2cam from the fifth line (`Tr: r00 r01 ... t2`) of a KITTI odometry calib.txt."""
    with open(calib_file) as f:
        lines = f.read().splitlines()
    v = lines[4].split()[1:13]
    T = np.eye(4)
    T[:3, :4] = np.array(v, dtype=np.float64).reshape(3, 4)
    return T


def _slerp_from_identity(t, q1):
    """Eigen::Quaterniond::Identity().slerp(t, q1), coefficients (w, x, y, z)."""
    q0 = np.array([1.0, 0.0, 0.0, 0.0])
    d = float(q0 @ q1)
    ad = abs(d)
    if ad >= 1.0 - np.finfo(np.float64).eps:
        s0, s1 = 1.0 - t, t
    else:
        th = np.arccos(ad)
        s0, s1 = np.sin((1.0 - t) * th) / np.sin(th), np.sin(t * th) / np.sin(th)
    if d < 0:
        s1 = -s1
    return s0 * q0 + s1 * q1


def _quat_wxyz(R):
    p = _pg.pose7(np.block([[R, np.zeros((3, 1))], [np.zeros((1, 3)), np.ones((1, 1))]]))      # (t, qx qy qz qw)
    return np.array([p[6], p[3], p[4], p[5]])


def _rot_wxyz(q):
    w, x, y, z = q                                                 # Eigen toRotationMatrix: no normalisation
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def save_pose(directory, keyframes, odoms, tf_velo2cam=None):
    """ggo_kf_odom.txt: the optimised keyframe poses in the camera frame.  ggo_wf_odom.txt: every frame, the keyframe's optimised
    pose carried along the odometry between keyframes plus a per-frame share of the correction the optimisation applied to the
    segment.  keyframes: list of dicts with `seq` and `estimate`; odoms: {seq: 4x4}.

    Restated as written, including: the slerp parameter is the frame count of the segment (`q0.slerp(seq1 - seq0, q1)`, :1118) and
    not its reciprocal, while the translation IS divided by it; frames of the last keyframe's segment run to `odoms.end()->first`
    (:1103), which is not defined by the language — the line above it in the reference names odoms.size(), which is what is used
    here."""
    C = np.eye(4) if tf_velo2cam is None else np.asarray(tf_velo2cam, dtype=np.float64)
    Ci = np.linalg.inv(C)
    os.makedirs(directory, exist_ok=True)
    save_kitti_poses(os.path.join(directory, "ggo_kf_odom.txt"), [C @ np.asarray(k["estimate"]) @ Ci for k in keyframes])
    rows = []
    if keyframes:
        align = np.linalg.inv(np.asarray(keyframes[0]["estimate"], dtype=np.float64))
    for i, k in enumerate(keyframes):
        seq0, seq1 = k["seq"], len(odoms)
        kf_pose = align @ np.asarray(k["estimate"], dtype=np.float64)
        if seq0 not in odoms:
            continue
        odom0 = np.asarray(odoms[seq0], dtype=np.float64)
        d_pose_odom = np.eye(4)
        if i < len(keyframes) - 1:
            nxt = keyframes[i + 1]
            d_pose = np.linalg.inv(kf_pose) @ (align @ np.asarray(nxt["estimate"], dtype=np.float64))
            seq1 = nxt["seq"]
            if seq1 not in odoms:
                continue
            d_odom = np.linalg.inv(odom0) @ np.asarray(odoms[seq1], dtype=np.float64)
            d_pose_odom = np.linalg.inv(d_odom) @ d_pose
            q = _slerp_from_identity(float(seq1 - seq0), _quat_wxyz(d_pose_odom[:3, :3]))
            d_pose_odom[:3, :3] = _rot_wxyz(q)
            d_pose_odom[:3, 3] *= 1.0 / (seq1 - seq0)
        for j in range(seq0, seq1):
            if j not in odoms:
                continue
            pose_s2k = np.linalg.inv(odom0) @ np.asarray(odoms[j], dtype=np.float64)
            pose_new = kf_pose @ pose_s2k if j == seq0 else kf_pose @ pose_s2k @ d_pose_odom
            rows.append(C @ pose_new @ Ci)
    save_kitti_poses(os.path.join(directory, "ggo_wf_odom.txt"), rows)
    return len(rows)


def dump(directory, graph_slam, keyframes, odoms=None, tf_velo2cam=None):
    """dump_service: graph.g2o (+ .kernels), one %06d directory per keyframe, special_nodes.csv, then save_pose.  The anchor and
    floor nodes belong to the GPS / floor constraints, which are outside the path: always -1."""
    os.makedirs(directory, exist_ok=True)
    graph_slam.save(os.path.join(directory, "graph.g2o"))
    for i, k in enumerate(keyframes):
        save_keyframe(os.path.join(directory, "%06d" % i), k)
    with open(os.path.join(directory, "special_nodes.csv"), "w") as f:
        f.write("anchor_node -1\nanchor_edge -1\nfloor_node -1\n")
    if odoms is not None:
        save_pose(directory, keyframes, odoms, tf_velo2cam)


def load_dump(directory, graph_slam):
    """The inverse of dump as the reference's map viewer reads it: graph.g2o into `graph_slam`, then every %06d directory in order;
    each keyframe's `node` is the loaded vertex with its id."""
    graph_slam.load(os.path.join(directory, "graph.g2o"))
    by_id = {v.id(): v for v in graph_slam._vertices}
    out = []
    i = 0
    while True:
        k = load_keyframe(os.path.join(directory, "%06d" % i))
        if k is None:
            break
        k["node"] = by_id.get(k["id"])
        if k["estimate"] is not None and k["node"] is not None:
            k["node"].setEstimate(k["estimate"])                                   # keyframe.cpp:193-196
        out.append(k)
        i += 1
    return out
