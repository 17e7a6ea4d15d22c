"""Host mirror of ``lv_slam::GraphSLAM`` (/root/reference/include/global_graph/graph_slam.hpp:40-149,
src/global_graph/graph_slam.cpp) for the part of it the hot path covers: SE(3) nodes, SE(3) edges with optional robust kernels,
``optimize``, g2o-text ``save`` / ``load``.  Vertices and edges are light Python handles (the reference hands out raw
``g2o::VertexSE3*`` / ``g2o::EdgeSE3*``); all arithmetic happens in liblvslam_b200.so on the GPU.
"""
import ctypes

import numpy as np

from . import _capi as C
from .synth import posegraph as _pg

_SOLVERS = {"lm_var": C.LVS_PGO_LM_CHOL, "lm_var_cholmod": C.LVS_PGO_LM_CHOL, "lm_var_csparse": C.LVS_PGO_LM_CHOL, "lm_fix6_3": C.LVS_PGO_LM_CHOL,
            "lm_fix6_3_cholmod": C.LVS_PGO_LM_CHOL, "lm_fix6_3_csparse": C.LVS_PGO_LM_CHOL, "gn_var": C.LVS_PGO_GN_CHOL,
            "gn_var_cholmod": C.LVS_PGO_GN_CHOL, "gn_var_csparse": C.LVS_PGO_GN_CHOL, "gn_fix6_3": C.LVS_PGO_GN_CHOL,
            "lm_pcg": C.LVS_PGO_LM_PCG, "gn_pcg": C.LVS_PGO_GN_PCG}


class VertexSE3:
    def __init__(self, vid, pose4x4):
        self._id = vid
        self._T = np.array(pose4x4, dtype=np.float64).reshape(4, 4)
        self._fixed = False

    def id(self):
        return self._id

    def estimate(self):
        return self._T.copy()

    def setEstimate(self, T):
        self._T = np.array(T, dtype=np.float64).reshape(4, 4)

    def setFixed(self, f):
        self._fixed = bool(f)

    def fixed(self):
        return self._fixed


class EdgeSE3:
    def __init__(self, v1, v2, rel4x4, info6x6):
        self.vertices = [v1, v2]
        self.measurement = np.array(rel4x4, dtype=np.float64).reshape(4, 4)
        self.information = np.array(info6x6, dtype=np.float64).reshape(6, 6)
        self.kernel = None          # (type, delta)


# unary priors on a VertexSE3 (include/g2o/edge_se3_prior{xy,xyz,quat,vec}.hpp): kind -> (LVS_PGO_EDGE_* code, g2o tag, measurement length)
EDGE_SE3, PRIOR_XY, PRIOR_XYZ, PRIOR_QUAT, PRIOR_VEC, SE3_PLANE = 0, 1, 2, 3, 4, 5
_PRIOR_TAGS = {PRIOR_XY: ("EDGE_SE3_PRIORXY", 2, 2), PRIOR_XYZ: ("EDGE_SE3_PRIORXYZ", 3, 3), PRIOR_QUAT: ("EDGE_SE3_PRIORQUAT", 4, 3),
               PRIOR_VEC: ("EDGE_SE3_PRIORVEC", 6, 3), SE3_PLANE: ("EDGE_SE3_PLANE", 4, 3)}


class VertexPlane:
    """g2o::VertexPlane as the floor detection uses it: one node, fixed at creation (global_graph_nodelet.cpp:601-604).  The estimate is a
    Plane3D: the four coefficients scaled to a unit normal."""

    def __init__(self, vid, coeffs):
        c = np.array(coeffs, dtype=np.float64).ravel()
        self._id, self._c, self._fixed = vid, c / np.linalg.norm(c[:3]), False

    def id(self):
        return self._id

    def estimate(self):
        return self._c.copy()

    def setFixed(self, f):
        self._fixed = bool(f)

    def fixed(self):
        return self._fixed


class EdgeSE3Prior:
    """One of EdgeSE3PriorXY / XYZ / Quat / Vec.  measurement: xy | xyz | (qx, qy, qz, qw) | direction(3) + measurement(3), stored as
    the edge's setMeasurement leaves it (PriorQuat: w >= 0; PriorVec: both halves normalised)."""

    def __init__(self, kind, v, measurement, information):
        self.kind = kind
        self.vertices = [v, v]
        m = np.array(measurement, dtype=np.float64).ravel()
        if kind == PRIOR_QUAT and m[3] < 0:
            m = -m
        if kind == PRIOR_VEC:
            m = np.concatenate([m[:3] / np.linalg.norm(m[:3]), m[3:] / np.linalg.norm(m[3:])])
        if kind == SE3_PLANE:
            m = m / np.linalg.norm(m[:3])
        self.measurement = m
        d = _PRIOR_TAGS[kind][2]
        self.information = np.array(information, dtype=np.float64).reshape(d, d)
        self.kernel = None
        self.plane = None           # SE3_PLANE: the VertexPlane


class GraphSLAM:
    def __init__(self, solver_type="lm_var", device=0):
        if solver_type not in _SOLVERS:
            raise ValueError("unknown solver type %r (the reference prints g2o's solver list here)" % solver_type)
        self._L = C.lib()
        self._solver_type = solver_type
        self._device = device
        self._h = None                      # the device object is created on first optimize()
        self._vertices, self._edges, self._planes = [], [], []
        self.last_stats = None

    def _handle(self):
        if not self._h:
            h = ctypes.c_void_p()
            C.check(self._L.lvs_pgo_create(_SOLVERS[self._solver_type], self._device, None, ctypes.byref(h)))
            self._h = h
        return self._h

    def close(self):
        if getattr(self, "_h", None):
            self._L.lvs_pgo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_solver(self, solver_type):
        if solver_type not in _SOLVERS:
            raise ValueError("unknown solver type %r" % solver_type)
        self.close()
        self._solver_type = solver_type

    def num_vertices(self):
        return len(self._vertices)

    def num_edges(self):
        return len(self._edges)

    def add_se3_node(self, pose4x4):
        # id = current vertex count (graph_slam.cpp:108); after load() of a file with other ids, the next free one
        vid = max(len(self._vertices) + len(self._planes), getattr(self, "_next_id", 0))
        v = VertexSE3(vid, pose4x4)
        self._next_id = vid + 1
        self._vertices.append(v)
        return v

    def add_se3_edge(self, v1, v2, relative_pose, information_matrix):
        e = EdgeSE3(v1, v2, relative_pose, information_matrix)
        self._edges.append(e)
        return e

    # GPS / IMU priors of the global-graph nodelet (graph_slam.cpp:194-240; global_graph_nodelet.cpp:420-427, 534-541)
    def add_se3_prior_xy_edge(self, v_se3, xy, information_matrix):
        e = EdgeSE3Prior(PRIOR_XY, v_se3, xy, information_matrix)
        self._edges.append(e)
        return e

    def add_se3_prior_xyz_edge(self, v_se3, xyz, information_matrix):
        e = EdgeSE3Prior(PRIOR_XYZ, v_se3, xyz, information_matrix)
        self._edges.append(e)
        return e

    def add_se3_prior_quat_edge(self, v_se3, quat_xyzw, information_matrix):
        e = EdgeSE3Prior(PRIOR_QUAT, v_se3, quat_xyzw, information_matrix)
        self._edges.append(e)
        return e

    def add_se3_prior_vec_edge(self, v_se3, direction, measurement, information_matrix):
        e = EdgeSE3Prior(PRIOR_VEC, v_se3, np.concatenate([np.asarray(direction, float), np.asarray(measurement, float)]), information_matrix)
        self._edges.append(e)
        return e

    # the floor constraint (graph_slam.cpp:116-126, 148-158; global_graph_nodelet.cpp:601-611): ONE plane node, fixed
    def add_plane_node(self, plane_coeffs):
        vid = max(len(self._vertices) + len(self._planes), getattr(self, "_next_id", 0))
        v = VertexPlane(vid, plane_coeffs)
        self._next_id = vid + 1
        self._planes.append(v)
        return v

    def add_se3_plane_edge(self, v_se3, v_plane, plane_coeffs, information_matrix):
        e = EdgeSE3Prior(SE3_PLANE, v_se3, plane_coeffs, information_matrix)
        e.plane = v_plane
        self._edges.append(e)
        return e

    def _floor_plane(self):
        """The plane of the SE3_PLANE edges: they must share one FIXED plane node (a free plane vertex is another vertex type, not on this path)."""
        planes = {id(e.plane): e.plane for e in self._edges if getattr(e, "kind", EDGE_SE3) == SE3_PLANE}
        if not planes:
            return None
        if len(planes) != 1 or not next(iter(planes.values())).fixed():
            raise NotImplementedError("EdgeSE3Plane is supported against ONE FIXED VertexPlane (the floor node of the global-graph nodelet)")
        return next(iter(planes.values())).estimate()

    def add_robust_kernel(self, edge, kernel_type, kernel_size):
        if kernel_type == "NONE":
            return
        if kernel_type != "Huber":
            print("warning : invalid robust kernel type: %s" % kernel_type)      # the reference warns and leaves the edge unkernelled
            return
        edge.kernel = (kernel_type, float(kernel_size))

    def _arrays(self):
        """The flat arrays lvs_pgo_set_graph takes (what shim/graph_slam_b200.cpp packs from the g2o containers), vectorised: the
        per-object version cost 0.4 s for 5 000 vertices / 19 599 edges, more than the whole LM run on the device."""
        nv, ne = len(self._vertices), len(self._edges)
        poses = _pg.pose7_batch(np.stack([v._T for v in self._vertices])) if nv else np.zeros((0, 7))
        fixed = np.fromiter((1 if v._fixed else 0 for v in self._vertices), dtype=np.uint8, count=nv)
        # rows of the pose array are positions in the vertex list, NOT vertex ids: a loaded g2o file may skip ids (plane / GPS nodes)
        row = {v._id: k for k, v in enumerate(self._vertices)}
        ij = np.fromiter((row[x] for e in self._edges for x in (e.vertices[0]._id, e.vertices[1]._id)), dtype=np.int32, count=2 * ne).reshape(ne, 2)
        iu = np.triu_indices(6)
        types = np.fromiter((getattr(e, "kind", EDGE_SE3) for e in self._edges), dtype=np.int32, count=ne)
        if not types.any():
            meas = _pg.pose7_batch(np.stack([e.measurement for e in self._edges])) if ne else np.zeros((0, 7))
            info = np.stack([e.information for e in self._edges])[:, iu[0], iu[1]] if ne else np.zeros((0, 21))
            types = None
        else:
            meas, info6 = np.zeros((ne, 7)), np.zeros((ne, 6, 6))
            binary = np.flatnonzero(types == EDGE_SE3)
            if len(binary):
                meas[binary] = _pg.pose7_batch(np.stack([self._edges[k].measurement for k in binary]))
            for k, e in enumerate(self._edges):
                if types[k] == EDGE_SE3:
                    info6[k] = e.information
                else:
                    meas[k, :len(e.measurement)] = e.measurement
                    d = e.information.shape[0]
                    info6[k, :d, :d] = e.information
            info = info6[:, iu[0], iu[1]]
        hub = np.fromiter((e.kernel[1] if e.kernel else 0.0 for e in self._edges), dtype=np.float64, count=ne)
        return poses, fixed, ij, meas, info, hub, types

    def optimize(self, num_iterations):
        if len(self._edges) < 1:
            return -1                                                             # graph_slam.cpp:302-305
        poses, fixed, ij, meas, info, hub, types = self._arrays()
        st = optimize_arrays(self._L, self._handle(), poses, fixed, ij, meas, info, hub, num_iterations, edge_type=types, floor_plane=self._floor_plane())
        out = np.zeros((len(self._vertices), 7))
        C.check(self._L.lvs_pgo_get_poses(self._handle(), out.ctypes.data))
        for v, T in zip(self._vertices, _pg.matrix_batch(out)):
            v._T = T
        self.last_stats = st
        print("chi2: (before)%g -> (after)%g" % (st["chi2_before"], st["chi2_after"]))
        return st["iterations"]

    # ---- g2o text format (types/slam3d/vertex_se3.cpp:49-64, edge_se3.cpp:44-76)
    def save(self, filename):
        with open(filename, "w") as f:
            for v in self._vertices:
                f.write("VERTEX_SE3:QUAT %d %s\n" % (v._id, " ".join(repr(float(x)) for x in _pg.pose7(v._T))))
                if v._fixed:
                    f.write("FIX %d\n" % v._id)
            for v in self._planes:                                # VertexPlane::write: coefficients and the colour
                f.write("VERTEX_PLANE %d %s 0 0 0\n" % (v._id, " ".join(repr(float(x)) for x in v._c)))
                if v._fixed:
                    f.write("FIX %d\n" % v._id)
            for e in self._edges:
                if getattr(e, "kind", EDGE_SE3) == SE3_PLANE:
                    up = [e.information[r, c] for r in range(3) for c in range(r, 3)]
                    f.write("EDGE_SE3_PLANE %d %d %s %s\n" % (e.vertices[0]._id, e.plane._id, " ".join(repr(float(x)) for x in e.measurement), " ".join(repr(float(x)) for x in up)))
                    continue
                if getattr(e, "kind", EDGE_SE3) != EDGE_SE3:      # write() of the prior edges: measurement, then the upper triangle
                    tag, _, d = _PRIOR_TAGS[e.kind]
                    m = e.measurement if e.kind != PRIOR_QUAT else e.measurement[[3, 0, 1, 2]]        # PriorQuat writes w x y z
                    up = [e.information[r, c] for r in range(d) for c in range(r, d)]
                    f.write("%s %d %s %s\n" % (tag, e.vertices[0]._id, " ".join(repr(float(x)) for x in m), " ".join(repr(float(x)) for x in up)))
                    continue
                up = [e.information[r, c] for r in range(6) for c in range(r, 6)]
                f.write("EDGE_SE3:QUAT %d %d %s %s\n" % (e.vertices[0]._id, e.vertices[1]._id, " ".join(repr(float(x)) for x in _pg.pose7(e.measurement)),
                                                      " ".join(repr(float(x)) for x in up)))
        with open(filename + ".kernels", "w") as f:                               # robust_kernel_io.cpp sidecar
            for e in self._edges:
                if e.kernel and getattr(e, "kind", EDGE_SE3) != EDGE_SE3:
                    f.write("1 %d %s %r\n" % (e.vertices[0]._id, e.kernel[0], e.kernel[1]))
                elif e.kernel:     # "<n vertices> <ids...> <type> <delta>" (g2o/robust_kernel_io.cpp:22-48)
                    f.write("2 %d %d %s %r\n" % (e.vertices[0]._id, e.vertices[1]._id, e.kernel[0], e.kernel[1]))
        return True

    def load(self, filename):
        self._vertices, self._edges, self._planes = [], [], []
        by_id = {}
        with open(filename) as f:
            for line in f:
                t = line.split()
                if not t:
                    continue
                if t[0] == "VERTEX_SE3:QUAT":
                    v = VertexSE3(int(t[1]), _pg.matrix(np.array(t[2:9], dtype=np.float64)))
                    by_id[v._id] = v
                    self._vertices.append(v)
                elif t[0] == "VERTEX_PLANE":
                    v = VertexPlane(int(t[1]), np.array(t[2:6], dtype=np.float64))
                    by_id[v._id] = v
                    self._planes.append(v)
                elif t[0] == "EDGE_SE3_PLANE":
                    up = np.array(t[7:13], dtype=np.float64)
                    info = np.zeros((3, 3))
                    info[np.triu_indices(3)] = up
                    info = info + np.triu(info, 1).T
                    e = EdgeSE3Prior(SE3_PLANE, by_id[int(t[1])], np.array(t[3:7], dtype=np.float64), info)
                    e.plane = by_id[int(t[2])]
                    self._edges.append(e)
                elif t[0] == "FIX":
                    for s in t[1:]:
                        by_id[int(s)].setFixed(True)
                elif t[0] == "EDGE_SE3:QUAT":
                    a, b = int(t[1]), int(t[2])
                    m = np.array(t[3:10], dtype=np.float64)
                    up = np.array(t[10:31], dtype=np.float64)
                    info = np.zeros((6, 6))
                    k = 0
                    for r in range(6):
                        for c in range(r, 6):
                            info[r, c] = info[c, r] = up[k]
                            k += 1
                    self._edges.append(EdgeSE3(by_id[a], by_id[b], _pg.matrix(m), info))
                elif t[0] in _PRIOR_BY_TAG:
                    kind = _PRIOR_BY_TAG[t[0]]
                    _, nm, d = _PRIOR_TAGS[kind]
                    m = np.array(t[2:2 + nm], dtype=np.float64)
                    if kind == PRIOR_QUAT:
                        m = m[[1, 2, 3, 0]]                   # the file holds w x y z
                    up = np.array(t[2 + nm:2 + nm + d * (d + 1) // 2], dtype=np.float64)
                    info = np.zeros((d, d))
                    k = 0
                    for r in range(d):
                        for c in range(r, d):
                            info[r, c] = info[c, r] = up[k]
                            k += 1
                    self._edges.append(EdgeSE3Prior(kind, by_id[int(t[1])], m, info))
        self._vertices.sort(key=lambda v: v._id)
        self._next_id = max((v._id for v in self._vertices + self._planes), default=-1) + 1
        try:
            with open(filename + ".kernels") as f:
                kern = {}                                     # (i, j) -> records in file order: one record is consumed per edge
                for line in f:
                    t = line.split()
                    if len(t) == 5 and t[0] == "2":           # KernelData (robust_kernel_io.cpp:51-62)
                        kern.setdefault((int(t[1]), int(t[2])), []).append((t[3], float(t[4])))
                    elif len(t) == 4 and t[0] == "1":         # a unary edge's record
                        kern.setdefault((int(t[1]), int(t[1]), "unary"), []).append((t[2], float(t[3])))
                    elif len(t) == 4:                         # sidecars written before the vertex count was added
                        kern.setdefault((int(t[0]), int(t[1])), []).append((t[2], float(t[3])))
                for e in self._edges:
                    unary = getattr(e, "kind", EDGE_SE3) != EDGE_SE3
                    q = kern.get((e.vertices[0]._id, e.vertices[1]._id, "unary") if unary else (e.vertices[0]._id, e.vertices[1]._id))
                    if q:
                        e.kernel = q.pop(0)                   # parallel edges between the same vertices take one record each
        except OSError:
            pass
        return True


def save_kitti_poses(filename, poses4x4):
    """One pose per line, the first three rows of the 4x4 matrix row-major with %le, the odometry node's own output format
    (src/lidar_odometry/scan_matching_odom_nodelet.cpp:157-160) and the KITTI ground-truth format it is compared with."""
    with open(filename, "w") as f:
        for T in poses4x4:
            T = np.asarray(T, dtype=np.float64)
            f.write(" ".join("%e" % T[r, c] for r in range(3) for c in range(4)) + "\n")


def load_kitti_poses(filename):
    out = []
    with open(filename) as f:
        for line in f:
            v = line.split()
            if len(v) != 12:
                continue
            T = np.eye(4)
            T[:3, :4] = np.array(v, dtype=np.float64).reshape(3, 4)
            out.append(T)
    return out


_PRIOR_BY_TAG = {v[0]: k for k, v in _PRIOR_TAGS.items()}


def optimize_arrays(L, h, poses7, fixed, ij, meas7, info21, huber, num_iterations, edge_type=None, floor_plane=None):
    """set_graph + optimize on flat arrays; returns the stats dict (used by GraphSLAM.optimize, the tests and the bench)."""
    poses7 = np.ascontiguousarray(poses7, dtype=np.float64)
    ij = np.ascontiguousarray(ij, dtype=np.int32)
    meas7 = np.ascontiguousarray(meas7, dtype=np.float64)
    info21 = np.ascontiguousarray(info21, dtype=np.float64)
    hub = np.ascontiguousarray(huber, dtype=np.float64) if huber is not None else None
    fx = np.ascontiguousarray(fixed, dtype=np.uint8) if fixed is not None else None
    ty = np.ascontiguousarray(edge_type, dtype=np.int32) if edge_type is not None else None
    if floor_plane is not None:
        fp = np.ascontiguousarray(floor_plane, dtype=np.float64)
        C.check(L.lvs_pgo_set_floor_plane(h, fp.ctypes.data))
    C.check(L.lvs_pgo_set_graph_typed(h, poses7.shape[0], poses7.ctypes.data, fx.ctypes.data if fx is not None else None, ij.shape[0], ij.ctypes.data,
                                      meas7.ctypes.data, info21.ctypes.data, hub.ctypes.data if hub is not None else None,
                                      ty.ctypes.data if ty is not None else None))
    st = C.PgoStats()
    rc = L.lvs_pgo_optimize(h, int(num_iterations), ctypes.byref(st))
    if rc != 0 and rc != -10:
        C.check(rc)
    return {k: getattr(st, k) for k, _ in C.PgoStats._fields_}


CHOL_STAT_NAMES = ("nnz_l_blocks", "fronts", "levels", "max_front", "arena_bytes", "factor_fma")


def chol_analyze(n_blocks, off_ij):
    """Host-side symbolic analysis of the direct solver (lvs_pgo_chol_analyze): no device needed.
    off_ij: (m, 2) int32 pairs (row < col) of the non-zero upper blocks.  Returns (stats dict, elimination order)."""
    off = np.ascontiguousarray(off_ij, dtype=np.int32).reshape(-1, 2)
    stats = (ctypes.c_longlong * 6)()
    perm = np.zeros(int(n_blocks), np.int32)
    C.check(C.lib().lvs_pgo_chol_analyze(int(n_blocks), off.shape[0], off.ctypes.data, stats, perm.ctypes.data))
    return dict(zip(CHOL_STAT_NAMES, [int(v) for v in stats])), perm


class PoseGraph:
    """Flat-array handle on the pose-graph C-ABI (tests / bench)."""

    def __init__(self, solver=C.LVS_PGO_LM_CHOL, device=0):
        self._L = C.lib()
        self._h = ctypes.c_void_p()
        C.check(self._L.lvs_pgo_create(solver, device, None, ctypes.byref(self._h)))
        self.nv = self.ne = 0

    def close(self):
        if getattr(self, "_h", None):
            self._L.lvs_pgo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_graph(self, poses7, ij, meas7, info21, huber=None, fixed=None, edge_type=None, floor_plane=None):
        if floor_plane is not None:
            fp = np.ascontiguousarray(floor_plane, dtype=np.float64)
            C.check(self._L.lvs_pgo_set_floor_plane(self._h, fp.ctypes.data))
        self._keep = [np.ascontiguousarray(poses7, dtype=np.float64), np.ascontiguousarray(ij, dtype=np.int32), np.ascontiguousarray(meas7, dtype=np.float64),
                      np.ascontiguousarray(info21, dtype=np.float64), np.ascontiguousarray(huber, dtype=np.float64) if huber is not None else None,
                      np.ascontiguousarray(fixed, dtype=np.uint8) if fixed is not None else None,
                      np.ascontiguousarray(edge_type, dtype=np.int32) if edge_type is not None else None]
        p, e, m, i, hb, fx, ty = self._keep
        self.nv, self.ne = p.shape[0], e.shape[0]
        C.check(self._L.lvs_pgo_set_graph_typed(self._h, self.nv, p.ctypes.data, fx.ctypes.data if fx is not None else None, self.ne, e.ctypes.data,
                                                m.ctypes.data, i.ctypes.data, hb.ctypes.data if hb is not None else None,
                                                ty.ctypes.data if ty is not None else None))

    def set_options(self, pcg_tolerance=0.0, pcg_max_iterations=0):
        C.check(self._L.lvs_pgo_set_solver_options(self._h, float(pcg_tolerance), int(pcg_max_iterations)))

    def optimize(self, max_iterations):
        st = C.PgoStats()
        rc = self._L.lvs_pgo_optimize(self._h, int(max_iterations), ctypes.byref(st))
        if rc != 0 and rc != -10:
            C.check(rc)
        out = {k: getattr(st, k) for k, _ in C.PgoStats._fields_}
        recs = (C.PgoIterRec * 2048)()
        n = ctypes.c_int(0)
        C.check(self._L.lvs_pgo_get_trace(self._h, recs, 2048, ctypes.byref(n)))
        out["trace"] = np.array([[recs[k].chi2, recs[k].lam, recs[k].trials, recs[k].pcg_iterations] for k in range(min(n.value, 2048))]).reshape(-1, 4)
        return out

    def poses(self):
        out = np.zeros((self.nv, 7))
        C.check(self._L.lvs_pgo_get_poses(self._h, out.ctypes.data))
        return out

    def errors(self):
        e, c, tot = np.zeros((self.ne, 6)), np.zeros(self.ne), ctypes.c_double(0)
        C.check(self._L.lvs_pgo_compute_errors(self._h, e.ctypes.data, c.ctypes.data, ctypes.byref(tot)))
        return e, c, tot.value

    def linearize(self):
        nf, no = ctypes.c_int(0), ctypes.c_int(0)
        C.check(self._L.lvs_pgo_system_size(self._h, ctypes.byref(nf), ctypes.byref(no)))
        Hd, Ho, b, off = np.zeros((nf.value, 6, 6)), np.zeros((no.value, 6, 6)), np.zeros(nf.value * 6), np.zeros((no.value, 2), np.int32)
        C.check(self._L.lvs_pgo_linearize(self._h, Hd.ctypes.data, off.ctypes.data, Ho.ctypes.data, b.ctypes.data))
        return dict(Hd=Hd, Ho=Ho, off=off, b=b)

    def chol_info(self):
        stats = (ctypes.c_longlong * 6)()
        C.check(self._L.lvs_pgo_chol_info(self._h, stats))
        return dict(zip(CHOL_STAT_NAMES, [int(v) for v in stats]))

    def solve(self, lam, tolerance=1e-24, max_iterations=0):
        nf = ctypes.c_int(0)
        C.check(self._L.lvs_pgo_system_size(self._h, ctypes.byref(nf), None))
        x, it = np.zeros(nf.value * 6), ctypes.c_int(0)
        C.check(self._L.lvs_pgo_solve(self._h, float(lam), float(tolerance), int(max_iterations), x.ctypes.data, ctypes.byref(it)))
        return x, it.value
