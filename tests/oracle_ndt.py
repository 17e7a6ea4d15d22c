"""ctypes binding of oracle/libndt_oracle.so — the CHECKER.  Test infrastructure only.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by
the product package (lv_slam_b200/).
"""
import ctypes
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ODIR = os.path.join(_ROOT, "oracle")
_SO = os.path.join(_ODIR, "libndt_oracle.so")

KDTREE, DIRECT26, DIRECT7, DIRECT1 = 0, 1, 2, 3
VAR_OMP, VAR_PCA, VAR_GROUND = 0, 1, 2

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _ODIR, "all"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        vp, i32, f32, f64, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_size_t
        L.ondt_create.restype = vp; L.ondt_create.argtypes = [i32]
        L.ondt_destroy.restype = None; L.ondt_destroy.argtypes = [vp]
        L.ondt_set_params.restype = None; L.ondt_set_params.argtypes = [vp, f32, f64, f64, f64, i32, i32, i32]
        L.ondt_set_target.restype = None; L.ondt_set_target.argtypes = [vp, vp, sz, sz]
        L.ondt_set_source.restype = None; L.ondt_set_source.argtypes = [vp, vp, sz, sz]
        L.ondt_get_grid.restype = None; L.ondt_get_grid.argtypes = [vp, vp, vp, vp]
        L.ondt_get_gauss.restype = None; L.ondt_get_gauss.argtypes = [vp, vp]
        L.ondt_num_leaves.restype = i32; L.ondt_num_leaves.argtypes = [vp]
        L.ondt_get_leaves.restype = None; L.ondt_get_leaves.argtypes = [vp] * 12
        L.ondt_get_leaf_angles.restype = None; L.ondt_get_leaf_angles.argtypes = [vp, vp]
        L.ondt_get_leaf_evecs.restype = None; L.ondt_get_leaf_evecs.argtypes = [vp, vp]
        L.ondt_neighbours.restype = i32; L.ondt_neighbours.argtypes = [vp, vp, i32, vp]
        L.ondt_lookup_keys.restype = None; L.ondt_lookup_keys.argtypes = [vp, vp, sz, sz, vp]
        L.ondt_transform.restype = None; L.ondt_transform.argtypes = [vp, sz, sz, vp, vp]
        L.ondt_eval_derivatives.restype = f64; L.ondt_eval_derivatives.argtypes = [vp, vp, vp, i32, vp, vp]
        L.ondt_eval_hessian.restype = None; L.ondt_eval_hessian.argtypes = [vp, vp, vp, vp]
        L.ondt_calculate_score.restype = f64; L.ondt_calculate_score.argtypes = [vp, vp]
        L.otransform_double.restype = None; L.otransform_double.argtypes = [vp, sz, sz, i32, vp, vp]
        L.oprefilter.restype = sz; L.oprefilter.argtypes = [vp, sz, sz, i32, f64, f64, i32, f32, vp, vp]
        L.ondt_fitness_score.restype = f64; L.ondt_fitness_score.argtypes = [vp, vp, f64, vp]
        L.ondt_align.restype = i32; L.ondt_align.argtypes = [vp, vp, vp, vp, vp]
        L.ondt_trace_len.restype = i32; L.ondt_trace_len.argtypes = [vp]
        L.ondt_get_trace.restype = None; L.ondt_get_trace.argtypes = [vp, vp]
        L.ose3_exp_matrix4f.restype = None; L.ose3_exp_matrix4f.argtypes = [vp, vp]
        L.ose3_exp.restype = None; L.ose3_exp.argtypes = [vp, vp, vp]
        L.ose3_log_from_matrix4f.restype = None; L.ose3_log_from_matrix4f.argtypes = [vp, vp]
        L.ose3_compose_log.restype = None; L.ose3_compose_log.argtypes = [vp, vp, vp]
        L.oracle_expf_restated.restype = ctypes.c_longlong; L.oracle_expf_restated.argtypes = [vp, vp, ctypes.c_longlong]
        L.olin_svd6_solve.restype = None; L.olin_svd6_solve.argtypes = [vp, vp, vp, vp]
        L.olin_sym3_eig.restype = None; L.olin_sym3_eig.argtypes = [vp, vp, vp]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _colmajor16(T):
    """4x4 (numpy row-major view) -> Eigen::Matrix4f memory order."""
    return np.ascontiguousarray(np.asarray(T, dtype=np.float32).T.reshape(16))


def _from_colmajor16(v):
    return np.asarray(v, dtype=np.float32).reshape(4, 4).T.copy()


class OracleNDT:
    """Mirror of the reference registration object's API, on the CPU restatement."""

    def __init__(self, variant=VAR_OMP, resolution=1.0, step_size=0.1, outlier_ratio=0.55, trans_eps=0.1, max_iter=35,
                 search=DIRECT7, num_threads=1):
        self.L = lib()
        self.h = self.L.ondt_create(variant)
        self.variant = variant
        self.params = dict(resolution=resolution, step_size=step_size, outlier_ratio=outlier_ratio, trans_eps=trans_eps,
                           max_iter=max_iter, search=search, num_threads=num_threads)
        self._push()

    def _push(self):
        p = self.params
        self.L.ondt_set_params(self.h, p["resolution"], p["step_size"], p["outlier_ratio"], p["trans_eps"], p["max_iter"],
                               p["search"], p["num_threads"])

    def set(self, **kw):
        self.params.update(kw)
        self._push()

    def __del__(self):
        try:
            self.L.ondt_destroy(self.h)
        except Exception:
            pass

    def set_target(self, xyz):
        a = _f32(xyz)
        self.L.ondt_set_target(self.h, a.ctypes.data, a.shape[0], a.shape[1])

    def set_source(self, xyz):
        a = _f32(xyz)
        self.n_src = a.shape[0]
        self.L.ondt_set_source(self.h, a.ctypes.data, a.shape[0], a.shape[1])

    def grid(self):
        mn, mx, dv = (np.zeros(3, np.int32) for _ in range(3))
        self.L.ondt_get_grid(self.h, mn.ctypes.data, mx.ctypes.data, dv.ctypes.data)
        return mn, mx, dv

    def gauss(self):
        d = np.zeros(3)
        self.L.ondt_get_gauss(self.h, d.ctypes.data)
        return d

    def leaves(self):
        n = self.L.ondt_num_leaves(self.h)
        out = dict(keys=np.zeros(n, np.int32), nr_points=np.zeros(n, np.int32), raw_points=np.zeros(n, np.int32),
                   mean=np.zeros((n, 3)), cov=np.zeros((n, 3, 3)), icov=np.zeros((n, 3, 3)), evals=np.zeros((n, 3)),
                   centroid=np.zeros((n, 3), np.float32), weight=np.zeros(n, np.int32), label=np.zeros(n, np.int32),
                   in_cloud=np.zeros(n, np.int32))
        order = ["keys", "nr_points", "raw_points", "mean", "cov", "icov", "evals", "centroid", "weight", "label", "in_cloud"]
        self.L.ondt_get_leaves(self.h, *[out[k].ctypes.data for k in order])
        return out

    def leaf_angles(self):
        """pclomp_ground: angle of every occupied cell's normal to the z axis [deg], -1 without an eigen-decomposition."""
        out = np.zeros(self.L.ondt_num_leaves(self.h))
        self.L.ondt_get_leaf_angles(self.h, out.ctypes.data)
        return out

    def leaf_evecs(self):
        """Eigenvectors of every occupied cell [n, 3, 3] (columns in eigenvalue order); identity below min_points_per_voxel."""
        out = np.zeros((self.L.ondt_num_leaves(self.h), 3, 3))
        self.L.ondt_get_leaf_evecs(self.h, out.ctypes.data)
        return out

    def neighbours(self, point, mode):
        """keys of the cells the direct search `mode` returns for one point, in push order"""
        p = np.ascontiguousarray(point[:3], dtype=np.float32)
        out = np.zeros(26, np.int32)
        n = self.L.ondt_neighbours(self.h, p.ctypes.data, int(mode), out.ctypes.data)
        return out[:n].copy()

    def lookup_keys(self, xyz):
        a = _f32(xyz)
        keys = np.zeros(a.shape[0], np.int32)
        self.L.ondt_lookup_keys(self.h, a.ctypes.data, a.shape[0], a.shape[1], keys.ctypes.data)
        return keys

    def eval_derivatives(self, p6, T=None, compute_hessian=True):
        p = np.ascontiguousarray(p6, dtype=np.float64)
        g, H = np.zeros(6), np.zeros((6, 6))
        Tm = _colmajor16(T) if T is not None else None
        s = self.L.ondt_eval_derivatives(self.h, p.ctypes.data, Tm.ctypes.data if Tm is not None else None,
                                         int(compute_hessian), g.ctypes.data, H.ctypes.data)
        return s, g, H

    def eval_hessian(self, p6, T=None):
        p = np.ascontiguousarray(p6, dtype=np.float64)
        H = np.zeros((6, 6))
        Tm = _colmajor16(T) if T is not None else None
        self.L.ondt_eval_hessian(self.h, p.ctypes.data, Tm.ctypes.data if Tm is not None else None, H.ctypes.data)
        return H

    def calculate_score(self, T):
        Tm = _colmajor16(T)
        return self.L.ondt_calculate_score(self.h, Tm.ctypes.data)

    def fitness_score(self, T, max_range=np.finfo(np.float64).max):
        """getFitnessScore(max_range) of T * source against the target points -> (score, correspondences)"""
        Tm = _colmajor16(T)
        n = ctypes.c_int(0)
        return self.L.ondt_fitness_score(self.h, Tm.ctypes.data, float(max_range), ctypes.byref(n)), n.value

    def align(self, guess=None, want_cloud=False):
        g = _colmajor16(np.eye(4) if guess is None else guess)
        fin = np.zeros(16, np.float32)
        stats = np.zeros(4)
        cloud = np.zeros((self.n_src, 3), np.float32) if want_cloud else None
        it = self.L.ondt_align(self.h, g.ctypes.data, fin.ctypes.data, stats.ctypes.data,
                               cloud.ctypes.data if want_cloud else None)
        res = dict(final=_from_colmajor16(fin), iterations=it, converged=bool(stats[0]), trans_probability=stats[1],
                   n_eval=int(stats[2]), n_hess=int(stats[3]), trace=self.trace())
        if want_cloud:
            res["cloud"] = cloud
        return res

    def trace(self):
        n = self.L.ondt_trace_len(self.h)
        t = np.zeros((n, 22))
        if n:
            self.L.ondt_get_trace(self.h, t.ctypes.data)
        return t


def prefilter(cloud, near=0.5, far=100.0, use_filter=True, leaf=0.1):
    """distance_filter + pcl::VoxelGrid restatement -> (points [m, n_fields], overflow flag)"""
    a = _f32(cloud)
    nf = 4 if a.shape[1] >= 4 else 3
    out = np.zeros((a.shape[0], nf), np.float32)
    fl = ctypes.c_int(0)
    m = lib().oprefilter(a.ctypes.data, a.shape[0], a.shape[1], nf, float(near), float(far), int(bool(use_filter)), float(leaf), out.ctypes.data, ctypes.byref(fl))
    return out[:m].copy(), fl.value


def window_map(clouds, transforms, leaf=0.1):
    """w_cloud = clouds[0] + sum transformPointCloud(clouds[k], transforms[k]) (double matrices), then VoxelGrid(leaf)
    (global_graph_nodelet.cpp:199-243)."""
    parts = []
    for c, T in zip(clouds, transforms):
        a = _f32(c)
        nf = 4 if a.shape[1] >= 4 else 3
        if T is None:
            parts.append(a[:, :nf].copy())
            continue
        out = np.zeros((a.shape[0], nf), np.float32)
        Tm = np.ascontiguousarray(np.asarray(T, dtype=np.float64).T.reshape(16))
        lib().otransform_double(a.ctypes.data, a.shape[0], a.shape[1], nf, Tm.ctypes.data, out.ctypes.data)
        parts.append(out)
    return prefilter(np.concatenate(parts), use_filter=False, leaf=leaf)[0]


def transform(xyz, T):
    a = _f32(xyz)
    out = np.zeros((a.shape[0], 3), np.float32)
    Tm = _colmajor16(T)
    lib().ondt_transform(a.ctypes.data, a.shape[0], a.shape[1], Tm.ctypes.data, out.ctypes.data)
    return out


def se3_exp_matrix4f(p6):
    p = np.ascontiguousarray(p6, dtype=np.float64)
    M = np.zeros(16, np.float32)
    lib().ose3_exp_matrix4f(p.ctypes.data, M.ctypes.data)
    return _from_colmajor16(M)


def se3_exp(p6):
    p = np.ascontiguousarray(p6, dtype=np.float64)
    q, t = np.zeros(4), np.zeros(3)
    lib().ose3_exp(p.ctypes.data, q.ctypes.data, t.ctypes.data)
    return q, t


def se3_log_from_matrix4f(T):
    M = _colmajor16(T)
    p = np.zeros(6)
    lib().ose3_log_from_matrix4f(M.ctypes.data, p.ctypes.data)
    return p


def se3_compose_log(delta6, p6):
    d = np.ascontiguousarray(delta6, dtype=np.float64)
    p = np.ascontiguousarray(p6, dtype=np.float64)
    o = np.zeros(6)
    lib().ose3_compose_log(d.ctypes.data, p.ctypes.data, o.ctypes.data)
    return o


_sref = False


def sophus_ref():
    """oracle/_ref/libsophus_ref.so: the reference's OWN Sophus sources (so3.cpp, se3.cpp from 3rdtools/Sophus-a621ff2-ubuntu18.04.zip) compiled by
    oracle/build_ref.sh against the Eigen stand-in of oracle/ref_stubs/.  None where neither the library nor the reference tree is present."""
    global _sref
    if _sref is False:
        so = os.path.join(_ODIR, "_ref", "libsophus_ref.so")
        if not os.path.exists(so) and os.path.exists("/root/reference/3rdtools/Sophus-a621ff2-ubuntu18.04.zip"):
            subprocess.call(["sh", os.path.join(_ODIR, "build_ref.sh")])
        if os.path.exists(so):
            L = ctypes.CDLL(so)
            vp = ctypes.c_void_p
            for name, n in (("sref_se3_exp", 3), ("sref_se3_exp_matrix", 2), ("sref_se3_log_of_Rt", 3), ("sref_compose_log", 3), ("sref_log_exp", 2),
                            ("sref_inverse", 3), ("sref_transform", 3)):
                getattr(L, name).restype = None
                getattr(L, name).argtypes = [vp] * n
            L.sref_selftest.restype = ctypes.c_double
            L.sref_selftest.argtypes = []
            _sref = L
        else:
            _sref = None
    return _sref


_nref = {}
_NREF_SO = {VAR_OMP: "libndt_ref.so", VAR_PCA: "libndt_pca_ref.so", VAR_GROUND: "libndt_ground_ref.so"}


def ndt_ref_lib(variant=VAR_OMP):
    """oracle/_ref/libndt{,_pca,_ground}_ref.so: the reference's OWN member functions of pclomp:: / pclpca::NormalDistributionsTransform /
    pclomp_ground::NormalDistributionsTransformGround (taken from ndt_omp_impl2.hpp / ndt_pca_impl2.hpp / ndt_ground_impl.hpp at build time)
    compiled in oracle/ndt_ref_harness.cpp.  None where neither the library nor the reference tree is present."""
    if variant not in _nref:
        so = os.path.join(_ODIR, "_ref", _NREF_SO[variant])
        if not os.path.exists(so) and os.path.exists("/root/reference/include/ndt_omp/ndt_omp_impl2.hpp"):
            subprocess.call(["sh", os.path.join(_ODIR, "build_ref.sh")])
        if os.path.exists(so):
            L = ctypes.CDLL(so)
            vp, i32, f32, f64, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_size_t
            L.nref_create.restype = vp; L.nref_create.argtypes = []
            L.nref_destroy.restype = None; L.nref_destroy.argtypes = [vp]
            L.nref_set_params.restype = None; L.nref_set_params.argtypes = [vp, f32, f64, f64, f64, i32, i32]
            L.nref_set_target_cells.restype = None; L.nref_set_target_cells.argtypes = [vp, i32] + [vp] * 9 + [f32, i32] + [vp] * 3
            L.nref_set_source.restype = None; L.nref_set_source.argtypes = [vp, vp, sz, sz]
            L.nref_eval_derivatives.restype = f64; L.nref_eval_derivatives.argtypes = [vp, vp, vp, i32, vp, vp]
            L.nref_eval_hessian.restype = None; L.nref_eval_hessian.argtypes = [vp, vp, vp, vp]
            L.nref_calculate_score.restype = f64; L.nref_calculate_score.argtypes = [vp, vp]
            L.nref_align.restype = i32; L.nref_align.argtypes = [vp, vp, vp, vp, vp, vp]
            _nref[variant] = L
        else:
            _nref[variant] = None
    return _nref[variant]


class ReferenceNDT:
    """The reference's own computeTransformation / computeDerivatives / ... (oracle/ndt_ref_harness.cpp) on the voxel cells of an OracleNDT."""

    def __init__(self, oracle):
        self.L = ndt_ref_lib(oracle.variant)
        self.h = self.L.nref_create()
        p = oracle.params
        self.L.nref_set_params(self.h, p["resolution"], p["step_size"], p["outlier_ratio"], p["trans_eps"], p["max_iter"], p["search"])
        lv = oracle.leaves()
        mn, mx, dv = oracle.grid()
        self._keep = [np.ascontiguousarray(lv[k]) for k in ("keys", "nr_points", "mean", "icov", "centroid", "in_cloud")] + [mn, mx, dv]
        extra = [np.ascontiguousarray(lv["weight"]), np.ascontiguousarray(oracle.leaf_evecs()), np.ascontiguousarray(lv["evals"])]
        self.L.nref_set_target_cells(self.h, len(lv["keys"]), *[a.ctypes.data for a in self._keep], p["resolution"], 6, *[a.ctypes.data for a in extra])
        self._keep += extra

    def __del__(self):
        try:
            self.L.nref_destroy(self.h)
        except Exception:
            pass

    def set_source(self, xyz):
        a = _f32(xyz)
        self.n_src = a.shape[0]
        self.L.nref_set_source(self.h, a.ctypes.data, a.shape[0], a.shape[1])

    def eval_derivatives(self, p6, T=None, compute_hessian=True):
        p = np.ascontiguousarray(p6, dtype=np.float64)
        g, H = np.zeros(6), np.zeros((6, 6))
        M = _colmajor16(T) if T is not None else None
        s = self.L.nref_eval_derivatives(self.h, p.ctypes.data, M.ctypes.data if M is not None else None, int(compute_hessian), g.ctypes.data, H.ctypes.data)
        return s, g, H

    def eval_hessian(self, p6, T=None):
        p = np.ascontiguousarray(p6, dtype=np.float64)
        H = np.zeros((6, 6))
        M = _colmajor16(T) if T is not None else None
        self.L.nref_eval_hessian(self.h, p.ctypes.data, M.ctypes.data if M is not None else None, H.ctypes.data)
        return H

    def calculate_score(self, T):
        M = _colmajor16(T)
        return self.L.nref_calculate_score(self.h, M.ctypes.data)

    def align(self, guess):
        g = _colmajor16(guess)
        fin = np.zeros(16, np.float32)
        it, tp = ctypes.c_int(0), ctypes.c_double(0)
        cloud = np.zeros((self.n_src, 3), np.float32)
        conv = self.L.nref_align(self.h, g.ctypes.data, fin.ctypes.data, ctypes.byref(it), ctypes.byref(tp), cloud.ctypes.data)
        return dict(final=_from_colmajor16(fin), iterations=it.value, converged=bool(conv), trans_probability=tp.value, cloud=cloud)


_vref = {}


class ReferenceVoxelGrid:
    """The reference's own VoxelGridCovariance::applyFilter and getNeighborhoodAtPoint{,7,1} (oracle/voxel_ref_harness.cpp); pca=True: those of
    pclpca::VoxelGridCovariance (with the per-leaf dimension label and weight)."""

    @staticmethod
    def available(pca=False):
        if pca not in _vref:
            so = os.path.join(_ODIR, "_ref", "libvoxel_pca_ref.so" if pca else "libvoxel_ref.so")
            if not os.path.exists(so) and os.path.exists("/root/reference/include/ndt_omp/voxel_grid_covariance_omp_impl.hpp"):
                subprocess.call(["sh", os.path.join(_ODIR, "build_ref.sh")])
            if os.path.exists(so):
                L = ctypes.CDLL(so)
                vp = ctypes.c_void_p
                L.vref_create.restype = vp; L.vref_create.argtypes = []
                L.vref_destroy.restype = None; L.vref_destroy.argtypes = [vp]
                L.vref_build.restype = ctypes.c_int; L.vref_build.argtypes = [vp, vp, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_float, ctypes.c_int, ctypes.c_double]
                L.vref_get_grid.restype = None; L.vref_get_grid.argtypes = [vp] * 4
                L.vref_get_leaves.restype = None; L.vref_get_leaves.argtypes = [vp] * 9
                L.vref_get_pca.restype = None; L.vref_get_pca.argtypes = [vp] * 3
                L.vref_neighbours.restype = ctypes.c_int; L.vref_neighbours.argtypes = [vp, vp, ctypes.c_int, vp]
                _vref[pca] = L
            else:
                _vref[pca] = None
        return _vref[pca] is not None

    def __init__(self, xyz, leaf=1.0, min_points=6, eig_mult=0.01, pca=False):
        assert self.available(pca)
        self.L = _vref[pca]
        self.h = self.L.vref_create()
        a = _f32(xyz)
        self.n = self.L.vref_build(self.h, a.ctypes.data, a.shape[0], a.shape[1], leaf, min_points, eig_mult)

    def __del__(self):
        try:
            self.L.vref_destroy(self.h)
        except Exception:
            pass

    def grid(self):
        mn, mx, dv = (np.zeros(3, np.int32) for _ in range(3))
        self.L.vref_get_grid(self.h, mn.ctypes.data, mx.ctypes.data, dv.ctypes.data)
        return mn, mx, dv

    def leaves(self):
        n = self.n
        out = dict(keys=np.zeros(n, np.int32), nr_points=np.zeros(n, np.int32), mean=np.zeros((n, 3)), cov=np.zeros((n, 3, 3)), icov=np.zeros((n, 3, 3)),
                   evecs=np.zeros((n, 3, 3)), evals=np.zeros((n, 3)), centroid=np.zeros((n, 3), np.float32))
        self.L.vref_get_leaves(self.h, *[out[k].ctypes.data for k in ("keys", "nr_points", "mean", "cov", "icov", "evecs", "evals", "centroid")])
        return out

    def neighbours(self, point, mode):
        p = np.ascontiguousarray(point[:3], dtype=np.float32)
        out = np.zeros(26, np.int32)
        n = self.L.vref_neighbours(self.h, p.ctypes.data, int(mode), out.ctypes.data)
        return out[:n].copy()

    def pca(self):
        """pclpca: (dimension label, int getDimension2d()) per cell"""
        label, weight = np.zeros(self.n, np.int32), np.zeros(self.n, np.int32)
        self.L.vref_get_pca(self.h, label.ctypes.data, weight.ctypes.data)
        return label, weight


_iref = False


def info_ref():
    """oracle/_ref/libinfo_ref.so: the reference's own information_matrix_calculator.cpp compiled against stand-in headers; None if unavailable."""
    global _iref
    if _iref is False:
        so = os.path.join(_ODIR, "_ref", "libinfo_ref.so")
        if not os.path.exists(so) and os.path.exists("/root/reference/src/global_graph/information_matrix_calculator.cpp"):
            subprocess.call(["sh", os.path.join(_ODIR, "build_ref.sh")])
        if os.path.exists(so):
            L = ctypes.CDLL(so)
            vp, sz, f64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_double
            L.iref_fitness_score.restype = f64; L.iref_fitness_score.argtypes = [vp, sz, sz, vp, sz, sz, vp, f64]
            L.iref_information_matrix.restype = None; L.iref_information_matrix.argtypes = [vp, sz, sz, vp, sz, sz, vp, ctypes.c_int, vp, vp, vp]
            _iref = L
        else:
            _iref = None
    return _iref


def ref_fitness_score(cloud1, cloud2, relpose, max_range=np.finfo(np.float64).max):
    a, b = _f32(cloud1), _f32(cloud2)
    T = np.ascontiguousarray(relpose, dtype=np.float64)
    return info_ref().iref_fitness_score(a.ctypes.data, a.shape[0], a.shape[1], b.ctypes.data, b.shape[0], b.shape[1], T.ctypes.data, float(max_range))


def ref_information_matrix(cloud1, cloud2, relpose, **params):
    a, b = _f32(cloud1), _f32(cloud2)
    T = np.ascontiguousarray(relpose, dtype=np.float64)
    names = (ctypes.c_char_p * max(len(params), 1))(*[k.encode() for k in params])
    vals = np.array([float(v) for v in params.values()] or [0.0])
    out = np.zeros((6, 6))
    info_ref().iref_information_matrix(a.ctypes.data, a.shape[0], a.shape[1], b.ctypes.data, b.shape[0], b.shape[1], T.ctypes.data, len(params), names,
                                        vals.ctypes.data, out.ctypes.data)
    return out


def svd6_solve(A, b):
    A = np.ascontiguousarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x, sv = np.zeros(6), np.zeros(6)
    lib().olin_svd6_solve(A.ctypes.data, b.ctypes.data, x.ctypes.data, sv.ctypes.data)
    return x, sv


def sym3_eig(A):
    A = np.ascontiguousarray(A, dtype=np.float64)
    ev, V = np.zeros(3), np.zeros((3, 3))
    lib().olin_sym3_eig(A.ctypes.data, ev.ctypes.data, V.ctypes.data)
    return ev, V


def expf_restated(x):
    """(restated glibc expf of x, number of inputs where it differs from the host libm's expf bit for bit)."""
    x = _f32(x)
    out = np.empty_like(x)
    bad = lib().oracle_expf_restated(x.ctypes.data, out.ctypes.data, x.size)
    return out, int(bad)
