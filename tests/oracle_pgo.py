"""ctypes binding of oracle/libpgo_oracle.so — the CHECKER for the pose-graph path.  Test infrastructure only."""
import ctypes
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ODIR = os.path.join(_ROOT, "oracle")
_SO = os.path.join(_ODIR, "libpgo_oracle.so")
_CS = os.path.join(_ODIR, "_ref", "libcsparse_ref.so")

SOLVER_CSPARSE, SOLVER_PCG, SOLVER_DENSE = 0, 1, 2
ALG_LM, ALG_GN = 0, 1
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.check_call(["make", "-s", "-C", _ODIR, "all"])
        if not os.path.exists(_CS) and os.path.isdir("/root/reference"):
            subprocess.call(["make", "-s", "-C", _ODIR, "ref"])
        L = ctypes.CDLL(_SO)
        vp, i32, f64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
        L.opgo_compute_dq_dR.restype = None; L.opgo_compute_dq_dR.argtypes = [vp, vp]
        L.opgo_oplus_matrix.restype = None; L.opgo_oplus_matrix.argtypes = [vp, vp, vp, vp]
        L.opgo_load_csparse.restype = i32; L.opgo_load_csparse.argtypes = [ctypes.c_char_p]
        L.opgo_have_csparse.restype = i32
        L.opgo_create.restype = vp
        L.opgo_destroy.argtypes = [vp]; L.opgo_destroy.restype = None
        L.opgo_set_graph.argtypes = [vp, i32, vp, vp, i32, vp, vp, vp, vp]; L.opgo_set_graph.restype = None
        L.opgo_set_graph_typed.argtypes = [vp, i32, vp, vp, i32, vp, vp, vp, vp, vp]; L.opgo_set_graph_typed.restype = None
        L.opgo_prior_error.argtypes = [i32, vp, vp, vp]; L.opgo_prior_error.restype = None
        L.opgo_set_floor_plane.argtypes = [vp, vp]; L.opgo_set_floor_plane.restype = None
        L.opgo_prior_jacobian.argtypes = [i32, vp, vp, vp]; L.opgo_prior_jacobian.restype = None
        L.opgo_get_poses.argtypes = [vp, vp]; L.opgo_get_poses.restype = None
        L.opgo_compute_errors.argtypes = [vp, vp, vp]; L.opgo_compute_errors.restype = f64
        L.opgo_linearize.argtypes = [vp, vp, vp]; L.opgo_linearize.restype = None
        L.opgo_num_free.argtypes = [vp]; L.opgo_num_free.restype = i32
        L.opgo_num_offdiag.argtypes = [vp]; L.opgo_num_offdiag.restype = i32
        L.opgo_get_system.argtypes = [vp, vp, vp, vp, vp]; L.opgo_get_system.restype = None
        L.opgo_solve.argtypes = [vp, f64, i32, f64, i32, vp]; L.opgo_solve.restype = i32
        L.opgo_last_pcg_iters.argtypes = [vp]; L.opgo_last_pcg_iters.restype = i32
        L.opgo_optimize.argtypes = [vp, i32, i32, i32, f64, i32, vp]; L.opgo_optimize.restype = i32
        L.opgo_trace_len.argtypes = [vp]; L.opgo_trace_len.restype = i32
        L.opgo_get_trace.argtypes = [vp, vp]; L.opgo_get_trace.restype = None
        L.opgo_edge_error.argtypes = [vp, vp, vp, vp]; L.opgo_edge_error.restype = None
        L.opgo_edge_jacobians.argtypes = [vp, vp, vp, vp, vp]; L.opgo_edge_jacobians.restype = None
        L.opgo_oplus.argtypes = [vp, vp, vp]; L.opgo_oplus.restype = None
        if os.path.exists(_CS):
            L.opgo_load_csparse(_CS.encode())
        _lib = L
    return _lib


def csparse_lnz(o):
    """nnz(L) of CSparse's symbolic factorisation (its own AMD ordering) after a SOLVER_CSPARSE solve on OraclePGO `o`; -1 if none."""
    L = lib()
    L.opgo_csparse_lnz.restype = ctypes.c_double
    L.opgo_csparse_lnz.argtypes = [ctypes.c_void_p]
    return float(L.opgo_csparse_lnz(o.h))


def have_csparse():
    return bool(lib().opgo_have_csparse())


def _c(a, dt=np.float64):
    return np.ascontiguousarray(a, dtype=dt)


class OraclePGO:
    def __init__(self):
        self.L = lib()
        self.h = self.L.opgo_create()

    def __del__(self):
        try:
            self.L.opgo_destroy(self.h)
        except Exception:
            pass

    def set_graph(self, poses7, edges_ij, meas7, info21, huber=None, fixed=None, edge_type=None, floor_plane=None):
        if floor_plane is not None:
            self.L.opgo_set_floor_plane(self.h, _c(floor_plane).ctypes.data)
        p, ij, m, inf = _c(poses7), _c(edges_ij, np.int32), _c(meas7), _c(info21)
        self.nv, self.ne = p.shape[0], ij.shape[0]
        hub = _c(huber) if huber is not None else None
        fx = _c(fixed, np.uint8) if fixed is not None else None
        ty = _c(edge_type, np.int32) if edge_type is not None else None
        self.L.opgo_set_graph_typed(self.h, self.nv, p.ctypes.data, fx.ctypes.data if fx is not None else None, self.ne, ij.ctypes.data, m.ctypes.data,
                                    inf.ctypes.data, hub.ctypes.data if hub is not None else None, ty.ctypes.data if ty is not None else None)

    def poses(self):
        out = np.zeros((self.nv, 7))
        self.L.opgo_get_poses(self.h, out.ctypes.data)
        return out

    def errors(self):
        e, c = np.zeros((self.ne, 6)), np.zeros(self.ne)
        tot = self.L.opgo_compute_errors(self.h, e.ctypes.data, c.ctypes.data)
        return e, c, tot

    def linearize(self):
        Ji, Jj = np.zeros((self.ne, 6, 6)), np.zeros((self.ne, 6, 6))
        self.L.opgo_linearize(self.h, Ji.ctypes.data, Jj.ctypes.data)
        nf, no = self.L.opgo_num_free(self.h), self.L.opgo_num_offdiag(self.h)
        Hd, Ho, b, off = np.zeros((nf, 6, 6)), np.zeros((no, 6, 6)), np.zeros(nf * 6), np.zeros((no, 2), np.int32)
        self.L.opgo_get_system(self.h, Hd.ctypes.data, off.ctypes.data, Ho.ctypes.data, b.ctypes.data)
        return dict(Ji=Ji, Jj=Jj, Hd=Hd, Ho=Ho, off=off, b=b)

    def solve(self, lam, solver=SOLVER_CSPARSE, pcg_tol=-1.0, pcg_max_iter=-1):
        nf = self.L.opgo_num_free(self.h)
        x = np.zeros(nf * 6)
        ok = self.L.opgo_solve(self.h, float(lam), solver, float(pcg_tol), pcg_max_iter, x.ctypes.data)
        return bool(ok), x, self.L.opgo_last_pcg_iters(self.h)

    def optimize(self, max_iters, algorithm=ALG_LM, solver=SOLVER_CSPARSE, pcg_tol=-1.0, pcg_max_iter=-1):
        st = np.zeros(5)
        it = self.L.opgo_optimize(self.h, max_iters, algorithm, solver, float(pcg_tol), pcg_max_iter, st.ctypes.data)
        n = self.L.opgo_trace_len(self.h)
        tr = np.zeros((n, 4))
        if n:
            self.L.opgo_get_trace(self.h, tr.ctypes.data)
        return dict(iterations=it, chi2_before=st[0], chi2_after=st[1], lam=st[2], trials=int(st[3]), robust_chi2_after=st[4], trace=tr)


def dense_system(lin, lam=0.0):
    nf = lin["Hd"].shape[0]
    M = np.zeros((nf * 6, nf * 6))
    for v in range(nf):
        M[v * 6:v * 6 + 6, v * 6:v * 6 + 6] = lin["Hd"][v]
    for (i, j), blk in zip(lin["off"], lin["Ho"]):
        M[i * 6:i * 6 + 6, j * 6:j * 6 + 6] = blk
        M[j * 6:j * 6 + 6, i * 6:i * 6 + 6] = blk.T
    return M + lam * np.eye(nf * 6)


def edge_error(z7, xi7, xj7):
    e = np.zeros(6)
    lib().opgo_edge_error(_c(z7).ctypes.data, _c(xi7).ctypes.data, _c(xj7).ctypes.data, e.ctypes.data)
    return e


def edge_jacobians(z7, xi7, xj7):
    Ji, Jj = np.zeros((6, 6)), np.zeros((6, 6))
    lib().opgo_edge_jacobians(_c(z7).ctypes.data, _c(xi7).ctypes.data, _c(xj7).ctypes.data, Ji.ctypes.data, Jj.ctypes.data)
    return Ji, Jj


def prior_error(kind, meas, x7):
    """computeError of EdgeSE3PriorXY / XYZ / Quat / Vec (kind 1-4), zero-padded to 6."""
    e, m = np.zeros(6), np.zeros(8)            # kind 5: the measured plane (4) followed by the fixed plane (4)
    m[:len(meas)] = meas
    lib().opgo_prior_error(int(kind), m.ctypes.data, _c(x7).ctypes.data, e.ctypes.data)
    return e


def prior_jacobian(kind, meas, x7):
    J, m = np.zeros((6, 6)), np.zeros(8)
    m[:len(meas)] = meas
    lib().opgo_prior_jacobian(int(kind), m.ctypes.data, _c(x7).ctypes.data, J.ctypes.data)
    return J


def oplus(x7, d6):
    out = np.zeros(7)
    lib().opgo_oplus(_c(x7).ctypes.data, _c(d6).ctypes.data, out.ctypes.data)
    return out
