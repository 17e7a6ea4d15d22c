"""CPU checks that pin the NDT oracle itself (no GPU): closed-form known answers, property checks and finite differences.

The reference ships no golden vector for this path, so these are the pins: the gauss constants the
reference's formulas give for its own parameters, se(3) round trips to 1e-10 (Sophus' own test tolerance), the small dense
solvers against numpy, a hand-built voxel, and finite differences of the score against the gradient / of the gradient against the
Hessian at poses where the reference's point Jacobian is the true derivative (rotation vector = 0).
"""
import json
import os

import numpy as np
import pytest

import oracle_ndt as O

HERE = os.path.dirname(os.path.abspath(__file__))


def test_gauss_constants_known_answers():
    # values of ndt_omp_impl2.hpp:93-100 for the reference's parameters
    o = O.OracleNDT(resolution=1.0, outlier_ratio=0.55)
    np.testing.assert_allclose(o.gauss(), [-2.217225244043, 0.433123004704, 0.597837000756], rtol=0, atol=5e-12)
    o = O.OracleNDT(resolution=0.5, outlier_ratio=0.55)
    np.testing.assert_allclose(o.gauss(), [-0.704446735814, 0.756362730327, -1.481604540924], rtol=0, atol=5e-12)


def test_se3_exp_log_round_trip():
    rng = np.random.default_rng(0)
    for _ in range(200):
        p = np.concatenate([rng.normal(0, 5, 3), rng.normal(0, 1, 3) * rng.uniform(0, 1)])
        q, t = O.se3_exp(p)
        assert abs(np.linalg.norm(q) - 1) < 1e-14
        back = O.se3_compose_log(np.zeros(6), p)          # log(exp(0) * exp(p))
        np.testing.assert_allclose(back, p, rtol=0, atol=1e-10)
    # tiny angles take the Taylor branches (so3.cpp:180-186, se3.cpp:176-180,208)
    p = np.array([1.0, -2.0, 0.5, 1e-12, -2e-12, 1e-13])
    np.testing.assert_allclose(O.se3_compose_log(np.zeros(6), p), p, rtol=0, atol=1e-10)
    # composition agrees with matrix products
    a, b = np.array([0.3, -0.1, 0.2, 0.02, 0.01, -0.4]), np.array([1.0, 2.0, -0.5, -0.3, 0.2, 0.1])
    Ta, Tb = O.se3_exp_matrix4f(a).astype(np.float64), O.se3_exp_matrix4f(b).astype(np.float64)
    Tab = O.se3_exp_matrix4f(O.se3_compose_log(a, b)).astype(np.float64)
    np.testing.assert_allclose(Tab, Ta @ Tb, atol=2e-6)


def _sref_call(fn, *arrays_out_last):
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in arrays_out_last]
    fn(*[a.ctypes.data for a in arrs])
    return arrs


def test_se3_restatement_against_the_reference_sophus_sources():
    """oracle/ose3.h against the reference's own Sophus (so3.cpp / se3.cpp of the vendored Sophus a621ff2, compiled by oracle/build_ref.sh
    against a stand-in for the Eigen operations they use): the three uses the NDT path makes of Sophus - SE3::exp(p).matrix(),
    SE3(R, t).log(), (SE3::exp(d) * SE3::exp(p)).log() (ndt_omp_impl2.hpp:119-120,161-166) - on random, tiny-angle (both sides of
    SMALL_EPS), near-pi and zero arguments.  Sophus' formulas, series coefficients, branches and normalisation points are pinned by this;
    Eigen's rounding is not (oracle/ref_stubs/eigen_min.h)."""
    R = O.sophus_ref()
    if R is None:
        pytest.skip("no compiled reference Sophus (needs /root/reference or a prebuilt oracle/_ref/libsophus_ref.so)")
    assert R.sref_selftest() < 1e-10                        # Sophus' own test_se3.cpp cases and bound on the compiled sources
    rng = np.random.default_rng(11)
    ps = [np.zeros(6)]
    for k in range(200):
        ps.append(rng.normal(0, [5, 5, 2, 0.5, 0.5, 0.5]))
    for ang in (1e-13, 9e-11, 1.1e-10, 1e-8, 1e-5):         # the Taylor branches switch at theta < 1e-10
        u = rng.normal(size=3); u /= np.linalg.norm(u)
        ps.append(np.concatenate([rng.normal(0, 3, 3), ang * u]))
    for k in range(20):                                     # rotations near pi (|w| of the quaternion near 0)
        u = rng.normal(size=3); u /= np.linalg.norm(u)
        ps.append(np.concatenate([rng.normal(0, 3, 3), (np.pi - rng.uniform(0, 1e-6)) * u]))
    worst = dict(exp=0.0, mat=0.0, log=0.0, comp=0.0)
    for p in ps:
        q, t = O.se3_exp(p)
        _, qr, tr = _sref_call(R.sref_se3_exp, p, np.zeros(4), np.zeros(3))
        worst["exp"] = max(worst["exp"], np.abs(q - qr).max(), np.abs(t - tr).max() / max(1.0, np.abs(tr).max()))
        _, Mr = _sref_call(R.sref_se3_exp_matrix, p, np.zeros(16))
        M = O.se3_exp_matrix4f(p)
        assert np.array_equal(M, Mr.reshape(4, 4).astype(np.float32))          # the float cast the reference applies (:161-163)
        # SE3(R, t).log() of the float matrix, as computeTransformation reads its guess
        Rd, td = M[:3, :3].astype(np.float64), M[:3, 3].astype(np.float64)
        _, _, lr = _sref_call(R.sref_se3_log_of_Rt, Rd.reshape(9), td, np.zeros(6))
        lo = O.se3_log_from_matrix4f(M)
        worst["log"] = max(worst["log"], np.abs(lo - lr).max() / max(1.0, np.abs(lr).max()))
        d = rng.normal(0, [0.1, 0.1, 0.1, 0.01, 0.01, 0.01])
        _, _, cr = _sref_call(R.sref_compose_log, d, p, np.zeros(6))
        worst["comp"] = max(worst["comp"], np.abs(O.se3_compose_log(d, p) - cr).max() / max(1.0, np.abs(cr).max()))
    assert max(worst.values()) <= 1e-15, worst


@pytest.mark.parametrize("search", [O.DIRECT7, O.DIRECT1, O.DIRECT26, O.KDTREE])
def test_restatement_against_the_reference_member_functions(small_pair, search):
    """The oracle against the reference's OWN code: computeTransformation, computeDerivatives, computePointDerivatives_AngleAxisd,
    updateDerivatives, computeHessian, updateHessian, computeStepLengthMT, updateIntervalMT, trialValueSelectionMT and calculateScore are
    taken verbatim from include/ndt_omp/ndt_omp_impl2.hpp at build time and compiled in oracle/ndt_ref_harness.cpp (with the reference's
    Sophus) against stand-ins for Eigen's interface, the class declaration and the voxel grid; both sides run one thread on the same
    voxel cells.  Bar: every number identical - derivative passes in all four search modes, the all-double Hessian, calculateScore, whole
    aligns (iterations, convergence flag, final transformation, probability, aligned cloud), the forced More-Thuente path included."""
    if O.ndt_ref_lib() is None:
        pytest.skip("no compiled reference NDT (needs /root/reference or a prebuilt oracle/_ref/libndt_ref.so)")
    tgt, src, guess, truth = small_pair
    o = O.OracleNDT(variant=O.VAR_OMP, search=search, num_threads=1, trans_eps=0.01, max_iter=30)
    o.set_target(tgt); o.set_source(src)
    r = O.ReferenceNDT(o); r.set_source(src)
    rng = np.random.default_rng(21)
    p0 = O.se3_log_from_matrix4f(guess)
    for k in range(3):
        p = p0 if k == 0 else p0 + rng.normal(0, [0.05, 0.05, 0.02, 0.004, 0.004, 0.01])
        T = guess if k == 0 else None
        for hess in (True, False):
            so, go, Ho = o.eval_derivatives(p, T, hess)
            sr, gr, Hr = r.eval_derivatives(p, T, hess)
            assert abs(so) > 100 and so == sr and np.array_equal(go, gr) and np.array_equal(Ho, Hr), (k, hess)
        assert np.array_equal(o.eval_hessian(p, T), r.eval_hessian(p, T))
    assert o.calculate_score(guess) == r.calculate_score(guess)
    ao, ar = o.align(guess, want_cloud=True), r.align(guess)
    assert ao["iterations"] == ar["iterations"] > 2 and ao["converged"] == ar["converged"]
    assert np.array_equal(ao["final"], ar["final"]) and ao["trans_probability"] == ar["trans_probability"] and np.array_equal(ao["cloud"], ar["cloud"])
    if search == O.DIRECT7:
        # step_size <= epsilon / 2 is the only way the More-Thuente loop and computeHessian run (the `(step_max - step_min) > 0` quirk)
        for ss, eps, g0 in ((0.2, 0.5, truth.astype(np.float32)), (0.2, 0.5, guess), (0.1, 0.3, truth.astype(np.float32))):
            o2 = O.OracleNDT(variant=O.VAR_OMP, search=search, num_threads=1, step_size=ss, trans_eps=eps, max_iter=6)
            o2.set_target(tgt); o2.set_source(src)
            r2 = O.ReferenceNDT(o2); r2.set_source(src)
            a2, b2 = o2.align(g0, want_cloud=True), r2.align(g0)
            assert a2["n_hess"] > 0 and a2["iterations"] == b2["iterations"] and a2["converged"] == b2["converged"]
            assert np.array_equal(a2["final"], b2["final"]) and a2["trans_probability"] == b2["trans_probability"] and np.array_equal(a2["cloud"], b2["cloud"])


@pytest.mark.parametrize("variant,search,kw", [(O.VAR_PCA, O.DIRECT1, {}), (O.VAR_PCA, O.DIRECT7, {}), (O.VAR_PCA, O.KDTREE, {}),
                                               (O.VAR_GROUND, O.DIRECT1, dict(resolution=2.0)), (O.VAR_GROUND, O.DIRECT7, dict(resolution=2.0)),
                                               (O.VAR_GROUND, O.DIRECT26, {}), (O.VAR_GROUND, O.KDTREE, dict(resolution=2.0)),
                                               (O.VAR_GROUND, O.DIRECT1, dict(resolution=10.0, max_iter=64))])
def test_pca_and_ground_restatements_against_the_reference_member_functions(small_pair, variant, search, kw):
    """The same pin for the two other registration classes: pclpca (include/ndt_pca/ndt_pca_impl2.hpp: the per-cell weight on the running
    sums) and pclomp_ground (include/ndt_omp/ndt_ground_impl.hpp: computeDerivatives_seg with flag_class = 1 - the double updateDerivatives call,
    the last-neighbour gate on the leaf normal, the zeroed rows and columns - and its computeTransformation without the first-iteration
    clause).  The per-cell weights and eigenvectors handed to the reference's code are the restatement's."""
    if O.ndt_ref_lib(variant) is None:
        pytest.skip("no compiled reference NDT")
    tgt, src, guess, truth = small_pair
    prm = dict(trans_eps=0.01, max_iter=30)
    prm.update(kw)
    o = O.OracleNDT(variant=variant, search=search, num_threads=1, **prm)
    o.set_target(tgt); o.set_source(src)
    r = O.ReferenceNDT(o); r.set_source(src)
    rng = np.random.default_rng(22)
    p0 = O.se3_log_from_matrix4f(guess)
    for k in range(3):
        p = p0 if k == 0 else p0 + rng.normal(0, [0.05, 0.05, 0.02, 0.004, 0.004, 0.01])
        T = guess if k == 0 else None
        for hess in (True, False):
            so, go, Ho = o.eval_derivatives(p, T, hess)
            sr, gr, Hr = r.eval_derivatives(p, T, hess)
            assert abs(so) > 10 and so == sr and np.array_equal(go, gr) and np.array_equal(Ho, Hr), (k, hess)
    ao, ar = o.align(guess, want_cloud=True), r.align(guess)
    assert ao["iterations"] == ar["iterations"] >= 1 and ao["converged"] == ar["converged"]
    assert np.array_equal(ao["final"], ar["final"]) and ao["trans_probability"] == ar["trans_probability"] and np.array_equal(ao["cloud"], ar["cloud"])
    # forced More-Thuente path (for pclomp_ground: computeStepLengthMT with flag_class, then the unmasked computeHessian)
    o2 = O.OracleNDT(variant=variant, search=search, num_threads=1, step_size=0.2, trans_eps=0.5, max_iter=4, resolution=prm.get("resolution", 1.0))
    o2.set_target(tgt); o2.set_source(src)
    r2 = O.ReferenceNDT(o2); r2.set_source(src)
    a2, b2 = o2.align(truth.astype(np.float32), want_cloud=True), r2.align(truth.astype(np.float32))
    assert a2["iterations"] == b2["iterations"] and a2["converged"] == b2["converged"] and np.array_equal(a2["final"], b2["final"])
    assert a2["trans_probability"] == b2["trans_probability"] and np.array_equal(a2["cloud"], b2["cloud"])


@pytest.mark.parametrize("leaf", [1.0, 0.5, 2.0, 0.7])
def test_voxel_build_against_the_reference_member_functions(small_pair, leaf):
    """VoxelGridCovariance::applyFilter and getNeighborhoodAtPoint{,7,1}, taken verbatim from voxel_grid_covariance_omp_impl.hpp at build
    time and compiled in oracle/voxel_ref_harness.cpp with the class declaration of voxel_grid_covariance_omp.h - in particular the Leaf
    constructor that starts cov_ at the identity - against the restatement: grid geometry, keys, counts, float centroids, means,
    covariances (after the eigenvalue inflation), inverse covariances, eigenvalues and eigenvectors of EVERY cell identical; NaN points
    and a strided cloud included; the direct searches return the same cells in the same order."""
    if not O.ReferenceVoxelGrid.available():
        pytest.skip("no compiled reference voxel grid")
    tgt = small_pair[0].copy()
    tgt[::97, 1] = np.nan                                   # skipped by applyFilter (the is_dense = false branch)
    cloud = np.zeros((len(tgt), 8), np.float32)
    cloud[:, :3] = tgt
    r = O.ReferenceVoxelGrid(cloud, leaf)
    o = O.OracleNDT(search=O.DIRECT7, num_threads=1, resolution=leaf)
    o.set_target(cloud)
    assert [a.tolist() for a in r.grid()] == [a.tolist() for a in o.grid()]
    rl, ol = r.leaves(), o.leaves()
    for k in ("keys", "nr_points", "centroid", "mean", "cov", "icov", "evals"):
        assert np.array_equal(rl[k], ol[k]), k
    assert np.array_equal(rl["evecs"], o.leaf_evecs())
    n6 = ol["nr_points"] == 6                                # the identity the sums start from: + (n - 1) / n^2 on the diagonal
    assert n6.any() and (ol["cov"][n6][:, [0, 1, 2], [0, 1, 2]].min(axis=1) > 5 / 36 - 1e-9).all()
    rng = np.random.default_rng(31)
    pts = tgt[rng.integers(0, len(tgt), 300)] + rng.normal(0, 0.3, (300, 3)).astype(np.float32)
    pts = pts[np.isfinite(pts).all(axis=1)]
    for mode_o, mode_r in ((O.DIRECT26, 1), (O.DIRECT7, 2), (O.DIRECT1, 3)):
        hits = 0
        for p in pts:
            a, b = o.neighbours(p, mode_o), r.neighbours(p, mode_r)
            assert np.array_equal(a, b), (mode_o, p)
            hits += len(a)
        assert hits > 50


@pytest.mark.parametrize("which", ["small", "full"])
def test_pca_voxel_weights_against_the_reference_member_functions(small_pair, scan_pair, which):
    """pclpca::VoxelGridCovariance::applyFilter (include/ndt_pca/voxel_grid_covariance_pca_impl.hpp, taken at build time): the dimension label
    and the integer weight `getDimension2d()` of every usable cell, and the whole grid, identical to the restatement's.  Cells that never
    reach the label block (fewer than six points) keep the constructor's dimension_2d_ = 0 in the reference and report weight 1 here (and
    on the device): they are invisible to every search of the path."""
    if not O.ReferenceVoxelGrid.available(pca=True):
        pytest.skip("no compiled reference voxel grid")
    tgt = (small_pair if which == "small" else scan_pair)[0]
    r = O.ReferenceVoxelGrid(tgt, 1.0, pca=True)
    o = O.OracleNDT(variant=O.VAR_PCA, search=O.DIRECT1, num_threads=1)
    o.set_target(tgt)
    rl, ol = r.leaves(), o.leaves()
    for k in ("keys", "nr_points", "centroid", "mean", "cov", "icov", "evals"):
        assert np.array_equal(rl[k], ol[k]), k
    label, weight = r.pca()
    usable = ol["nr_points"] >= 6
    assert usable.sum() > 300 and np.array_equal(label[usable], ol["label"][usable]) and np.array_equal(weight[usable], ol["weight"][usable])
    assert len(set(label[usable].tolist())) >= 2 and weight[usable].max() > 10
    assert not weight[~usable].any() and (ol["weight"][~usable] == 1).all()


def test_information_matrix_against_the_reference_source(small_pair):
    """lv_slam::InformationMatrixCalculator - the reference's own src/global_graph/information_matrix_calculator.cpp compiled whole against stand-in
    ROS / PCL / Eigen headers - against the restatement of its fitness score (oracle) and the host mirror of its weighting
    (lv_slam_b200/information_matrix.py): constructor defaults, the launch file's threshold, the constant-matrix switch."""
    if O.info_ref() is None:
        pytest.skip("no compiled reference information matrix calculator")
    from lv_slam_b200.information_matrix import InformationMatrixCalculator
    tgt, src, guess, truth = small_pair
    tgt, src = tgt[::3], src[::3]                           # the stand-in kd-tree is exhaustive
    o = O.OracleNDT(num_threads=8)
    o.set_target(tgt); o.set_source(src)
    big = float(np.finfo(np.float64).max)
    for T in (truth, guess.astype(np.float64)):
        for mr in (big, 0.25, 1e-4):
            want = O.ref_fitness_score(tgt, src, T, mr)
            got, cnt = o.fitness_score(np.asarray(T, dtype=np.float64).astype(np.float32), mr)
            assert got == want or abs(got - want) <= 1e-15 * abs(want), (mr, got, want)
            assert (cnt == 0) == (want == big)
    for prm in ({}, dict(fitness_score_thresh=2.0), dict(var_gain_a=10.0, min_stddev_x=0.2, max_stddev_q=0.5), dict(use_const_inf_matrix=1, const_stddev_x=0.3)):
        want = O.ref_information_matrix(tgt, src, guess.astype(np.float64), **prm)
        kw = dict(prm)
        if "use_const_inf_matrix" in kw:
            kw["use_const_inf_matrix"] = True
        m = InformationMatrixCalculator(**kw)
        fs = O.ref_fitness_score(tgt, src, guess.astype(np.float64))
        got = m.information_from_fitness(fs) if not m.use_const_inf_matrix else m.calc_information_matrix(None, None, None)
        assert np.array_equal(got, want), (prm, got.diagonal(), want.diagonal())


def test_eigen_stand_in_against_numpy():
    """oracle/ref_stubs/eigen_min.h - the interface stand-in the reference's own code is compiled against - does what the Eigen operations of the
    same name do: small-matrix algebra, quaternion <-> matrix, rotation of a vector, angle-axis products, isometry product and inverse."""
    import ctypes
    import subprocess
    root = os.path.dirname(HERE)
    so = os.path.join(root, "oracle", "_ref", "libeigen_min_selftest.so")
    if not os.path.exists(so):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=gnu++17", "-ffp-contract=off", "-fPIC", "-shared", "-I" + os.path.join(root, "oracle", "ref_stubs"), "-o", so,
                               os.path.join(root, "oracle", "ref_stubs", "selftest_api.cpp")])
    E = ctypes.CDLL(so)
    vp, f64 = ctypes.c_void_p, ctypes.c_double
    E.est_matrix3.restype = None; E.est_matrix3.argtypes = [vp] * 5
    E.est_quaternion.restype = None; E.est_quaternion.argtypes = [vp, vp, vp, f64, f64, vp]
    E.est_isometry.restype = None; E.est_isometry.argtypes = [vp] * 3
    rng = np.random.default_rng(61)

    def rot(q):
        w, x, y, z = q
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])

    def qmul(a, b):
        return np.array([a[0] * b[0] - a[1:] @ b[1:], *(a[0] * b[1:] + b[0] * a[1:] + np.cross(a[1:], b[1:]))])

    for k in range(50):
        A, B = rng.normal(size=(3, 3)) + 2 * np.eye(3), rng.normal(size=(3, 3))
        v, w = rng.normal(size=3), rng.normal(size=3)
        out = np.zeros(44)
        E.est_matrix3(np.ascontiguousarray(A).ctypes.data, np.ascontiguousarray(B).ctypes.data, v.ctypes.data, w.ctypes.data, out.ctypes.data)
        np.testing.assert_allclose(out[0:9].reshape(3, 3), A @ B, atol=1e-14)
        np.testing.assert_allclose(out[9:18].reshape(3, 3), A.T @ B, atol=1e-14)
        np.testing.assert_allclose(out[18:27].reshape(3, 3), np.linalg.inv(A), atol=1e-12)
        np.testing.assert_allclose(out[27:30], A @ v, atol=1e-14)
        np.testing.assert_allclose([out[30], out[34]], [v @ w, np.linalg.norm(v)], atol=1e-14)
        np.testing.assert_allclose(out[31:34], np.cross(v, w), atol=1e-14)
        np.testing.assert_allclose(out[35:44].reshape(3, 3), np.outer(v, w), atol=1e-14)
        q, p = rng.normal(size=4), rng.normal(size=4)
        q /= np.linalg.norm(q); p /= np.linalg.norm(p)
        if k % 3 == 0:
            q[0] = abs(q[0]) * 0.01; q /= np.linalg.norm(q)                  # trace of the rotation below zero: the other branches of the conversion
        a, b = rng.uniform(-3, 3), rng.uniform(-1.5, 1.5)
        out = np.zeros(29)
        E.est_quaternion(q.ctypes.data, p.ctypes.data, v.ctypes.data, a, b, out.ctypes.data)
        R = rot(q)
        np.testing.assert_allclose(out[0:9].reshape(3, 3), R, atol=1e-14)
        back = out[9:13]
        assert min(np.abs(back - q).max(), np.abs(back + q).max()) < 1e-12     # the matrix determines the quaternion up to sign
        np.testing.assert_allclose(out[13:16], R @ v, atol=1e-13)
        np.testing.assert_allclose(out[16:20], qmul(q, p), atol=1e-14)
        Rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
        Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
        np.testing.assert_allclose(out[20:29].reshape(3, 3), Rz @ Ry, atol=1e-14)
        Ta, Tb = np.eye(4), np.eye(4)
        Ta[:3, :3], Ta[:3, 3] = rot(q), rng.normal(0, 5, 3)
        Tb[:3, :3], Tb[:3, 3] = rot(p), rng.normal(0, 5, 3)
        out = np.zeros(44)
        E.est_isometry(np.ascontiguousarray(Ta).ctypes.data, np.ascontiguousarray(Tb).ctypes.data, out.ctypes.data)
        np.testing.assert_allclose(out[0:16].reshape(4, 4), Ta @ Tb, atol=1e-13)
        np.testing.assert_allclose(out[16:32].reshape(4, 4), np.linalg.inv(Ta), atol=1e-13)
        np.testing.assert_allclose(out[32:35], Ta[:3, 3], atol=0)
        np.testing.assert_allclose(out[35:44].reshape(3, 3), Ta[:3, :3], atol=0)


def test_log_of_float_guess_matches_matrix():
    T = np.eye(4, dtype=np.float32)
    T[0, 3] = 1.5                                          # the reference's first-frame guess (scan_matching_odom_nodelet.cpp:199-200)
    np.testing.assert_allclose(O.se3_log_from_matrix4f(T), [1.5, 0, 0, 0, 0, 0], atol=1e-15)


def test_small_dense_solvers_against_numpy():
    rng = np.random.default_rng(1)
    for _ in range(50):
        A = rng.normal(size=(6, 6)) * rng.uniform(0.1, 1e4, size=(6, 1))
        b = rng.normal(size=6)
        x, sv = O.svd6_solve(A, b)
        np.testing.assert_allclose(x, np.linalg.solve(A, b), rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(sv, np.linalg.svd(A, compute_uv=False), rtol=1e-10)
    # rank-deficient: pseudo-inverse semantics of JacobiSVD::solve
    A = np.zeros((6, 6)); A[:3, :3] = np.diag([2.0, 3.0, 4.0]); b = np.arange(1.0, 7.0)
    x, _ = O.svd6_solve(A, b)
    np.testing.assert_allclose(x, np.linalg.pinv(A) @ b, atol=1e-14)
    for _ in range(50):
        M = rng.normal(size=(3, 3)); S = M @ M.T
        ev, V = O.sym3_eig(S)
        np.testing.assert_allclose(ev, np.linalg.eigvalsh(S), rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(V @ np.diag(ev) @ V.T, S, atol=1e-12)


def test_single_voxel_hand_case():
    """Eight points in one 1 m cell: mean, the one-pass covariance ON TOP OF THE IDENTITY the Leaf constructor starts cov_ with
    (voxel_grid_covariance_omp.h:98-106; applyFilter :240,:329-330), the (n-1)/n factor, inverse; and a cell where the eigenvalue inflation fires."""
    pts = np.array([[0.2, 0.2, 0.5], [0.8, 0.2, 0.5], [0.2, 0.8, 0.5], [0.8, 0.8, 0.5], [0.5, 0.5, 0.5], [0.4, 0.6, 0.5], [0.6, 0.4, 0.5],
                    [0.5, 0.5, 0.5]], dtype=np.float32)
    o = O.OracleNDT()
    o.set_target(pts)
    lv = o.leaves()
    assert lv["keys"].tolist() == [0] and lv["nr_points"].tolist() == [8]
    p = pts.astype(np.float64)
    mean = p.mean(axis=0)
    np.testing.assert_allclose(lv["mean"][0], mean, atol=1e-15)
    cov = ((p - mean).T @ (p - mean) + np.eye(3)) / 8 * (7 / 8)      # (I + sum p p^T - 2 s m^T) / n + m m^T, times (n-1)/n
    np.testing.assert_allclose(lv["cov"][0], cov, atol=1e-12)
    ev = np.linalg.eigvalsh(cov)
    assert ev[0] > 0.01 * ev[2]                             # the identity keeps the planar patch away from the inflation: 7/64 on the diagonal
    np.testing.assert_allclose(lv["evals"][0], ev, rtol=1e-9)
    np.testing.assert_allclose(lv["icov"][0] @ lv["cov"][0], np.eye(3), atol=1e-9)
    # 160 points of a 10 m cell: the identity's share, 159 / 160^2, is below 1 % of the largest eigenvalue and the smallest one is inflated
    big = np.tile(pts * 10, (20, 1)).astype(np.float32)
    o10 = O.OracleNDT(resolution=10.0)
    o10.set_target(big)
    l10 = o10.leaves()
    pb = big.astype(np.float64)
    covb = ((pb - pb.mean(axis=0)).T @ (pb - pb.mean(axis=0)) + np.eye(3)) / 160 * (159 / 160)
    evb = np.linalg.eigvalsh(covb)
    assert l10["nr_points"].tolist() == [160] and evb[0] < 0.01 * evb[2]
    np.testing.assert_allclose(l10["evals"][0], [0.01 * evb[2], evb[1], evb[2]], rtol=1e-9)
    np.testing.assert_allclose(l10["icov"][0] @ l10["cov"][0], np.eye(3), atol=1e-9)
    # five points: below min_points_per_voxel, the cell exists but is not usable
    o.set_target(pts[:5])
    lv = o.leaves()
    assert lv["nr_points"].tolist() == [5] and not lv["icov"].any()
    o.set_source(pts)
    s, g, H = o.eval_derivatives(np.zeros(6))
    assert s == 0 and not g.any() and not H.any()


def test_voxel_index_matches_reference_formula(small_pair):
    tgt = small_pair[0]
    o = O.OracleNDT()
    o.set_target(tgt)
    mn, mx, dv = o.grid()
    inv = np.float32(1.0)
    ijk = (np.floor(tgt * inv) - mn.astype(np.float32)).astype(np.int32)
    keys = ijk[:, 0] + ijk[:, 1] * dv[0] + ijk[:, 2] * dv[0] * dv[1]
    uk, cnt = np.unique(keys, return_counts=True)
    lv = o.leaves()
    assert np.array_equal(lv["keys"], uk) and np.array_equal(lv["raw_points"], cnt)
    assert np.array_equal(o.lookup_keys(tgt), keys)         # leaf 1.0 is a power of two: mul and div index forms agree


def _fd(f, p, i, h):
    e = np.zeros(6); e[i] = h
    return (f(p + e) - f(p - e)) / (2 * h)


def _interior_points(src, T, margin):
    """Source points whose transformed position stays `margin` away from every voxel face: the NDT score is only piecewise
    smooth (a point changes neighbourhood when it crosses a face), so finite differences are taken where no point crosses."""
    x = O.transform(src, T).astype(np.float64)
    fr = x - np.floor(x)
    return src[(np.minimum(fr, 1 - fr).min(axis=1) > margin)]


@pytest.mark.parametrize("variant,search", [(O.VAR_OMP, O.DIRECT7), (O.VAR_OMP, O.DIRECT1), (O.VAR_PCA, O.DIRECT1), (O.VAR_PCA, O.DIRECT7)])
def test_gradient_and_hessian_against_finite_differences(small_pair, variant, search):
    tgt, src, guess, truth = small_pair
    o = O.OracleNDT(variant=variant, search=search, num_threads=1)
    o.set_target(tgt)
    score = lambda p: o.eval_derivatives(p, None, False)[0]
    grad = lambda p: o.eval_derivatives(p, None, False)[1]
    ht, hr = 2e-3, 1e-4
    # (a) p = 0: the reference's Jacobian [I | -[R x]x] is the exact derivative in all six directions
    p0 = np.zeros(6)
    o.set_source(_interior_points(src, np.eye(4, dtype=np.float32), 0.03))
    s, g, H = o.eval_derivatives(p0)
    assert abs(s) > 100
    for i in range(6):
        fd = _fd(score, p0, i, ht if i < 3 else hr)
        assert abs(fd - g[i]) <= 5e-3 * np.abs(g[:3] if i < 3 else g[3:]).max(), (i, fd, g[i])
    # (b) a pure translation: exact for the translation block anywhere; the Hessian block is the derivative of the gradient
    pt = np.array([0.4, -0.2, 0.05, 0, 0, 0.0])
    o.set_source(_interior_points(src, O.se3_exp_matrix4f(pt), 0.03))
    s, g, H = o.eval_derivatives(pt)
    for i in range(3):
        assert abs(_fd(score, pt, i, ht) - g[i]) <= 5e-3 * np.abs(g[:3]).max()
        col = np.array([_fd(lambda q: grad(q)[j], pt, i, ht) for j in range(3)])
        np.testing.assert_allclose(col, H[:3, i], rtol=0, atol=1e-2 * np.abs(H[:3, :3]).max())
    assert np.abs(H - H.T).max() > 0                        # quirk 9: the rot-rot block is not symmetric
    np.testing.assert_allclose(H[:3, :3], H[:3, :3].T, rtol=0, atol=1e-5 * np.abs(H[:3, :3]).max())


def test_double_hessian_close_to_float_hessian(small_pair):
    tgt, src, guess, truth = small_pair
    o = O.OracleNDT(search=O.KDTREE, num_threads=1)
    o.set_target(tgt); o.set_source(src[::5])
    p = O.se3_log_from_matrix4f(guess)
    _, _, Hf = o.eval_derivatives(p)
    Hd = o.eval_hessian(p)
    np.testing.assert_allclose(Hf, Hd, rtol=0, atol=1e-4 * np.abs(Hd).max())     # same neighbours, float32 vs float64 terms


def test_align_recovers_known_motion_and_quirks(scan_pair):
    tgt, src, guess, truth = scan_pair
    o = O.OracleNDT(trans_eps=0.01, max_iter=64, search=O.DIRECT7, num_threads=os.cpu_count() or 1)
    o.set_target(tgt); o.set_source(src)
    r = o.align(guess)
    assert r["converged"] and r["iterations"] >= 2          # quirk 7: at least two iterations
    assert r["n_eval"] == r["iterations"] + 1 and r["n_hess"] == 0   # `interval_converged = (step_max - step_min) > 0` disables the MT loop
    assert np.abs(r["final"][:3, 3] - truth[:3, 3]).max() < 0.05
    assert (np.abs(r["trace"][:, 12]) <= 0.1 + 1e-12).all() # steps are clamped to step_size
    # final_transformation_ is exp(x_t) of the LAST line-search point, i.e. p_before + dir * step of the last iteration (quirk 2)
    last = r["trace"][-1]
    np.testing.assert_array_equal(r["final"], O.se3_exp_matrix4f(last[0:6] + last[6:12] * last[12]))
    # max_iterations + 2 iterations are possible (quirk 7)
    o2 = O.OracleNDT(trans_eps=1e-9, max_iter=3, search=O.DIRECT7, num_threads=1)
    o2.set_target(tgt[::8]); o2.set_source(src[::8])
    assert o2.align(guess)["iterations"] == 5


def test_golden_regression_of_the_oracle(small_pair):
    """tests/golden/ndt_small_pair.json was written by tests/golden/make_golden.py from THIS oracle (not from the reference,
    which cannot be built here): it pins the oracle against accidental edits."""
    path = os.path.join(HERE, "golden", "ndt_small_pair.json")
    gold = json.load(open(path))
    tgt, src, guess, truth = small_pair
    o = O.OracleNDT(trans_eps=0.01, max_iter=30, search=O.DIRECT7, num_threads=1)
    o.set_target(tgt); o.set_source(src)
    lv = o.leaves()
    assert len(lv["keys"]) == gold["n_cells"] and int((lv["nr_points"] >= 6).sum()) == gold["n_usable"]
    assert int(lv["keys"].astype(np.int64).sum()) == gold["key_sum"]
    s, g, H = o.eval_derivatives(O.se3_log_from_matrix4f(guess), guess)
    np.testing.assert_allclose(s, gold["score"], rtol=1e-12)
    np.testing.assert_allclose(g, gold["gradient"], rtol=1e-10, atol=1e-9 * np.abs(g).max())
    np.testing.assert_allclose(H, np.array(gold["hessian"]), rtol=1e-10, atol=1e-9 * np.abs(H).max())
    r = o.align(guess)
    assert r["iterations"] == gold["iterations"]
    np.testing.assert_allclose(r["final"], np.array(gold["final"], dtype=np.float32), atol=1e-6)


def test_fitness_score_known_answers():
    """getFitnessScore restatement on a hand case: squared nearest-neighbour distances, capped by max_range, mean in double."""
    tgt = np.array([[0, 0, 0], [1, 0, 0], [0, 2, 0], [np.nan, 0, 0]], np.float32)
    src = np.array([[0.1, 0, 0], [0.9, 0.1, 0], [5, 5, 5], [0, np.inf, 0]], np.float32)
    o = O.OracleNDT()
    o.set_target(tgt); o.set_source(src)
    I = np.eye(4, dtype=np.float32)
    d = [np.float32(0.1) ** 2, np.float32(np.float32(0.1) ** 2 + np.float32(0.1) ** 2)]
    d0 = np.float32(np.float32(0.1) * np.float32(0.1))
    d1 = np.float32(np.float32(np.float32(0.9) - np.float32(1.0)) ** 2 + np.float32(0.1) * np.float32(0.1))
    d2 = np.float32(25 + 9 + 25)
    s, n = o.fitness_score(I)
    assert n == 3 and abs(s - (float(d0) + float(d1) + float(d2)) / 3) < 1e-15 * s      # the non-finite source point is skipped
    s, n = o.fitness_score(I, 1.0)
    assert n == 2 and s == (float(d0) + float(d1)) / 2
    s, n = o.fitness_score(I, 1e-6)
    assert n == 0 and s == np.finfo(np.float64).max
    T = I.copy(); T[0, 3] = -0.1                                                      # moves the first point onto a target point
    s, n = o.fitness_score(T, 0.5)
    x1 = np.float32(np.float32(0.9) + np.float32(-0.1))
    e1 = np.float32(np.float32(x1 - np.float32(1.0)) ** 2 + np.float32(0.1) * np.float32(0.1))
    assert n == 2 and s == (0.0 + float(e1)) / 2


def test_prefilter_known_answers():
    """distance_filter + pcl::VoxelGrid restatement on a hand case (prefiltering_nodelet.cpp:164-181, 138-148)."""
    pts = np.array([[0.2, 0.0, 0.0, 1.0],      # |p| = 0.2 < near: dropped
                    [1.02, 0.0, 0.0, 2.0],     # leaf (10, 0, 0) at 0.1 m
                    [1.04, 0.0, 0.0, 4.0],     # same leaf
                    [3.0, 4.0, 0.0, 5.0],      # |p| = 5
                    [60.0, 80.0, 0.0, 7.0],    # |p| = 100: not < far, dropped
                    [np.nan, 1.0, 1.0, 9.0],   # norm is NaN: dropped
                    [-1.0, -1.0, 0.5, 3.0]], np.float32)
    out, fl = O.prefilter(pts, near=0.5, far=100.0, leaf=0.1)
    assert fl == 0 and out.shape == (3, 4)
    # ascending leaf index = x fastest, then y, then z (divb_mul = 1, dx, dx*dy): z = 0 leaves before the z = 0.5 one, and within
    # z = 0 the y = 0 row before y = 4
    a = np.float32(np.float32(1.02) + np.float32(1.04)) / np.float32(2)
    assert np.array_equal(out[0], np.array([a, 0, 0, 3.0], np.float32))
    assert np.array_equal(out[1], pts[3]) and np.array_equal(out[2], pts[6])
    # no downsampling: the kept points in input order; no distance filter: NaN rows survive the filter and vanish in the grid
    out, _ = O.prefilter(pts, near=0.5, far=100.0, leaf=0.0)
    assert np.array_equal(out, pts[[1, 2, 3, 6]])
    out, _ = O.prefilter(pts, use_filter=False, leaf=0.1)
    assert out.shape[0] == 5
    # PCL's overflow guard: 1e-4 m leaves over a 140 m extent do not fit int32 -> the cloud passes through, flag set
    out, fl = O.prefilter(pts[:, :3], near=0.5, far=1000.0, leaf=1e-4)
    assert fl == 1 and np.array_equal(out, pts[[1, 2, 3, 4, 6], :3])


def test_golden_regression_of_fitness_and_prefilter(small_pair):
    """tests/golden/aux_small_pair.json (written by make_golden.py from this oracle): the restatements of getFitnessScore and of the
    prefilter chain pinned against accidental edits; the correspondence counts and the output size are exact."""
    gold = json.load(open(os.path.join(HERE, "golden", "aux_small_pair.json")))
    tgt, src, guess, truth = small_pair
    o = O.OracleNDT(trans_eps=0.01, max_iter=30, search=O.DIRECT7, num_threads=1)
    o.set_target(tgt); o.set_source(src)
    big = float(np.finfo(np.float64).max)
    for name, T, mr in (("truth", truth, big), ("guess", guess, big), ("guess_capped", guess, 0.25)):
        sc, cnt = o.fitness_score(T, mr)
        assert cnt == gold["fitness"][name]["correspondences"]
        np.testing.assert_allclose(sc, gold["fitness"][name]["score"], rtol=1e-12)
    cloud = np.concatenate([tgt[:, :3], np.random.default_rng(3).random((len(tgt), 1), dtype=np.float32)], axis=1).astype(np.float32)
    pf, _ = O.prefilter(cloud, 0.5, 100.0, True, 0.1)
    g = gold["prefilter"]
    assert len(cloud) == g["n_in"] and len(pf) == g["n_out"]
    np.testing.assert_allclose(pf.astype(np.float64).sum(axis=0), g["column_sums"], rtol=1e-12)
    np.testing.assert_array_equal(pf[0], np.array(g["first"], np.float32))
    np.testing.assert_array_equal(pf[-1], np.array(g["last"], np.float32))


def test_expf_restatement_matches_the_host_libm():
    """updateDerivatives calls exp on a float, which binds to glibc's expf (oracle/ndt_oracle.cpp header).  The oracle calls the
    host's expf; the device path restates the routine (csrc/lvs_math.cuh: glibc_expf).  The restatement must be the host routine
    bit for bit: 4M random arguments over the range the score can produce, a dense run of consecutive floats, the special cases.
    (tools/expf_sweep.c runs all 2.24e9 finite floats of [-104, 88.8].)"""
    rng = np.random.default_rng(11)
    x = np.concatenate([
        -rng.random(2_000_000, dtype=np.float32) * 104.0,
        -np.exp(rng.uniform(-40, 4.7, 1_000_000)).astype(np.float32),
        rng.random(500_000, dtype=np.float32) * 88.8,
        np.arange(0xC0000000, 0xC0000000 + 500_000, dtype=np.uint32).view(np.float32),
        np.array([0.0, -0.0, -87.3, -87.4, -103.27, -103.9, -103.98, -104.5, -1e30, -np.inf, 88.7, 88.73, 1e30, np.inf, np.nan,
                  -float.fromhex("0x1.f8cbb2p+5"), 1e-45, -1e-45], dtype=np.float32)])
    out, bad = O.expf_restated(x)
    assert bad == 0
    assert out[-4] != out[-4]                       # NaN in, NaN out
    fin = np.isfinite(x) & (np.abs(x) < 88)
    ref = np.exp(x[fin].astype(np.float64))
    ok = ref > 1e-37
    assert np.abs(out[fin][ok] / ref[ok] - 1).max() < 1.2e-7       # <= 1 ulp of float


# ---------------------------------------------------------------------------------------------------------------------
# pclomp_ground::NormalDistributionsTransformGround (include/ndt_omp/ndt_ground_impl.hpp): horizontal-voxel NDT for z / roll / pitch

_OFF7 = np.array([[0, 0, 0], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]])


def _last_neighbour_angle(o, pts, search):
    """numpy restatement of "the LAST cell of the point's neighbourhood decides": per point the angle2xy of the last usable leaf the
    direct search would push (probe order of getNeighborhoodAtPoint7 / 1), 0 for an empty neighbourhood (ndt_ground_impl.hpp:484)."""
    lv, ang = o.leaves(), o.leaf_angles()
    mn, mx, dv = o.grid()
    usable = {int(k): a for k, a, n in zip(lv["keys"], ang, lv["nr_points"]) if n >= 6}
    cell = np.floor(pts[:, :3].astype(np.float32) / np.float32(o.params["resolution"])).astype(np.int64)
    out = np.zeros(len(pts))
    for i, c in enumerate(cell):
        for off in (_OFF7 if search == O.DIRECT7 else _OFF7[:1]):
            q = c + off
            if (q < mn).any() or (q > mx).any():
                continue
            key = int((q[0] - mn[0]) + (q[1] - mn[1]) * dv[0] + (q[2] - mn[2]) * dv[0] * dv[1])
            if key in usable:
                out[i] = usable[key]
    return out


def test_ground_leaf_normals_against_numpy(small_pair):
    tgt = small_pair[0]
    o = O.OracleNDT(variant=O.VAR_GROUND, search=O.DIRECT1, num_threads=1)
    o.set_target(tgt)
    lv, ang = o.leaves(), o.leaf_angles()
    has = lv["in_cloud"] == 1
    assert (ang[~has] == -1).all() and has.sum() > 100
    checked = 0
    for k in np.nonzero(has & (lv["nr_points"] >= 6))[0]:
        w, v = np.linalg.eigh(lv["cov"][k])
        if w[1] - w[0] < 1e-6 * w[2]:
            continue                                        # the normal of a (near-)degenerate pair is not defined
        ref = np.degrees(np.arccos(min(1.0, abs(v[2, 0]))))
        assert abs(ref - ang[k]) < 1e-4, (k, ref, ang[k])   # 180/3.1415926 vs 180/pi: 2e-8 relative
        checked += 1
    assert checked > 100
    assert 0.05 < np.mean(ang[has] < 10) < 0.95             # the scene has both ground and wall leaves


@pytest.mark.parametrize("search", [O.DIRECT1, O.DIRECT7])
def test_ground_derivatives_are_the_gated_doubled_masked_omp_sums(small_pair, search):
    """computeDerivatives_seg with flag_class = 1 (ndt_ground_impl.hpp:363-572) against its definition in terms of the pclomp pass:
    points whose last neighbour is within 10 degrees of horizontal, score once, gradient and Hessian twice (updateDerivatives is called
    twice per cell, :519,522), rows / columns x, y, yaw zeroed (:554-561)."""
    tgt, src, guess, truth = small_pair
    og = O.OracleNDT(variant=O.VAR_GROUND, search=search, num_threads=1)
    oo = O.OracleNDT(variant=O.VAR_OMP, search=search, num_threads=1)
    og.set_target(tgt); oo.set_target(tgt)
    p = O.se3_log_from_matrix4f(guess)
    trans = O.transform(src, guess)
    keep = _last_neighbour_angle(og, trans, search) < 10
    assert 100 < keep.sum() < len(src) - 100
    og.set_source(src); oo.set_source(src[keep])
    sg, gg, Hg = og.eval_derivatives(p, guess)
    so, go, Ho = oo.eval_derivatives(p, guess)
    m = np.array([0, 0, 1, 1, 1, 0.0])
    np.testing.assert_allclose(sg, so, rtol=1e-13)
    np.testing.assert_allclose(gg, 2 * go * m, rtol=1e-12, atol=1e-12 * np.abs(go).max())
    np.testing.assert_allclose(Hg, 2 * Ho * np.outer(m, m), rtol=1e-12, atol=1e-12 * np.abs(Ho).max())
    assert (gg[[0, 1, 5]] == 0).all() and (Hg[[0, 1, 5], :] == 0).all() and (Hg[:, [0, 1, 5]] == 0).all()
    # the score-only pass (line search trials) gates the same points
    s2, g2, H2 = og.eval_derivatives(p, guess, False)
    assert s2 == sg and np.array_equal(g2, gg) and (H2 == 0).all()


def test_ground_align_solves_z_roll_pitch_only(small_pair):
    tgt = small_pair[0]
    for res, mot, tol_z, tol_r in ((2.0, [0, 0, 0.05, 0, 0, 0.0], 2e-3, 2e-4), (4.0, [0, 0, 0.12, 0.006, -0.004, 0.0], 0.01, 1e-3)):
        motion = O.se3_exp_matrix4f(np.array(mot))
        src = O.transform(tgt[::2], np.linalg.inv(motion.astype(np.float64)).astype(np.float32))
        o = O.OracleNDT(variant=O.VAR_GROUND, resolution=res, trans_eps=0.001, max_iter=64, search=O.DIRECT1, num_threads=os.cpu_count() or 1)
        o.set_target(tgt); o.set_source(src)
        r = o.align(np.eye(4, dtype=np.float32))
        assert r["converged"] and r["iterations"] < 64 and r["n_hess"] == 0 and r["n_eval"] == r["iterations"] + 1
        tr = r["trace"]
        assert (tr[:, 6 + 0] == 0).all() and (tr[:, 6 + 1] == 0).all() and (tr[:, 6 + 5] == 0).all()      # unit direction: no x, y, yaw component
        pf = O.se3_log_from_matrix4f(r["final"])
        assert abs(pf[2] - mot[2]) < tol_z and abs(pf[3] - mot[3]) < tol_r and abs(pf[4] - mot[4]) < tol_r, pf
    # the step test has no "not in the first iteration" clause (ndt_ground_impl.hpp:173): a start at the optimum ends after ONE iteration
    o2 = O.OracleNDT(variant=O.VAR_GROUND, resolution=2.0, trans_eps=0.05, max_iter=64, search=O.DIRECT1, num_threads=1)
    o2.set_target(tgt); o2.set_source(tgt[::2])
    r2 = o2.align(np.eye(4, dtype=np.float32))
    assert r2["converged"] and r2["iterations"] == 1
    o3 = O.OracleNDT(variant=O.VAR_OMP, resolution=2.0, trans_eps=0.05, max_iter=64, search=O.DIRECT1, num_threads=1)
    o3.set_target(tgt); o3.set_source(tgt[::2])
    assert o3.align(np.eye(4, dtype=np.float32))["iterations"] == 2


def test_golden_regression_of_the_ground_oracle(small_pair):
    """tests/golden/ndt_ground_small_pair.json (tests/golden/make_golden.py): ground_s2k's configuration on the small pair."""
    gold = json.load(open(os.path.join(HERE, "golden", "ndt_ground_small_pair.json")))
    tgt, src, guess, truth = small_pair
    o = O.OracleNDT(variant=O.VAR_GROUND, resolution=10.0, trans_eps=0.01, max_iter=64, search=O.DIRECT1, num_threads=1)
    o.set_target(tgt); o.set_source(src)
    ang = o.leaf_angles()
    hz = (ang >= 0) & (ang < 10)
    assert len(ang) == gold["n_cells"] and int((ang >= 0).sum()) == gold["n_with_normal"] and int(hz.sum()) == gold["n_horizontal"]
    assert int(o.leaves()["keys"][hz].astype(np.int64).sum()) == gold["horizontal_key_sum"]
    s, g, H = o.eval_derivatives(O.se3_log_from_matrix4f(guess), guess)
    np.testing.assert_allclose(s, gold["score"], rtol=1e-12)
    np.testing.assert_allclose(g, gold["gradient"], rtol=1e-10, atol=1e-9 * np.abs(g).max())
    np.testing.assert_allclose(H, np.array(gold["hessian"]), rtol=1e-10, atol=1e-9 * np.abs(H).max())
    r = o.align(guess)
    assert r["iterations"] == gold["iterations"]
    np.testing.assert_allclose(r["final"], np.array(gold["final"], dtype=np.float32), atol=1e-6)
