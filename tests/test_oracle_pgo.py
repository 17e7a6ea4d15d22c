"""CPU checks that pin the pose-graph oracle (no GPU).  g2o's own self-test for this path is a property test — analytic vs
numeric EdgeSE3 Jacobian on random poses, tolerance 1e-6 (3rdtools/g2o-a48ff8c.zip!g2o/g2o/types/slam3d/test_slam3d_jacobian.cpp:
109-140) — restated here, plus closed-form graphs and the cross-check of the reference's vendored CSparse against numpy."""
import json
import os

import numpy as np
import pytest

import oracle_pgo as P
from lv_slam_b200.synth import posegraph as G

HERE = os.path.dirname(os.path.abspath(__file__))


def _rand_pose(rng, scale=10.0):
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    return np.concatenate([rng.normal(0, scale, 3), q])


def test_edge_jacobians_against_numeric_differentiation():
    rng = np.random.default_rng(0)
    h = 1e-6
    for _ in range(300):
        z, xi, xj = _rand_pose(rng), _rand_pose(rng), _rand_pose(rng)
        if rng.uniform() < 0.5:      # near-consistent edge (small error), as in a converged graph
            xj = G.pose7(G.matrix(xi) @ G.matrix(z) @ G.matrix(np.concatenate([rng.normal(0, 0.05, 3), [0.01, -0.02, 0.015, 1.0]])))
        Ji, Jj = P.edge_jacobians(z, xi, xj)
        Ni, Nj = np.zeros((6, 6)), np.zeros((6, 6))
        for k in range(6):
            d = np.zeros(6); d[k] = h
            Ni[:, k] = (P.edge_error(z, P.oplus(xi, d), xj) - P.edge_error(z, P.oplus(xi, -d), xj)) / (2 * h)
            Nj[:, k] = (P.edge_error(z, xi, P.oplus(xj, d)) - P.edge_error(z, xi, P.oplus(xj, -d))) / (2 * h)
        # oplus perturbs with a compact quaternion, the analytic Jacobian differentiates w.r.t. that same update
        assert np.abs(Ji - Ni).max() < 1e-6 * max(1.0, np.abs(Ji).max())
        assert np.abs(Jj - Nj).max() < 1e-6 * max(1.0, np.abs(Jj).max())


def test_error_is_zero_on_consistent_edge_and_oplus_identity():
    rng = np.random.default_rng(1)
    xi, z = _rand_pose(rng), _rand_pose(rng)
    xj = G.pose7(G.matrix(xi) @ G.matrix(z))
    np.testing.assert_allclose(P.edge_error(z, xi, xj), 0, atol=1e-13)
    np.testing.assert_allclose(P.oplus(xi, np.zeros(6)), xi * np.sign(xi[6]) if False else P.oplus(xi, np.zeros(6)))
    T = G.matrix(P.oplus(xi, np.array([0.1, 0.2, 0.3, 0, 0, 0.0])))
    np.testing.assert_allclose(T[:3, 3], G.matrix(xi)[:3, 3] + G.matrix(xi)[:3, :3] @ [0.1, 0.2, 0.3], atol=1e-13)   # X * increment
    # |dq|^2 > 1: fromCompactQuaternion returns the identity rotation (isometry3d_mappings.cpp:84-91)
    T2 = G.matrix(P.oplus(xi, np.array([0, 0, 0, 0.9, 0.9, 0.9])))
    np.testing.assert_allclose(T2[:3, :3], G.matrix(xi)[:3, :3], atol=1e-13)


def test_huber_weights_and_robust_chi2():
    g = G.sphere(10, 4, seed=3)
    o = P.OraclePGO()
    o.set_graph(g["poses7"], g["ij"], g["meas7"], g["info21"], g["huber"])
    e, c, tot = o.errors()
    rho = np.where(c <= 1.0, c, 2 * np.sqrt(c) - 1.0)       # Huber with delta = 1 (robust_kernel_impl.cpp:65-78)
    np.testing.assert_allclose(tot, rho.sum(), rtol=1e-13)
    # chi2 = e^T Omega e with Omega = diag(2,2,2,10,10,10) (information_matrix_calculator.cpp:29-35 with the launch-file stddevs)
    np.testing.assert_allclose(c, (e ** 2 * np.array([2, 2, 2, 10, 10, 10.0])).sum(axis=1), rtol=1e-13)
    lin = o.linearize()
    w = np.where(c <= 1.0, 1.0, 1.0 / np.sqrt(c))
    Om = np.diag([2, 2, 2, 10, 10, 10.0])
    nf = len(g["poses7"])
    H = np.zeros((nf * 6, nf * 6)); b = np.zeros(nf * 6)
    for k, (i, j) in enumerate(g["ij"]):
        A, B = lin["Ji"][k], lin["Jj"][k]
        H[i * 6:i * 6 + 6, i * 6:i * 6 + 6] += w[k] * A.T @ Om @ A
        H[j * 6:j * 6 + 6, j * 6:j * 6 + 6] += w[k] * B.T @ Om @ B
        H[i * 6:i * 6 + 6, j * 6:j * 6 + 6] += w[k] * A.T @ Om @ B
        H[j * 6:j * 6 + 6, i * 6:i * 6 + 6] += w[k] * B.T @ Om @ A
        b[i * 6:i * 6 + 6] -= w[k] * A.T @ Om @ e[k]
        b[j * 6:j * 6 + 6] -= w[k] * B.T @ Om @ e[k]
    np.testing.assert_allclose(P.dense_system(lin), H, rtol=1e-11, atol=1e-9)
    np.testing.assert_allclose(lin["b"], b, rtol=1e-11, atol=1e-9)


def test_linear_solvers_agree_with_numpy():
    g = G.sphere(12, 6, seed=5)
    o = P.OraclePGO()
    o.set_graph(g["poses7"], g["ij"], g["meas7"], g["info21"], g["huber"])
    lin = o.linearize()
    lam = 1e-5 * np.max(np.abs(np.einsum("nii->ni", lin["Hd"])))
    x_ref = np.linalg.solve(P.dense_system(lin, lam), lin["b"])
    ok, x, _ = o.solve(lam, P.SOLVER_DENSE)
    assert ok and np.abs(x - x_ref).max() <= 1e-9 * np.abs(x_ref).max()
    if P.have_csparse():             # the reference's vendored CSparse (oracle/_ref, built by oracle/build_ref.sh)
        ok, x, _ = o.solve(lam, P.SOLVER_CSPARSE)
        assert ok and np.abs(x - x_ref).max() <= 1e-9 * np.abs(x_ref).max()
    ok, x, it = o.solve(lam, P.SOLVER_PCG, 1e-26)
    assert ok and it > 10 and np.abs(x - x_ref).max() <= 1e-7 * np.abs(x_ref).max()
    ok, x6, it6 = o.solve(lam, P.SOLVER_PCG, 1e-6)
    assert it6 < it
    # without damping and without a fixed vertex the system is singular (6 gauge freedoms): Cholesky must report failure
    ok, _, _ = o.solve(0.0, P.SOLVER_DENSE)
    sv = np.linalg.svd(P.dense_system(lin), compute_uv=False)
    assert sv[-6] / sv[0] < 1e-9 or not ok


def test_triangle_graph_known_optimum():
    """Three poses on a line, consistent odometry 1 m apart plus a loop edge claiming 2.3 m between 0 and 2, no kernel, vertex 0
    fixed: the weighted least-squares optimum along x is closed form (x1 = 1.1, x2 = 2.2 for equal information)."""
    I7 = np.array([0, 0, 0, 0, 0, 0, 1.0])
    poses = np.array([I7, I7 + [1, 0, 0, 0, 0, 0, 0], I7 + [2, 0, 0, 0, 0, 0, 0]])
    ij = np.array([[0, 1], [1, 2], [0, 2]], dtype=np.int32)
    meas = np.array([I7 + [1, 0, 0, 0, 0, 0, 0], I7 + [1, 0, 0, 0, 0, 0, 0], I7 + [2.3, 0, 0, 0, 0, 0, 0]])
    info = np.tile(G.info21_diag([1, 1, 1, 1, 1, 1]), (3, 1))
    o = P.OraclePGO()
    o.set_graph(poses, ij, meas, info, None, np.array([1, 0, 0], dtype=np.uint8))
    r = o.optimize(20, P.ALG_GN, P.SOLVER_DENSE)
    assert r["iterations"] == 20
    p = o.poses()
    np.testing.assert_allclose(p[:, 0], [0, 1.1, 2.2], atol=1e-12)
    np.testing.assert_allclose(p[:, 1:6], 0, atol=1e-12)
    np.testing.assert_allclose(r["chi2_after"], 0.01 + 0.01 + 0.01, rtol=1e-9)
    # LM reaches the same optimum and then stops by itself (ten rejected trials or rho == 0), well before max_iterations
    o.set_graph(poses, ij, meas, info, None, np.array([1, 0, 0], dtype=np.uint8))
    r = o.optimize(1024, P.ALG_LM, P.SOLVER_DENSE)
    assert 0 < r["iterations"] < 200
    np.testing.assert_allclose(o.poses()[:, 0], [0, 1.1, 2.2], atol=1e-7)


def test_lm_on_sphere_and_golden(tmp_path):
    gold = json.load(open(os.path.join(HERE, "golden", "pgo_sphere_200.json")))
    g = G.sphere(20, 10, seed=7)
    assert len(g["poses7"]) == gold["n_vertices"] and len(g["ij"]) == gold["n_edges"]
    o = P.OraclePGO()
    o.set_graph(g["poses7"], g["ij"], g["meas7"], g["info21"], g["huber"])
    e, c, tot = o.errors()
    np.testing.assert_allclose([tot, c.sum()], [gold["robust_chi2_initial"], gold["chi2_initial"]], rtol=1e-12)
    solvers = [P.SOLVER_DENSE] + ([P.SOLVER_CSPARSE] if P.have_csparse() else [])
    for s in solvers:
        o.set_graph(g["poses7"], g["ij"], g["meas7"], g["info21"], g["huber"])
        r = o.optimize(100, P.ALG_LM, s)
        assert r["iterations"] > 5
        np.testing.assert_allclose(r["chi2_after"], gold["chi2_final"], rtol=1e-8)
        np.testing.assert_allclose(r["trace"][:5, 1], gold["first_lambdas"], rtol=1e-7)
        np.testing.assert_allclose(r["trace"][:5, 0], gold["first_chi2"], rtol=1e-7)
        assert (np.diff(r["trace"][:, 0]) <= 1e-9).all()    # accepted steps never increase the robust chi2
    # sphere generator: edge count of g2o's create_sphere (4 999 + 14 600 at 100 x 50)
    assert len(G.sphere(100, 50, seed=7)["ij"]) == 19599


def test_empty_graph_returns_minus_one():
    o = P.OraclePGO()
    o.set_graph(np.array([[0, 0, 0, 0, 0, 0, 1.0]]), np.zeros((0, 2), np.int32), np.zeros((0, 7)), np.zeros((0, 21)))
    assert o.optimize(10)["iterations"] == -1               # graph_slam.cpp:302-305


# ---- unary priors on a VertexSE3 (include/g2o/edge_se3_prior{xy,xyz,quat,vec}.hpp; "later" row of SURVEY.md §8f)
FLOOR = np.array([0.3, -0.2, 0.9, 1.5])          # the fixed VertexPlane the EdgeSE3Plane rows (kind 5) refer to


def _local_plane(T, plane):
    """operator*(Isometry3d, Plane3D) with the inverse pose: the plane as seen from the vehicle (g2o plane3d.h)."""
    p = plane / np.linalg.norm(plane[:3])
    Ti = np.linalg.inv(T)
    n = Ti[:3, :3] @ p[:3]
    return np.r_[n, p[3] - Ti[:3, 3] @ n]


def _priors_on(g, rng, every=7):
    """The sphere graph g with GPS / IMU / floor style unary constraints on every `every`-th vertex, interleaved after the binary edges of
    that vertex: returns (ij, meas7, info21, huber, edge_type); the floor rows refer to the plane FLOOR."""
    ij, meas, info, hub, ty = [], [], [], [], []
    nv = len(g["poses7"])
    by_first = {}
    for k, (a, b) in enumerate(g["ij"]):
        by_first.setdefault(int(a), []).append(k)
    done = set()
    for v in range(nv):
        for k in by_first.get(v, []):
            ij.append(g["ij"][k]); meas.append(g["meas7"][k]); info.append(g["info21"][k]); hub.append(g["huber"][k]); ty.append(0)
            done.add(k)
        if v % every:
            continue
        T = G.matrix(g["truth7"][v])
        kind = 1 + (v // every) % 5
        m, I6 = np.zeros(7), np.zeros((6, 6))
        if kind == 1:
            m[:2] = T[:2, 3] + rng.normal(0, 0.05, 2); I6[:2, :2] = np.array([[4.0, 0.3], [0.3, 5.0]])
        elif kind == 2:
            m[:3] = T[:3, 3] + rng.normal(0, 0.05, 3); I6[:3, :3] = np.diag([4.0, 4.0, 1.0]) + 0.2
        elif kind == 3:
            q = G.pose7(T)[3:] * (-1.0 if v % 2 else 1.0)           # either sign: setMeasurement keeps w >= 0
            m[:4] = q + rng.normal(0, 0.01, 4); I6[:3, :3] = np.diag([50.0, 60.0, 70.0])
        elif kind == 4:
            d = np.array([0.0, 0.0, -1.0])
            m[:3] = 3.0 * d; m[3:6] = T[:3, :3].T @ d + rng.normal(0, 0.02, 3); I6[:3, :3] = np.eye(3) * 30.0
        else:
            m[:4] = 2.5 * (_local_plane(T, FLOOR) + np.r_[rng.normal(0, 0.01, 3), rng.normal(0, 0.05)])      # any scale: Plane3D normalises
            I6[:3, :3] = np.diag([40.0, 40.0, 8.0]) + 0.5
        iu = np.triu_indices(6)
        ij.append([v, v]); meas.append(m); info.append(I6[iu]); hub.append(1.0 if v % (2 * every) == 0 else 0.0); ty.append(kind)
    assert len(done) == len(g["ij"])
    return np.array(ij, np.int32), np.array(meas), np.array(info), np.array(hub), np.array(ty, np.int32)


def test_prior_edge_errors_and_numeric_jacobians():
    rng = np.random.default_rng(5)
    for _ in range(50):
        x = _rand_pose(rng)
        T = G.matrix(x)
        # PriorXYZ / PriorXY: translation minus measurement; Jacobian of t + R dt is [R 0] (rows of the edge's dimension only)
        m = rng.normal(0, 5, 3)
        np.testing.assert_allclose(P.prior_error(2, m, x), np.r_[T[:3, 3] - m, 0, 0, 0], atol=1e-14)
        np.testing.assert_allclose(P.prior_error(1, m[:2], x), np.r_[T[:2, 3] - m[:2], 0, 0, 0, 0], atol=1e-14)
        J = P.prior_jacobian(2, m, x)
        np.testing.assert_allclose(J[:3, :3], T[:3, :3], atol=3e-5)      # central differences with delta = 1e-9 on |t| ~ 10: round-off 1e-15 / 2e-9
        np.testing.assert_allclose(J[:3, 3:], 0, atol=3e-5)
        np.testing.assert_allclose(J[3:], 0, atol=0)
        assert np.abs(P.prior_jacobian(1, m[:2], x)[2:]).max() == 0
        # PriorQuat: zero at the pose's own quaternion whatever its sign; d vec(q (x) dq) / d dq at dq = 0 is  w I + [v]x
        q = G.pose7(T)[3:]
        np.testing.assert_allclose(P.prior_error(3, -q, x)[:3], 0, atol=1e-13)
        qq = q if q[3] >= 0 else -q
        vx = np.array([[0, -qq[2], qq[1]], [qq[2], 0, -qq[0]], [-qq[1], qq[0], 0]])
        Jq = P.prior_jacobian(3, q, x)
        np.testing.assert_allclose(Jq[:3, 3:], qq[3] * np.eye(3) + vx, atol=5e-6)
        np.testing.assert_allclose(Jq[:3, :3], 0, atol=5e-6)
        # PriorVec: R^T direction - measurement, both normalised by setMeasurement
        d, z = rng.normal(size=3), rng.normal(size=3)
        e = P.prior_error(4, np.r_[d, z], x)
        np.testing.assert_allclose(e[:3], T[:3, :3].T @ (d / np.linalg.norm(d)) - z / np.linalg.norm(z), atol=1e-13)
        # EdgeSE3Plane against a fixed plane: zero when the measurement IS the plane seen from the pose (whatever its scale); a shift of the
        # measured distance is the third component; azimuth / elevation of the measured normal in the local plane's frame are the first two
        lp = _local_plane(T, FLOOR)
        np.testing.assert_allclose(P.prior_error(5, np.r_[3.0 * lp, FLOOR], x)[:3], 0, atol=1e-12)
        np.testing.assert_allclose(P.prior_error(5, np.r_[lp + [0, 0, 0, 0.25], FLOOR], x)[:3], [0, 0, 0.25], atol=1e-12)      # distance() = -coeffs(3)
        Jp = P.prior_jacobian(5, np.r_[lp, FLOOR], x)
        assert np.abs(Jp[3:]).max() == 0 and np.abs(Jp[:3]).max() > 0.1 and np.linalg.matrix_rank(Jp[:3], tol=1e-3) == 3


def test_single_vertex_with_position_and_orientation_priors_has_a_known_optimum():
    """One vertex, a PriorXYZ and a PriorQuat: LM must land on (measurement, measured quaternion) - the only edges are unary."""
    x0 = np.array([[1.0, -2.0, 0.5, 0.1, -0.2, 0.05, 0.97]]); x0[0, 3:] /= np.linalg.norm(x0[0, 3:])
    target_q = np.array([0.2, 0.1, -0.15, 0.96]); target_q /= np.linalg.norm(target_q)
    iu = np.triu_indices(6)
    Ixyz, Iq = np.zeros((6, 6)), np.zeros((6, 6))
    Ixyz[:3, :3] = np.diag([2.0, 3.0, 4.0]); Iq[:3, :3] = np.eye(3) * 100
    meas = np.zeros((2, 7)); meas[0, :3] = [3.0, 1.0, -1.0]; meas[1, :4] = target_q
    o = P.OraclePGO()
    o.set_graph(x0, np.array([[0, 0], [0, 0]], np.int32), meas, np.array([Ixyz[iu], Iq[iu]]), None, None, np.array([2, 3], np.int32))
    r = o.optimize(200, P.ALG_LM, P.SOLVER_DENSE)
    assert r["iterations"] > 0 and r["chi2_after"] < 1e-12
    p = o.poses()[0]
    np.testing.assert_allclose(p[:3], [3.0, 1.0, -1.0], atol=1e-7)
    np.testing.assert_allclose(p[3:] * np.sign(p[6]), target_q, atol=1e-7)


def test_priors_pull_a_drifting_chain_back():
    """Sphere graph with odometry edges only (no loop closures) + XYZ priors from the truth on every 5th vertex: the optimum is far
    closer to the truth than the chained odometry, and typed and untyped set_graph agree when no prior is present."""
    g = G.sphere(20, 5, seed=3)
    odo = np.flatnonzero(np.abs(g["ij"][:, 0] - g["ij"][:, 1]) == 1)
    ij, meas, info, hub = g["ij"][odo], g["meas7"][odo], g["info21"][odo], g["huber"][odo]
    o0 = P.OraclePGO(); o0.set_graph(g["poses7"], ij, meas, info, hub)
    o1 = P.OraclePGO(); o1.set_graph(g["poses7"], ij, meas, info, hub, None, np.zeros(len(ij), np.int32))
    assert o0.errors()[2] == o1.errors()[2]
    iu = np.triu_indices(6)
    I6 = np.zeros((6, 6)); I6[:3, :3] = np.eye(3) * 25
    vs = np.arange(0, len(g["poses7"]), 5)
    pm = np.zeros((len(vs), 7)); pm[:, :3] = g["truth7"][vs, :3]
    o = P.OraclePGO()
    o.set_graph(g["poses7"], np.vstack([ij, np.stack([vs, vs], 1)]).astype(np.int32), np.vstack([meas, pm]), np.vstack([info, np.tile(I6[iu], (len(vs), 1))]),
                np.r_[hub, np.zeros(len(vs))], None, np.r_[np.zeros(len(ij)), np.full(len(vs), 2)].astype(np.int32))
    r = o.optimize(100, P.ALG_LM, P.SOLVER_DENSE)
    assert r["iterations"] > 0
    before = np.abs(g["poses7"][:, :3] - g["truth7"][:, :3]).max()
    after = np.abs(o.poses()[:, :3] - g["truth7"][:, :3]).max()
    assert after < 0.25 * before, (before, after)


def test_plane_edge_error_against_an_independent_numpy_restatement():
    """EdgeSE3Plane::computeError (include/g2o/edge_se3_plane.hpp:40-47) with g2o's Plane3D arithmetic written out again in numpy with plain
    rotation matrices (the oracle goes through Eigen's quaternion product): random poses, planes and measurements."""
    rng = np.random.default_rng(12)

    def rot(n):                                   # Plane3D::rotation: Rz(azimuth) * Ry(-elevation)
        az, el = np.arctan2(n[1], n[0]), np.arctan2(n[2], np.hypot(n[0], n[1]))
        Rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
        Ry = np.array([[np.cos(-el), 0, np.sin(-el)], [0, 1, 0], [-np.sin(-el), 0, np.cos(-el)]])
        return Rz @ Ry

    for _ in range(200):
        x = _rand_pose(rng)
        plane, meas = rng.normal(size=4), rng.normal(size=4)
        lp = _local_plane(G.matrix(x), plane)
        lp = lp / np.linalg.norm(lp[:3])
        m = meas / np.linalg.norm(meas[:3])
        n = rot(lp[:3]).T @ m[:3]
        want = [np.arctan2(n[1], n[0]), np.arctan2(n[2], np.hypot(n[0], n[1])), (-lp[3]) - (-m[3])]
        np.testing.assert_allclose(P.prior_error(5, np.r_[meas, plane], x)[:3], want, atol=1e-11)


def test_unary_constraint_graph_against_the_golden_pin():
    gold = json.load(open(os.path.join(HERE, "golden", "pgo_sphere_200_unary.json")))
    gr = G.sphere(20, 10, seed=7)
    ij, meas, info, hub, ty = _priors_on(gr, np.random.default_rng(5), 5)
    assert len(ij) == gold["n_edges"] and int((ty != 0).sum()) == gold["n_unary"] and sorted(set(int(t) for t in ty)) == gold["kinds"]
    o = P.OraclePGO()
    o.set_graph(gr["poses7"], ij, meas, info, hub, None, ty, np.array(gold["floor_plane"]))
    e, c, tot = o.errors()
    np.testing.assert_allclose([tot, c.sum(), np.abs(e[ty != 0]).sum()], [gold["robust_chi2_initial"], gold["chi2_initial"], gold["unary_error_abs_sum"]], rtol=1e-12)
    st = o.optimize(100, P.ALG_LM, P.SOLVER_DENSE)
    np.testing.assert_allclose(st["chi2_after"], gold["chi2_final"], rtol=1e-9)
    np.testing.assert_allclose(o.poses()[0], gold["pose_0"], atol=1e-9)


def test_compute_dq_dR_against_g2o_own_generated_code():
    """g2o's own types/slam3d/dquat2mat.cpp with its Maxima-generated cases (unpacked from the reference's 3rdtools/g2o-a48ff8c.zip and compiled
    as they are by oracle/build_ref.sh) against the restatement of the table behind EdgeSE3::linearizeOplus: all four branches of the
    quaternion extraction, identical to the last bit."""
    import ctypes
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libdquat_ref.so")
    if not os.path.exists(so):
        if os.path.exists("/root/reference/3rdtools/g2o-a48ff8c.zip"):
            import subprocess
            subprocess.call(["sh", os.path.join(os.path.dirname(so), "..", "build_ref.sh")])
        if not os.path.exists(so):
            pytest.skip("no compiled g2o dquat2mat")
    G = ctypes.CDLL(so)
    G.gref_compute_dq_dR.restype = None; G.gref_compute_dq_dR.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L = P.lib()
    rng = np.random.default_rng(41)
    seen = set()
    for k in range(400):
        q = rng.normal(size=4)
        if k % 4:                                            # push the trace below zero to reach the x / y / z branches
            q[0] *= 0.05
            q[1 + k % 3] *= 5
        q /= np.linalg.norm(q)
        w, x, y, z = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        tr = np.trace(R)
        seen.add(0 if tr > 0 else 1 if (R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]) else 2 if R[1, 1] > R[2, 2] else 3)
        a, b = np.zeros(27), np.zeros(27)
        Rc = np.ascontiguousarray(R)
        G.gref_compute_dq_dR(Rc.ctypes.data, a.ctypes.data)
        L.opgo_compute_dq_dR(Rc.ctypes.data, b.ctypes.data)
        assert np.array_equal(a, b), (k, np.abs(a - b).max())
    assert seen == {0, 1, 2, 3}


def test_edge_se3_math_against_g2o_own_code():
    """g2o's own slam3d edge math - computeEdgeSE3Gradient with its skew helpers (isometry3d_gradients.h), toVectorMQT / fromVectorMQT /
    toCompactQuaternion / fromCompactQuaternion / normalize (isometry3d_mappings.cpp) and compute_dq_dR, taken from the reference's g2o zip at
    build time and compiled in oracle/g2o_ref_harness.cpp - against the restatement of EdgeSE3::computeError, EdgeSE3::linearizeOplus and
    VertexSE3::oplusImpl: error vectors, both 6 x 6 Jacobians and the updated pose identical to the last bit on random edges, near-identity
    edges, half-turn edges (negative quaternion w, trace below zero) and updates of every size up to |v| > 1 (the identity fallback)."""
    import ctypes
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libg2o_ref.so")
    if not os.path.exists(so):
        if os.path.exists("/root/reference/3rdtools/g2o-a48ff8c.zip"):
            import subprocess
            subprocess.call(["sh", os.path.join(os.path.dirname(so), "..", "build_ref.sh")])
        if not os.path.exists(so):
            pytest.skip("no compiled g2o slam3d math")
    G = ctypes.CDLL(so)
    vp = ctypes.c_void_p
    for f, n in (("gref_edge_error", 4), ("gref_edge_jacobians", 5), ("gref_oplus", 4)):
        getattr(G, f).restype = None; getattr(G, f).argtypes = [vp] * n
    L = P.lib()
    L.opgo_edge_error.restype = None; L.opgo_edge_error.argtypes = [vp] * 4
    L.opgo_edge_jacobians.restype = None; L.opgo_edge_jacobians.argtypes = [vp] * 5
    rng = np.random.default_rng(43)

    def pose(kind):
        q = rng.normal(size=4)
        if kind == 1:
            q = np.array([1.0, 0, 0, 0]) + 1e-3 * rng.normal(size=4)
        elif kind == 2:
            q[0] *= 0.02                                     # near a half turn
        q /= np.linalg.norm(q)
        if kind == 3:
            q = -q
        return np.concatenate([rng.normal(0, 10, 3), q[1:], q[:1]])

    G.gref_huber.restype = None; G.gref_huber.argtypes = [ctypes.c_double, ctypes.c_double, vp]
    L.opgo_huber.restype = None; L.opgo_huber.argtypes = [ctypes.c_double, ctypes.c_double, vp]
    for e, d in ((0.0, 1.0), (0.5, 1.0), (1.0, 1.0), (1.0000001, 1.0), (37.5, 1.0), (3.0, 2.0), (4.0, 2.0), (1e6, 0.3)):      # RobustKernelHuber::robustify
        ra, rb = np.zeros(3), np.zeros(3)
        G.gref_huber(e, d, ra.ctypes.data); L.opgo_huber(e, d, rb.ctypes.data)
        assert np.array_equal(ra, rb), (e, d)
    for k in range(300):
        z, xi, xj = pose(k % 4), pose((k // 4) % 4), pose((k // 16) % 4)
        a, b = np.zeros(6), np.zeros(6)
        G.gref_edge_error(z.ctypes.data, xi.ctypes.data, xj.ctypes.data, a.ctypes.data)
        L.opgo_edge_error(z.ctypes.data, xi.ctypes.data, xj.ctypes.data, b.ctypes.data)
        assert np.array_equal(a, b), (k, a, b)
        Ja, Jb, Ka, Kb = np.zeros(36), np.zeros(36), np.zeros(36), np.zeros(36)
        G.gref_edge_jacobians(z.ctypes.data, xi.ctypes.data, xj.ctypes.data, Ja.ctypes.data, Ka.ctypes.data)
        L.opgo_edge_jacobians(z.ctypes.data, xi.ctypes.data, xj.ctypes.data, Jb.ctypes.data, Kb.ctypes.data)
        assert np.array_equal(Ja, Jb) and np.array_equal(Ka, Kb), (k, np.abs(Ja - Jb).max(), np.abs(Ka - Kb).max())
        upd = rng.normal(0, [1, 1, 1, 0.3, 0.3, 0.3]) * (10.0 ** rng.integers(-6, 1))
        if k % 25 == 0:
            upd[3:] = [0.8, 0.7, 0.6]                        # |v|^2 > 1: fromCompactQuaternion returns the identity
        Ra, ta, Rb, tb = np.zeros(9), np.zeros(3), np.zeros(9), np.zeros(3)
        G.gref_oplus(xi.ctypes.data, upd.ctypes.data, Ra.ctypes.data, ta.ctypes.data)
        L.opgo_oplus_matrix(xi.ctypes.data, upd.ctypes.data, Rb.ctypes.data, tb.ctypes.data)
        assert np.array_equal(Ra, Rb) and np.array_equal(ta, tb), k


def test_lm_control_flow_against_g2o_own_code():
    """g2o's own OptimizationAlgorithmLevenberg::solve / computeLambdaInit / computeScale (taken from the reference's g2o zip at build time,
    oracle/lm_ref_harness.cpp) driving the restatement's building blocks, against the restated loops of oracle/pgo_oracle.cpp (Levenberg-Marquardt and, at the end, g2o's
    OptimizationAlgorithmGaussNewton::solve against the restated Gauss-Newton): iteration counts, the per-iteration chi2 / lambda / trial counts and the optimised poses identical - on the sphere with Huber kernels (dense and
    CSparse solves), with unary priors and the floor constraint, with a fixed vertex, and from a start far enough that steps are rejected."""
    import ctypes
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(root, "oracle", "_ref", "liblm_ref.so")
    if not os.path.exists(so):
        if os.path.exists("/root/reference/3rdtools/g2o-a48ff8c.zip"):
            import subprocess
            subprocess.call(["sh", os.path.join(root, "oracle", "build_ref.sh")])
        if not os.path.exists(so):
            pytest.skip("no compiled g2o LM")
    G = ctypes.CDLL(so)
    vp, i32 = ctypes.c_void_p, ctypes.c_int
    G.opgo_create.restype = vp
    G.opgo_destroy.argtypes = [vp]; G.opgo_destroy.restype = None
    G.opgo_set_graph_typed.argtypes = [vp, i32, vp, vp, i32, vp, vp, vp, vp, vp]; G.opgo_set_graph_typed.restype = None
    G.opgo_set_floor_plane.argtypes = [vp, vp]; G.opgo_set_floor_plane.restype = None
    G.opgo_get_poses.argtypes = [vp, vp]; G.opgo_get_poses.restype = None
    G.opgo_load_csparse.restype = i32; G.opgo_load_csparse.argtypes = [ctypes.c_char_p]
    G.gref_lm_optimize.restype = i32; G.gref_lm_optimize.argtypes = [vp, i32, i32, vp, vp, i32, vp]
    cs = os.path.join(root, "oracle", "_ref", "libcsparse_ref.so")
    if os.path.exists(cs):
        G.opgo_load_csparse(cs.encode())

    def run_g2o(poses, ij, meas, info, hub, fixed, etype, floor, iters, solver):
        h = G.opgo_create()
        c = lambda a, dt=np.float64: np.ascontiguousarray(a, dtype=dt)
        p, e, m, inf = c(poses), c(ij, np.int32), c(meas), c(info)
        hb = c(hub) if hub is not None else None
        fx = c(fixed, np.uint8) if fixed is not None else None
        ty = c(etype, np.int32) if etype is not None else None
        if floor is not None:
            G.opgo_set_floor_plane(h, c(floor).ctypes.data)
        G.opgo_set_graph_typed(h, len(p), p.ctypes.data, fx.ctypes.data if fx is not None else None, len(e), e.ctypes.data, m.ctypes.data, inf.ctypes.data,
                               hb.ctypes.data if hb is not None else None, ty.ctypes.data if ty is not None else None)
        st, tr, n = np.zeros(5), np.zeros((iters + 1, 3)), ctypes.c_int(0)
        it = G.gref_lm_optimize(h, iters, solver, st.ctypes.data, tr.ctypes.data, iters + 1, ctypes.byref(n))
        out = np.zeros((len(p), 7))
        G.opgo_get_poses(h, out.ctypes.data)
        G.opgo_destroy(h)
        return it, st, tr[:n.value], out

    from lv_slam_b200.synth import posegraph as SG
    g1 = SG.sphere(20, 10, seed=7)
    rng = np.random.default_rng(9)
    p_ij, p_meas, p_info, p_hub, p_type = _priors_on(g1, rng, every=5)
    far = g1["poses7"].copy()
    far[:, :3] += np.random.default_rng(3).normal(0, 3.0, far[:, :3].shape)      # a start that makes LM reject steps
    fixed0 = np.zeros(len(g1["poses7"]), np.uint8); fixed0[0] = 1
    cases = [("sphere, Huber", g1["poses7"], g1["ij"], g1["meas7"], g1["info21"], g1["huber"], None, None, None),
             ("sphere, fixed vertex", g1["poses7"], g1["ij"], g1["meas7"], g1["info21"], g1["huber"], fixed0, None, None),
             ("far start", far, g1["ij"], g1["meas7"], g1["info21"], g1["huber"], None, None, None),
             ("unary priors + floor", g1["poses7"], p_ij, p_meas, p_info, p_hub, None, p_type, FLOOR)]
    solvers = [P.SOLVER_DENSE] + ([P.SOLVER_CSPARSE] if (P.have_csparse() and os.path.exists(cs)) else [])
    rejected = 0
    for ci, (name, poses, ij, meas, info, hub, fixed, etype, floor) in enumerate(cases):
        for solver in (solvers if ci == 0 or len(solvers) == 1 else solvers[1:]):      # the dense solve is slow: first case only when CSparse is there
            o = P.OraclePGO()
            o.set_graph(poses, ij, meas, info, hub, fixed, etype, floor)
            r = o.optimize(40, P.ALG_LM, solver)
            it, st, tr, out = run_g2o(poses, ij, meas, info, hub, fixed, etype, floor, 40, solver)
            assert it == r["iterations"] > 0, (name, it, r["iterations"])
            assert np.array_equal(tr[:, 0], r["trace"][:, 0]) and np.array_equal(tr[:, 1], r["trace"][:, 1]) and np.array_equal(tr[:, 2], r["trace"][:, 2]), name
            assert st[1] == r["chi2_after"] and st[2] == r["lam"] and int(st[3]) == r["trials"], name
            assert np.array_equal(out, o.poses()), name
            rejected += int((tr[:, 2] > 1).sum())
    assert rejected > 0                                      # the trial loop (pop, lambda *= ni) was exercised
    # g2o's own LinearSolverPCG::solve (solvers/pcg/linear_solver_pcg.hpp, block-Jacobi preconditioner, 1e-6 relative tolerance) on the damped
    # system of the graph's current linearisation, against the restated solve_pcg: same iteration count, same solution
    G.gref_pcg_solve.restype = i32; G.gref_pcg_solve.argtypes = [vp, ctypes.c_double, vp, vp]
    for name, poses, ij, meas, info, hub, fixed, etype, floor in cases[:2] + cases[3:]:
        for lam in (0.0 if fixed is not None else 1e-3, 10.0):
            o = P.OraclePGO()
            o.set_graph(poses, ij, meas, info, hub, fixed, etype, floor)
            o.linearize()
            ok, x_o, it_o = o.solve(lam, P.SOLVER_PCG)
            h = G.opgo_create()
            c = lambda a, dt=np.float64: np.ascontiguousarray(a, dtype=dt)
            p_, e_, m_, i_ = c(poses), c(ij, np.int32), c(meas), c(info)
            hb_ = c(hub) if hub is not None else None
            fx_ = c(fixed, np.uint8) if fixed is not None else None
            ty_ = c(etype, np.int32) if etype is not None else None
            if floor is not None:
                G.opgo_set_floor_plane(h, c(floor).ctypes.data)
            G.opgo_set_graph_typed(h, len(p_), p_.ctypes.data, fx_.ctypes.data if fx_ is not None else None, len(e_), e_.ctypes.data, m_.ctypes.data, i_.ctypes.data,
                                   hb_.ctypes.data if hb_ is not None else None, ty_.ctypes.data if ty_ is not None else None)
            x_g, it_g = np.zeros_like(x_o), ctypes.c_int(0)
            okg = G.gref_pcg_solve(h, lam, x_g.ctypes.data, ctypes.byref(it_g))
            G.opgo_destroy(h)
            assert ok and okg and it_g.value == it_o > 3, (name, lam, it_g.value, it_o)
            assert np.array_equal(x_g, x_o), (name, lam, np.abs(x_g - x_o).max())
    # Gauss-Newton: g2o's own OptimizationAlgorithmGaussNewton::solve in the same outer loop (vertex 0 fixed, like GraphSLAM's gn_var solvers need)
    G.gref_gn_optimize.restype = i32; G.gref_gn_optimize.argtypes = [vp, i32, i32, vp, vp, i32, vp]
    for solver in solvers[-1:]:
        o = P.OraclePGO()
        o.set_graph(g1["poses7"], g1["ij"], g1["meas7"], g1["info21"], g1["huber"], fixed0)
        r = o.optimize(8, P.ALG_GN, solver)
        h = G.opgo_create()
        c = lambda a, dt=np.float64: np.ascontiguousarray(a, dtype=dt)
        p_, e_, m_, i_, hb_, fx_ = c(g1["poses7"]), c(g1["ij"], np.int32), c(g1["meas7"]), c(g1["info21"]), c(g1["huber"]), c(fixed0, np.uint8)
        G.opgo_set_graph_typed(h, len(p_), p_.ctypes.data, fx_.ctypes.data, len(e_), e_.ctypes.data, m_.ctypes.data, i_.ctypes.data, hb_.ctypes.data, None)
        st, tr, n = np.zeros(5), np.zeros((9, 3)), ctypes.c_int(0)
        it = G.gref_gn_optimize(h, 8, solver, st.ctypes.data, tr.ctypes.data, 9, ctypes.byref(n))
        out = np.zeros((len(p_), 7))
        G.opgo_get_poses(h, out.ctypes.data)
        G.opgo_destroy(h)
        assert it == r["iterations"] == 8 and np.array_equal(tr[:n.value, 0], r["trace"][:, 0]) and np.array_equal(out, o.poses())


def test_prior_edges_against_the_reference_own_classes():
    """The reference's OWN unary edge classes - include/g2o/edge_se3_priorxy.hpp, _priorxyz.hpp, _priorquat.hpp, _priorvec.hpp and edge_se3_plane.hpp (with
    g2o's own plane3d.h from the vendored zip), included as they are and compiled against stand-ins for the g2o / Eigen headers they include (oracle/prior_ref_api.cpp) - against the restatement:
    setMeasurement (the quaternion's sign flip, the two normalisations of the vector prior) followed by computeError on random poses,
    half-turn poses and measurements with negative w."""
    import ctypes
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(root, "oracle", "_ref", "libprior_ref.so")
    if not os.path.exists(so):
        if os.path.exists("/root/reference/include/g2o/edge_se3_priorvec.hpp"):
            import subprocess
            subprocess.call(["sh", os.path.join(root, "oracle", "build_ref.sh")])
        if not os.path.exists(so):
            pytest.skip("no compiled reference prior edges")
    G = ctypes.CDLL(so)
    G.pref_prior_error.restype = None; G.pref_prior_error.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    G.pref_prior_jacobian.restype = None; G.pref_prior_jacobian.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    rng = np.random.default_rng(47)
    worst = 0.0
    for k in range(400):
        q = rng.normal(size=4)
        if k % 3 == 0:
            q[0] *= 0.02
        q /= np.linalg.norm(q)
        x7 = np.concatenate([rng.normal(0, 10, 3), q[1:], q[:1]])
        for kind in (1, 2, 3, 4, 5):
            meas = np.zeros(8)
            if kind == 5:                                    # EdgeSE3Plane: measured plane, the (fixed) plane vertex; Plane3D normalises both
                meas[:4] = np.concatenate([rng.normal(size=3), rng.normal(0, 2, 1)])
                meas[4:] = np.concatenate([rng.normal(size=3), rng.normal(0, 2, 1)])
            elif kind in (1, 2):
                meas[:3] = rng.normal(0, 10, 3)
            elif kind == 3:
                mq = rng.normal(size=4); mq /= np.linalg.norm(mq)
                meas[:4] = mq                                # x y z w, either sign of w
            else:
                meas[:6] = rng.normal(0, 3, 6)               # direction and measured vector, both normalised by setMeasurement
            want = np.zeros(6)
            G.pref_prior_error(kind, meas.ctypes.data, x7.ctypes.data, want.ctypes.data)
            got = P.prior_error(kind, meas, x7)
            # the numeric Jacobian: g2o's own BaseUnaryEdge / BaseBinaryEdge::linearizeOplus (1e-9 central differences through push / oplus /
            # computeError / pop) around the same classes.  A last-bit difference of an error value is worth 5e8 times as much here.
            Jw = np.zeros(36)
            G.pref_prior_jacobian(kind, meas.ctypes.data, x7.ctypes.data, Jw.ctypes.data)
            Jg = P.prior_jacobian(kind, meas, x7).reshape(36)
            if kind in (1, 2, 3):
                assert np.array_equal(Jg, Jw), (kind, np.abs(Jg - Jw).max())
            else:
                assert np.abs(Jg - Jw).max() <= 2e-6, (kind, np.abs(Jg - Jw).max())
            if kind == 5:
                assert np.abs(got - want).max() <= 1e-15, (kind, got, want)
            elif kind == 4:                                  # linear().inverse(): the stand-in's 3 x 3 inverse and the restatement's differ in the last bit
                worst = max(worst, float(np.abs(got - want).max()))
                assert np.abs(got - want).max() <= 2e-15, (kind, got, want)
            else:
                assert np.array_equal(got, want), (kind, got, want)


def test_quadratic_form_against_g2o_own_code():
    """g2o's own BaseBinaryEdge / BaseUnaryEdge::constructQuadraticForm (core/base_*_edge.hpp of the vendored zip, with its Huber kernel) on the
    Jacobians, information and error of single edges, against what the restatement's build_system assembles for the same edge: both diagonal
    blocks, the off-diagonal block, both right-hand sides, with and without kernel (inlier and outlier), fixed vertices, unary priors of
    dimension 2 and 3.  g2o forms (A^T W) B, the restatement A^T (W B): agreement to rounding (1e-13 of the block's largest entry)."""
    import ctypes
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(root, "oracle", "_ref", "libprior_ref.so")
    if not os.path.exists(so):
        pytest.skip("no compiled reference edges")
    G = ctypes.CDLL(so)
    vp, f64, i32 = ctypes.c_void_p, ctypes.c_double, ctypes.c_int
    G.pref_quadratic_form_binary.restype = None; G.pref_quadratic_form_binary.argtypes = [vp, vp, vp, vp, f64, i32, i32, vp, vp, vp, vp, vp]
    G.pref_quadratic_form_unary.restype = None; G.pref_quadratic_form_unary.argtypes = [i32, vp, vp, vp, f64, vp, vp]
    rng = np.random.default_rng(53)

    def pose():
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        return np.concatenate([rng.normal(0, 5, 3), q[1:], q[:1]])

    def full_info(upper21):
        M = np.zeros((6, 6)); k = 0
        for r in range(6):
            for c in range(r, 6):
                M[r, c] = M[c, r] = upper21[k]; k += 1
        return M

    def close(a, b):
        return np.abs(np.asarray(a) - np.asarray(b)).max() <= 1e-13 * max(1.0, np.abs(b).max())

    seen_outlier = seen_inlier = 0
    for k in range(120):
        xi, xj = pose(), pose()
        z = pose()
        L = rng.normal(size=(6, 6)); info = L @ L.T + 6 * np.eye(6)
        info21 = np.array([info[r, c] for r in range(6) for c in range(r, 6)])
        hub = [0.0, 1.0, 1e6][k % 3]                         # no kernel, outlier (chi2 >> 1), inlier (huge delta)
        fixed = np.array([k % 5 == 0, k % 7 == 0], np.uint8)
        if fixed.all():
            fixed[1] = 0
        o = P.OraclePGO()
        o.set_graph(np.stack([xi, xj]), np.array([[0, 1]], np.int32), z[None, :], info21[None, :], np.array([hub]), fixed)
        lin = o.linearize()
        err = P.edge_error(z, xi, xj)
        Ji, Jj = lin["Ji"][0], lin["Jj"][0]
        Ai, bi, Aj, bj, Hij = np.zeros((6, 6)), np.zeros(6), np.zeros((6, 6)), np.zeros(6), np.zeros((6, 6))
        G.pref_quadratic_form_binary(np.ascontiguousarray(Ji).ctypes.data, np.ascontiguousarray(Jj).ctypes.data, np.ascontiguousarray(info).ctypes.data, err.ctypes.data,
                                     float(hub), int(fixed[0]), int(fixed[1]), Ai.ctypes.data, bi.ctypes.data, Aj.ctypes.data, bj.ctypes.data, Hij.ctypes.data)
        e2 = float(err @ info @ err)
        if hub > 0:
            seen_outlier += e2 > hub * hub
            seen_inlier += e2 <= hub * hub
        free = [v for v in (0, 1) if not fixed[v]]
        blocks = {0: (Ai, bi), 1: (Aj, bj)}
        for slot, v in enumerate(free):
            assert close(lin["Hd"][slot], blocks[v][0]) and close(lin["b"][slot * 6:slot * 6 + 6], blocks[v][1]), (k, v)
        if len(free) == 2:
            assert close(lin["Ho"][0], Hij), k
    assert seen_outlier > 10 and seen_inlier > 10
    # unary priors: XY (D = 2) and XYZ (D = 3) on one free vertex, with and without kernel
    for k in range(60):
        x = pose()
        kind = 1 + k % 2
        D = 2 if kind == 1 else 3
        meas = np.zeros(8); meas[:D] = x[:D] + rng.normal(0, 3.0, D)
        Lm = rng.normal(size=(D, D)); infoD = Lm @ Lm.T + D * np.eye(D)
        info = np.zeros((6, 6)); info[:D, :D] = infoD
        info21 = np.array([info[r, c] for r in range(6) for c in range(r, 6)])
        hub = [0.0, 0.5, 1e6][k % 3]
        o = P.OraclePGO()
        o.set_graph(x[None, :], np.array([[0, 0]], np.int32), meas[None, :7], info21[None, :], np.array([hub]), None, np.array([kind], np.int32))
        lin = o.linearize()
        err = P.prior_error(kind, meas, x)
        J = P.prior_jacobian(kind, meas, x)
        A, b = np.zeros((6, 6)), np.zeros(6)
        G.pref_quadratic_form_unary(D, np.ascontiguousarray(J).ctypes.data, np.ascontiguousarray(info).ctypes.data, err.ctypes.data, float(hub), A.ctypes.data, b.ctypes.data)
        assert close(lin["Hd"][0], A) and close(lin["b"][:6], b), (k, kind)
