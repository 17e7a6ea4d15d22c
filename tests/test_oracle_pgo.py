"""CPU checks that pin the pose-graph oracle (no GPU).  g2o's own self-test for this path is a property test — analytic vs
numeric EdgeSE3 Jacobian on random poses, tolerance 1e-6 (3rdtools/g2o-a48ff8c.zip!g2o/g2o/types/slam3d/test_slam3d_jacobian.cpp:
109-140) — restated here, plus closed-form graphs and the cross-check of the reference's vendored CSparse against numpy."""
import json
import os

import numpy as np

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
