"""The CUDA path against the COMMITTED fixtures of tests/golden/ (written by tests/golden/make_golden.py from the CPU oracle; the
reference ships no golden vector of its own).  The other GPU tests compare with the oracle run live; this file
closes the triangle fixture <-> oracle (tests/test_oracle_*.py, CPU) <-> device, at the same tolerances.  It sorts last on
purpose: everything it checks is implied by tests that ran before it."""
import json
import os

import numpy as np
import pytest

import oracle_ndt as O
from lv_slam_b200.synth import posegraph as G

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _gold(name):
    return json.load(open(os.path.join(HERE, "golden", name)))


def _ndt(small_pair):
    import lv_slam_b200 as L
    tgt, src, guess, truth = small_pair
    n = L.NormalDistributionsTransform(variant=O.VAR_OMP)
    n.setTransformationEpsilon(0.01)
    n.setMaximumIterations(30)
    n.setNeighborhoodSearchMethod(O.DIRECT7)
    n.setResolution(1.0)
    n.setStepSize(0.1)
    n.setInputTarget(tgt)
    n.setInputSource(src)
    return n


def test_ndt_against_the_golden_pair(small_pair):
    gold = _gold("ndt_small_pair.json")
    tgt, src, guess, truth = small_pair
    n = _ndt(small_pair)
    c = n.cells()
    assert len(c["keys"]) == gold["n_cells"] and int((c["nr_points"] >= 6).sum()) == gold["n_usable"]      # bit-exact index work
    assert int(c["keys"].astype(np.int64).sum()) == gold["key_sum"]
    s, g, H = n.eval_derivatives(O.se3_log_from_matrix4f(guess), guess, True)
    assert abs(s - gold["score"]) <= 1e-9 * abs(gold["score"])
    gg, gH = np.array(gold["gradient"]), np.array(gold["hessian"])
    assert np.max(np.abs(np.asarray(g) - gg)) <= 1e-9 * np.max(np.abs(gg))
    assert np.max(np.abs(np.asarray(H) - gH)) <= 1e-9 * np.max(np.abs(gH))
    n.align(guess)
    r = n.result()
    assert r["iterations"] == gold["iterations"]
    fin = np.array(gold["final"])
    assert np.max(np.abs(r["final"][:3, 3] - fin[:3, 3])) <= 1e-4                  # north-star tolerance: 1e-4 m
    assert np.max(np.abs(np.asarray(r["final"], dtype=np.float64)[:3, :3] - fin[:3, :3])) <= 1e-5


def test_ground_ndt_against_the_golden_pair(small_pair):
    import lv_slam_b200 as L
    gold = _gold("ndt_ground_small_pair.json")
    tgt, src, guess, truth = small_pair
    n = L.NormalDistributionsTransformGround()          # ground_s2k, scan_matching_odom_nodelet.cpp:121-126
    n.setResolution(10.0); n.setNeighborhoodSearchMethod(O.DIRECT1); n.setTransformationEpsilon(0.01); n.setMaximumIterations(64)
    n.setInputTarget(tgt); n.setInputSource(src)
    hz = n.cell_horizontal() == 1
    assert len(hz) == gold["n_cells"] and int(hz.sum()) == gold["n_horizontal"]
    assert int(n.cells()["keys"][hz].astype(np.int64).sum()) == gold["horizontal_key_sum"]
    s, g, H = n.eval_derivatives(O.se3_log_from_matrix4f(guess), guess, True)
    gg, gH = np.array(gold["gradient"]), np.array(gold["hessian"])
    assert abs(s - gold["score"]) <= 1e-9 * abs(gold["score"])
    assert np.max(np.abs(np.asarray(g) - gg)) <= 1e-9 * np.max(np.abs(gg)) and np.max(np.abs(np.asarray(H) - gH)) <= 1e-9 * np.max(np.abs(gH))
    n.align(guess)
    r = n.result()
    fin = np.array(gold["final"])
    assert r["iterations"] == gold["iterations"]
    assert np.max(np.abs(r["final"][:3, 3] - fin[:3, 3])) <= 1e-4 and np.max(np.abs(np.asarray(r["final"], dtype=np.float64)[:3, :3] - fin[:3, :3])) <= 1e-5


def test_fitness_and_prefilter_against_the_golden_pair(small_pair):
    import lv_slam_b200 as L
    gold = _gold("aux_small_pair.json")
    tgt, src, guess, truth = small_pair
    n = _ndt(small_pair)
    big = float(np.finfo(np.float64).max)
    for name, T, mr in (("truth", truth, big), ("guess", guess, big), ("guess_capped", guess, 0.25)):
        sc, cnt = n.getFitnessScore(mr, T=T, with_count=True)
        assert cnt == gold["fitness"][name]["correspondences"]
        assert abs(sc - gold["fitness"][name]["score"]) <= 1e-11 * gold["fitness"][name]["score"]
    cloud = np.concatenate([tgt[:, :3], np.random.default_rng(3).random((len(tgt), 1), dtype=np.float32)], axis=1).astype(np.float32)
    pf = L.Prefilter(distance_near_thresh=0.5, distance_far_thresh=100.0, downsample_resolution=0.1).filter(cloud)
    g = gold["prefilter"]
    assert len(cloud) == g["n_in"] and len(pf) == g["n_out"]
    np.testing.assert_allclose(np.asarray(pf, dtype=np.float64).sum(axis=0), g["column_sums"], rtol=1e-12)
    assert np.array_equal(pf[0], np.array(g["first"], np.float32)) and np.array_equal(pf[-1], np.array(g["last"], np.float32))


def test_pose_graph_against_the_golden_sphere():
    import lv_slam_b200 as L
    gold = _gold("pgo_sphere_200.json")
    gr = G.sphere(20, 10, seed=7)
    assert len(gr["poses7"]) == gold["n_vertices"] and len(gr["ij"]) == gold["n_edges"]
    pg = L.PoseGraph(0)
    pg.set_graph(gr["poses7"], gr["ij"], gr["meas7"], gr["info21"], gr["huber"])
    e, c, tot = pg.errors()
    assert abs(tot - gold["robust_chi2_initial"]) <= 1e-9 * gold["robust_chi2_initial"]
    assert abs(float(np.sum(c)) - gold["chi2_initial"]) <= 1e-9 * gold["chi2_initial"]
    st = pg.optimize(100)
    assert abs(st["chi2_after"] - gold["chi2_final"]) <= 1e-6 * gold["chi2_final"]
    k = min(5, len(st["trace"]), len(gold["first_chi2"]))
    assert np.allclose(st["trace"][:k, 0], gold["first_chi2"][:k], rtol=1e-6)
    assert np.allclose(st["trace"][:k, 1], gold["first_lambdas"][:k], rtol=1e-6)


def test_unary_constraint_graph_against_the_golden_pin():
    """The sphere with GPS / IMU priors and the floor constraint (lvs_pgo_set_graph_typed): the device against the committed pin, at the
    tolerances of the live comparison (numeric Jacobians: tests/test_pgo_gpu.py)."""
    import lv_slam_b200 as L
    from test_oracle_pgo import _priors_on
    gold = _gold("pgo_sphere_200_unary.json")
    gr = G.sphere(20, 10, seed=7)
    ij, meas, info, hub, ty = _priors_on(gr, np.random.default_rng(5), 5)
    assert len(ij) == gold["n_edges"] and int((ty != 0).sum()) == gold["n_unary"]
    pg = L.PoseGraph(0)
    pg.set_graph(gr["poses7"], ij, meas, info, hub, None, ty, np.array(gold["floor_plane"]))
    e, c, tot = pg.errors()
    assert abs(tot - gold["robust_chi2_initial"]) <= 1e-9 * gold["robust_chi2_initial"]
    assert abs(float(np.sum(c)) - gold["chi2_initial"]) <= 1e-9 * gold["chi2_initial"]
    assert abs(float(np.abs(e[ty != 0]).sum()) - gold["unary_error_abs_sum"]) <= 1e-9 * gold["unary_error_abs_sum"]
    st = pg.optimize(100)
    assert abs(st["chi2_after"] - gold["chi2_final"]) <= 1e-5 * gold["chi2_final"]
    assert np.abs(pg.poses()[0][:3] - np.array(gold["pose_0"])[:3]).max() <= 1e-5 and np.abs(pg.poses()[-1][:3] - np.array(gold["pose_last"])[:3]).max() <= 1e-5
