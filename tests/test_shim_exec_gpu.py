"""The reference-side bindings of shim/ EXECUTED on the GPU: tests/shim_exec/shim_driver.cpp makes the calls a
lv_slam nodelet makes - pcl::Registration::setInputTarget / setInputSource / align / getFinalTransformation / hasConverged, the
fitness score, InformationMatrixCalculator::calc_fitness_score, the prefilter, GraphSLAM::optimize on a g2o graph - through the
interface stand-ins of tests/shim_stubs (PCL, Eigen and g2o are not in this image), and the results are compared with the CPU oracle."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle_ndt as O
import oracle_pgo as P

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rot_angle(Ra, Rb):
    d = Ra.astype(np.float64).T @ Rb.astype(np.float64)
    v = 0.5 * np.array([d[2, 1] - d[1, 2], d[0, 2] - d[2, 0], d[1, 0] - d[0, 1]])
    return float(np.arcsin(min(1.0, np.linalg.norm(v))))


def _build(tmp, variant):
    stubs = os.path.join(ROOT, "tests", "shim_stubs")
    exe = str(tmp / ("shim_driver_" + variant))
    cmd = ["g++", "-std=c++14", "-Wall", "-Wextra", "-Werror", "-O1"] + ({"pca": ["-DLVS_SHIM_PCA"], "ground": ["-DLVS_SHIM_GROUND"]}.get(variant, [])) + [
        "-I" + stubs, "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "shim"),
        os.path.join(ROOT, "tests", "shim_exec", "shim_driver.cpp"), os.path.join(ROOT, "shim", "graph_slam_b200.cpp"), os.path.join(ROOT, "shim", "aux_b200.cpp"),
        "-o", exe, "-L" + os.path.join(ROOT, "lv_slam_b200"), "-llvslam_b200", "-Wl,-rpath," + os.path.join(ROOT, "lv_slam_b200")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def _run(exe, d):
    r = subprocess.run([exe, str(d)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    out = {}
    for ln in r.stdout.splitlines():
        t = ln.split()
        if not t:
            continue
        if t[0] == "pose":
            out.setdefault("poses", []).append([float(x) for x in t[2:]])
        elif t[0] in ("iterations:", "chi2:", "time:"):          # GraphSLAM::optimize prints these like the reference does
            continue
        else:
            out[t[0]] = t[1:]
    return out


@pytest.mark.parametrize("variant", ["omp", "pca", "ground"])
def test_shims_run_on_the_gpu_and_match_the_oracle(tmp_path, small_pair, variant):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    from lv_slam_b200.synth import posegraph as G
    tgt, src, guess, truth = small_pair
    tgt.astype(np.float32).tofile(tmp_path / "tgt.f32"); src.astype(np.float32).tofile(tmp_path / "src.f32")
    np.ascontiguousarray(guess.T, dtype=np.float32).tofile(tmp_path / "guess.f32")                  # column-major, Eigen's order
    # a sphere graph with Huber kernels and the reference's GPS / IMU style unary priors (XY, XYZ, Quat, Vec) on every fifth vertex
    from test_oracle_pgo import _priors_on, FLOOR
    gr = G.sphere(20, 10, seed=7)
    p_ij, p_meas, p_info, p_hub, p_type = _priors_on(gr, np.random.default_rng(9), every=5)
    gr["poses7"].astype(np.float64).tofile(tmp_path / "poses7.f64"); p_meas.astype(np.float64).tofile(tmp_path / "meas7.f64")
    p_info.astype(np.float64).tofile(tmp_path / "info21.f64"); p_ij.astype(np.int32).tofile(tmp_path / "ij.i32")
    p_type.astype(np.int32).tofile(tmp_path / "etype.i32"); p_hub.astype(np.float64).tofile(tmp_path / "huber.f64")
    FLOOR.astype(np.float64).tofile(tmp_path / "floor.f64")
    out = _run(_build(tmp_path, variant), tmp_path)

    # ---- registration vs the CPU restatement with the same settings
    if variant == "ground":     # ground_s2k as scan_matching_odom_nodelet.cpp:121-126 configures it
        o = O.OracleNDT(variant=O.VAR_GROUND, resolution=10.0, trans_eps=0.01, max_iter=64, search=O.DIRECT1, num_threads=8)
    else:
        o = O.OracleNDT(variant=O.VAR_PCA if variant == "pca" else O.VAR_OMP, trans_eps=0.01, max_iter=30,
                        search=O.DIRECT1 if variant == "pca" else O.DIRECT7, num_threads=8)
    o.set_target(tgt); o.set_source(src)
    r = o.align(guess, want_cloud=True)
    F = np.array([float(x) for x in out["final"]], dtype=np.float32).reshape(4, 4).T
    assert int(out["iterations"][0]) == r["iterations"] and bool(int(out["converged"][0])) == r["converged"]
    assert np.max(np.abs(F[:3, 3] - r["final"][:3, 3])) <= 1e-4 and _rot_angle(F[:3, :3], r["final"][:3, :3]) <= 1e-5
    assert abs(float(out["trans_probability"][0]) - r["trans_probability"]) <= 1e-6 * abs(r["trans_probability"])
    fs, fc = o.fitness_score(r["final"], float(np.finfo(np.float64).max))
    assert abs(float(out["fitness"][0]) - fs) <= 1e-9 * fs
    fs2, _ = o.fitness_score(r["final"], 0.25)
    assert abs(float(out["fitness_capped"][0]) - fs2) <= 1e-9 * fs2
    assert int(out["aligned_n"][0]) == len(src)
    np.testing.assert_allclose([float(x) for x in out["aligned_first"]], r["cloud"][0], atol=2e-4)
    np.testing.assert_allclose([float(x) for x in out["aligned_last"]], r["cloud"][-1], atol=2e-4)
    assert int(out["second_iterations"][0]) == o.align(r["final"])["iterations"]

    if variant == "ground":
        return
    if variant == "pca":
        lv = o.leaves()
        c = out["cells"]              # "cells <n> usable <u> weight_sum <w>"
        assert int(c[0]) == len(lv["keys"]) and int(c[2]) == int((lv["nr_points"] >= 6).sum()) and int(c[4]) == int(lv["weight"].astype(np.int64).sum())
        return

    # ---- the stages either side of the path
    assert abs(float(out["info_fitness"][0]) - o.fitness_score(F, float(np.finfo(np.float64).max))[0]) <= 1e-9 * fs
    cloud = np.concatenate([src[:, :3], (np.arange(len(src)) % 7).astype(np.float32)[:, None]], axis=1).astype(np.float32)
    pf, _ = O.prefilter(cloud, 0.5, 100.0, True, 0.1)
    n_pf, sx, si = int(out["prefilter"][0]), float(out["prefilter"][1]), float(out["prefilter"][2])
    assert n_pf == len(pf) and abs(sx - pf[:, 0].astype(np.float64).sum()) <= 1e-3 and abs(si - pf[:, 3].astype(np.float64).sum()) <= 1e-3

    # ---- pose graph: GraphSLAM::optimize on the g2o containers vs the CPU restatement of g2o's LM
    op = P.OraclePGO()
    op.set_graph(gr["poses7"], p_ij, p_meas, p_info, p_hub, None, p_type, FLOOR)
    ro = op.optimize(100, P.ALG_LM, P.SOLVER_CSPARSE if P.have_csparse() else P.SOLVER_DENSE)
    assert int(out["pgo_iterations"][0]) > 0 and int(out["pgo_empty"][0]) == -1
    got = np.array(out["poses"])
    want = op.poses()
    dmax = float(np.abs(got[:, :3] - want[:, :3]).max())          # the priors fix the gauge: the poses themselves agree
    assert dmax <= 1e-5, dmax                                      # (numeric Jacobians of the prior edges: see tests/test_pgo_gpu.py)
    assert np.abs(np.abs(np.sum(got[:, 3:] * want[:, 3:], axis=1)) - 1.0).max() <= 1e-10
    assert ro["iterations"] > 0
