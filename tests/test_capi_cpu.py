"""No-GPU checks of the boundary: the shared library loads, exports every symbol include/lvslam_b200.h declares, reports
errors the documented way, and fails loudly (LVS_ERR_NO_DEVICE) instead of computing anything when no B200 is present.
Also the host-side logic that needs no device: g2o text save/load, keyframe planning, multi-rank plumbing over gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "lvslam_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lvs_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from lv_slam_b200 import _capi, build
    build.build()
    L = ctypes.CDLL(_capi.LIB_PATH)
    names = _declared_functions()
    assert len(names) >= 40
    for n in names:
        assert hasattr(L, n), "missing export " + n
    out = subprocess.check_output(["nm", "-D", "--defined-only", _capi.LIB_PATH]).decode()
    exported = set(re.findall(r" T (\w+)", out))
    assert set(names) <= exported
    # nothing but the C-ABI leaks out of the library
    assert all(s.startswith("lvs_") for s in exported), sorted(s for s in exported if not s.startswith("lvs_"))[:5]


def test_library_carries_sm100a_code_only():
    from lv_slam_b200 import _capi
    out = subprocess.run(["cuobjdump", "--list-elf", _capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_defaults_and_error_strings():
    from lv_slam_b200 import _capi as C
    L = C.lib()
    p = C.NdtParams()
    L.lvs_ndt_default_params(ctypes.byref(p))
    # constructor defaults of the reference class (ndt_omp_impl2.hpp:54-83)
    assert (p.resolution, p.step_size, p.outlier_ratio, p.transformation_epsilon, p.max_iterations, p.search_method) == (1.0, 0.1, 0.55, 0.1, 35, C.LVS_DIRECT7)
    assert p.min_points_per_voxel == 6 and p.min_covar_eigvalue_mult == 0.01
    assert L.lvs_status_string(0) == b"ok" and b"device" in L.lvs_status_string(-2)


@pytest.mark.skipif("__import__('torch').cuda.is_available()")
def test_no_cpu_fallback_without_a_device():
    import lv_slam_b200 as M
    with pytest.raises(M.LvsError) as e:
        M.NormalDistributionsTransform()
    assert e.value.status == -2
    with pytest.raises(M.LvsError) as e:
        M.NdtBatch(1, 1)
    assert e.value.status == -2
    with pytest.raises(M.LvsError) as e:
        M.PoseGraph()
    assert e.value.status == -2
    assert M.lib().lvs_device_count() == 0


def test_invalid_arguments_are_rejected_before_touching_the_device():
    from lv_slam_b200 import _capi as C
    L = C.lib()
    h = ctypes.c_void_p()
    p = C.NdtParams()
    L.lvs_ndt_default_params(ctypes.byref(p))
    p.resolution = -1.0
    assert L.lvs_ndt_create(ctypes.byref(p), 0, None, ctypes.byref(h)) == -1
    assert b"resolution" in L.lvs_last_error()
    p.resolution = 1.0; p.search_method = 9
    assert L.lvs_ndt_create(ctypes.byref(p), 0, None, ctypes.byref(h)) == -1
    assert L.lvs_pgo_create(17, 0, None, ctypes.byref(h)) == -1
    assert L.lvs_ndt_align(None, None, None) == -1 and L.lvs_pgo_optimize(None, 1, None) == -1
    assert L.lvs_ndt_destroy(None) == 0 and L.lvs_pgo_destroy(None) == 0


def test_graph_slam_text_round_trip(tmp_path):
    import lv_slam_b200 as M
    from lv_slam_b200.synth import posegraph as G
    g = G.sphere(8, 3, seed=2)
    gs = M.GraphSLAM("lm_var_cholmod")
    vs = [gs.add_se3_node(G.matrix(p)) for p in g["poses7"]]
    vs[0].setFixed(True)
    for k, ((a, b), m) in enumerate(zip(g["ij"], g["meas7"])):
        e = gs.add_se3_edge(vs[a], vs[b], G.matrix(m), np.diag([2, 2, 2, 10, 10, 10.0]))
        gs.add_robust_kernel(e, "Huber" if k % 2 == 0 else "NONE", 1.0)
    path = str(tmp_path / "g.g2o")
    assert gs.save(path)
    txt = open(path).read().splitlines()
    assert txt[0].startswith("VERTEX_SE3:QUAT 0 ") and sum(l.startswith("EDGE_SE3:QUAT") for l in txt) == len(g["ij"])
    assert len(txt[-1].split()) == 3 + 7 + 21
    gs2 = M.GraphSLAM()
    assert gs2.load(path)
    assert gs2.num_vertices() == gs.num_vertices() and gs2.num_edges() == gs.num_edges()
    p1, f1, ij1, m1, i1, h1, _ = gs._arrays()
    p2, f2, ij2, m2, i2, h2, _ = gs2._arrays()
    np.testing.assert_allclose(p2, p1, atol=1e-15); np.testing.assert_allclose(m2, m1, atol=1e-15)
    assert np.array_equal(ij1, ij2) and np.array_equal(f1, f2) and np.array_equal(h1, h2) and np.array_equal(i1, i2)
    assert M.GraphSLAM().optimize(5) == -1                  # no edges (graph_slam.cpp:302-305), decided before any device call
    # the robust-kernel sidecar has the reference's own line format "<n vertices> <ids...> <type> <delta>" (robust_kernel_io.cpp:22-62)
    kl = open(path + ".kernels").read().splitlines()
    assert len(kl) == (len(g["ij"]) + 1) // 2 and all(len(l.split()) == 5 and l.split()[0] == "2" and l.split()[3] == "Huber" for l in kl)
    # and a file written by the reference (g2o text + sidecar) loads
    ref = str(tmp_path / "ref.g2o")
    open(ref, "w").write("VERTEX_SE3:QUAT 0 0 0 0 0 0 0 1\nVERTEX_SE3:QUAT 1 1 0 0 0 0 0 1\nEDGE_SE3:QUAT 1 0 -1 0 0 0 0 0 1 " +
                         " ".join("2" if k in (0, 6, 11) else "10" if k in (15, 18, 20) else "0" for k in range(21)) + "\n")
    open(ref + ".kernels", "w").write("2 1 0 Huber 1\n")
    gs3 = M.GraphSLAM()
    assert gs3.load(ref) and gs3.num_edges() == 1 and gs3._arrays()[5][0] == 1.0 and np.array_equal(gs3._arrays()[2], [[1, 0]])
    # a foreign file whose SE3 ids are not 0..n-1 (a dump that also held a floor-plane node with id 2), with two parallel edges
    # of which only the first has a kernel record: rows follow the vertex list, kernels are consumed one per edge, new ids do not collide
    gap = str(tmp_path / "gap.g2o")
    info = " ".join("2" if k in (0, 6, 11) else "10" if k in (15, 18, 20) else "0" for k in range(21))
    open(gap, "w").write("VERTEX_SE3:QUAT 0 0 0 0 0 0 0 1\nVERTEX_SE3:QUAT 1 1 0 0 0 0 0 1\nVERTEX_SE3:QUAT 3 2 0 0 0 0 0 1\nVERTEX_SE3:QUAT 4 3 0 0 0 0 0 1\n"
                         "EDGE_SE3:QUAT 3 1 -1 0 0 0 0 0 1 " + info + "\nEDGE_SE3:QUAT 4 3 -1 0 0 0 0 0 1 " + info + "\nEDGE_SE3:QUAT 4 3 -1 0 0 0 0 0 1 " + info + "\n")
    open(gap + ".kernels", "w").write("2 4 3 Huber 0.5\n")
    gs4 = M.GraphSLAM()
    assert gs4.load(gap)
    p4, f4, ij4, m4, i4, h4, _ = gs4._arrays()
    assert np.array_equal(ij4, [[2, 1], [3, 2], [3, 2]]) and p4[2, 0] == 2.0 and p4[3, 0] == 3.0
    assert h4.tolist() == [0.0, 0.5, 0.0]
    assert gs4.add_se3_node(np.eye(4)).id() == 5
    # KITTI pose lines (scan_matching_odom_nodelet.cpp:157-160)
    from lv_slam_b200.graph_slam import load_kitti_poses, save_kitti_poses
    Ts = [G.matrix(p) for p in g["poses7"][:5]]
    save_kitti_poses(str(tmp_path / "odom.txt"), Ts)
    back = load_kitti_poses(str(tmp_path / "odom.txt"))
    assert len(back) == 5 and len(open(str(tmp_path / "odom.txt")).readline().split()) == 12
    np.testing.assert_allclose(np.array(back), np.array(Ts), rtol=0, atol=1e-6 * 200)


def test_keyframe_plan_follows_the_nodelet_policy():
    from lv_slam_b200 import synth
    poses = [synth.pose_matrix(synth.traj_pose(f)) for f in range(40)]
    plan = synth.keyframe_plan(poses)
    assert [f for f, _, _ in plan] == list(range(2, 40))
    keys = [k for _, k, _ in plan]
    assert keys[0] == 0 and sorted(set(keys)) == [0, 9, 18, 27, 36]      # 1.2 m per frame: a new keyframe once 10 m are exceeded
    for f, k, g in plan:
        truth = np.linalg.inv(poses[k]) @ poses[f]
        assert np.abs(g[:3, 3] - truth[:3, 3]).max() < 0.05              # constant-velocity guess


_WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import torch.distributed as dist
from lv_slam_b200 import dist as D
rank, local, world = D.env_rank()
dist.init_process_group("gloo")
start, span = D.frame_range(rank, 6)
mx = D.max_over_ranks(10.0 + rank, world)
sm = D.sum_over_ranks(span - 2, world)
import torch
lo = [None] * world
dist.all_gather_object(lo, (start, span))
# point-sharding plumbing: the 64-byte IPC handles travel in rank order, the source chunks tile the cloud
blobs = D.all_gather_bytes(bytes([rank]) * 64, world)
chunks = [D.shard_range(125001, r, world) for r in range(world)]
if rank == 0:
    print("RESULT", mx, sm, lo)
    print("SHARD", [b[0] for b in blobs], [len(b) for b in blobs], chunks)
dist.destroy_process_group()
"""


def test_two_rank_plumbing_over_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29517", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][0]
    assert "11.0 12.0 [(0, 8), (8, 8)]" in line            # max over ranks, total pairs, disjoint frame ranges
    shard = [l for l in out.stdout.splitlines() if l.startswith("SHARD")][0]
    assert "[0, 1] [64, 64] [(0, 62500), (62500, 125001)]" in shard


def test_graph_slam_prior_edges_round_trip(tmp_path):
    """The reference's GPS / IMU prior edges (graph_slam.cpp:194-240) in the host mirror: setMeasurement semantics, the flat arrays
    lvs_pgo_set_graph_typed takes, and the g2o text tags the reference registers (graph_slam.cpp:31-35).  No device needed."""
    from lv_slam_b200.graph_slam import GraphSLAM, PRIOR_QUAT, PRIOR_VEC
    gs = GraphSLAM("lm_var")
    T = np.eye(4); T[:3, 3] = [1, 2, 3]
    a, b = gs.add_se3_node(np.eye(4)), gs.add_se3_node(T)
    gs.add_se3_edge(a, b, T, np.eye(6))
    e1 = gs.add_se3_prior_xy_edge(b, [1.5, 2.5], np.diag([4.0, 5.0]))
    gs.add_se3_prior_xyz_edge(a, [0.1, 0.2, 0.3], np.eye(3) * 2)
    e3 = gs.add_se3_prior_quat_edge(b, [0.5, -0.5, -0.5, -0.5], np.eye(3) * 10)
    e4 = gs.add_se3_prior_vec_edge(a, [0, 0, -2.0], [0.0, 3.0, -4.0], np.eye(3))
    gs.add_robust_kernel(e1, "Huber", 1.5)
    floor = gs.add_plane_node([0.0, 0.0, 2.0, 0.0])                      # the floor node of the nodelet: id 2, normalised, fixed
    e5 = gs.add_se3_plane_edge(b, floor, [0.0, 0.0, 1.0, -1.7], np.eye(3) * 5)
    with pytest.raises(NotImplementedError):
        gs._floor_plane()                                                # a FREE plane vertex is not on this path
    floor.setFixed(True)
    assert floor.id() == 2 and np.allclose(gs._floor_plane(), [0, 0, 1, 0]) and gs.add_se3_node(np.eye(4)).id() == 3
    assert e3.kind == PRIOR_QUAT and np.allclose(e3.measurement, [-0.5, 0.5, 0.5, 0.5])                       # w >= 0
    assert e4.kind == PRIOR_VEC and np.allclose(e4.measurement, [0, 0, -1, 0, 0.6, -0.8])                # both halves normalised
    poses, fixed, ij, meas, info, hub, types = gs._arrays()
    assert types.tolist() == [0, 1, 2, 3, 4, 5] and ij.tolist() == [[0, 1], [1, 1], [0, 0], [1, 1], [0, 0], [1, 1]] and hub.tolist() == [0, 1.5, 0, 0, 0, 0]
    assert np.allclose(meas[5, :4], [0, 0, 1, -1.7]) and info[5, 0] == 5.0
    assert np.allclose(meas[1, :2], [1.5, 2.5]) and np.allclose(meas[3, :4], [-0.5, 0.5, 0.5, 0.5]) and info[1, 0] == 4.0 and info[1, 6] == 5.0 and info[1, 11] == 0.0
    f = str(tmp_path / "g.g2o")
    assert gs.save(f)
    text = open(f).read()
    assert "EDGE_SE3_PRIORXY 1 1.5 2.5 4.0 0.0 5.0" in text and "EDGE_SE3_PRIORQUAT 1 0.5 -0.5 0.5 0.5" in text and "EDGE_SE3_PRIORVEC 0" in text
    assert "VERTEX_PLANE 2 0.0 0.0 1.0 0.0 0 0 0" in text and "FIX 2" in text and "EDGE_SE3_PLANE 1 2 0.0 0.0 1.0 -1.7 5.0 0.0 0.0 5.0 0.0 5.0" in text
    gs2 = GraphSLAM("lm_var")
    assert gs2.load(f) and gs2.num_edges() == 6 and np.allclose(gs2._floor_plane(), [0, 0, 1, 0])
    p2 = gs2._arrays()
    for x, y in zip(gs._arrays(), p2):
        assert np.allclose(x, y)


def _elimination_game(n, off, order):
    """Independent symbolic Cholesky: eliminate in `order`, connecting the remaining neighbours of every pivot."""
    pos = np.empty(n, np.int64)
    pos[order] = np.arange(n)
    adj = [set() for _ in range(n)]
    for a, b in off:
        adj[a].add(b); adj[b].add(a)
    nnz = n
    for p in order:
        later = [v for v in adj[p] if pos[v] > pos[p]]
        nnz += len(later)
        for v in later:
            adj[v].discard(p)
            adj[v].update(u for u in later if u != v)
    return nnz


def test_direct_solver_symbolic_analysis_on_the_host():
    """lvs_pgo_chol_analyze (minimum-degree ordering, elimination tree, supernodes) needs no device: the fill it reports is the
    fill of an independent elimination game under its ordering, and it beats the natural ordering on the sphere graph."""
    from lv_slam_b200.graph_slam import chol_analyze
    from lv_slam_b200.synth import posegraph as G
    g = G.sphere(20, 10, seed=7)
    n = len(g["poses7"])
    off = np.array(sorted({(min(a, b), max(a, b)) for a, b in np.asarray(g["ij"]) if a != b}), np.int32)
    st, perm = chol_analyze(n, off)
    assert sorted(perm.tolist()) == list(range(n))
    assert st["nnz_l_blocks"] == _elimination_game(n, off, perm)
    assert st["nnz_l_blocks"] < _elimination_game(n, off, np.arange(n))
    assert 1 <= st["levels"] <= st["fronts"] <= n and st["max_front"] % 6 == 1 and st["arena_bytes"] > 0
    # a chain (pure odometry) has no fill at all, and an empty pattern is n independent 1x1 fronts
    chain = np.array([(i, i + 1) for i in range(49)], np.int32)
    st, perm = chol_analyze(50, chain)
    assert st["nnz_l_blocks"] == 50 + 49
    st, perm = chol_analyze(7, np.zeros((0, 2), np.int32))
    assert st["nnz_l_blocks"] == 7 and st["fronts"] == 7 and st["levels"] == 1
    with pytest.raises(Exception):
        chol_analyze(5, np.array([(3, 1)], np.int32))     # not (row < col)


def test_relaxed_supernodes_keep_the_fill_and_shorten_the_tree():
    """Relaxed amalgamation (a chain child joins its parent when that pads its columns with few explicit zeros) must not change the
    ordering or the fill that is reported - only the fronts: fewer of them, fewer levels, some more arithmetic (host only)."""
    from lv_slam_b200.synth import posegraph as G
    g = G.sphere(50, 20, seed=11)
    n = len(g["poses7"])
    off = np.array(sorted({(min(a, b), max(a, b)) for a, b in np.asarray(g["ij"]) if a != b}), np.int32)
    code = ("import sys, json, numpy as np; sys.path.insert(0, %r); from lv_slam_b200.graph_slam import chol_analyze; "
            "off = np.load(sys.argv[1]); st, perm = chol_analyze(int(sys.argv[2]), off); print(json.dumps([st, perm.tolist()]))" % ROOT)
    import json, tempfile
    with tempfile.TemporaryDirectory() as d:
        np.save(os.path.join(d, "off.npy"), off)
        res = {}
        for relax in ("0", "12"):
            out = subprocess.run([sys.executable, "-c", code, os.path.join(d, "off.npy"), str(n)], capture_output=True, text=True,
                                 env=dict(os.environ, LVS_CHOL_RELAX=relax), timeout=300)
            assert out.returncode == 0, out.stderr[-1500:]
            res[relax] = json.loads(out.stdout.strip().splitlines()[-1])
    (s0, p0), (s1, p1) = res["0"], res["12"]
    assert p0 == p1 and s0["nnz_l_blocks"] == s1["nnz_l_blocks"] == _elimination_game(n, off, np.array(p0))
    assert s1["fronts"] < s0["fronts"] and s1["levels"] <= s0["levels"] and s0["factor_fma"] <= s1["factor_fma"] <= 1.1 * s0["factor_fma"]
    assert s1["max_front"] == s0["max_front"]


def test_host_marshalling_helpers():
    """CloudBatch / pack_guesses / AlignResults: the per-call host work of the batched API is a C call over prepared arrays, and the
    result records are numpy views of the C structs (no device needed)."""
    from lv_slam_b200 import _capi as C
    from lv_slam_b200.ndt import AlignResults, CloudBatch, pack_guesses
    clouds = [np.zeros((5, 3), np.float32), np.ones((7, 3), np.float32)]
    cb = CloudBatch(clouds)
    assert cb.n == 2 and cb.stride == 12 and cb.on_device == 0 and list(cb.counts) == [5, 7] and cb.ptrs[1] == clouds[1].ctypes.data
    wide = np.zeros((4, 8), np.float32)
    assert CloudBatch([wide[:, :3]]).stride == 32
    with pytest.raises(ValueError):
        CloudBatch([clouds[0], wide[:, :3]])
    T = np.arange(16, dtype=np.float32).reshape(4, 4)
    g = pack_guesses([T, np.eye(4)])
    assert g.shape == (2, 16) and g.dtype == np.float32 and g[0, 1] == T[1, 0] and g[0, 12] == T[0, 3]      # column-major
    raw = (C.NdtResult * 2)()
    for i in range(2):
        raw[i].iterations, raw[i].n_eval, raw[i].converged, raw[i].score = 5 + i, 1 + i, i, 1.5 * i
        for k in range(16):
            raw[i].final_transformation[k] = k + 100 * i
    r = AlignResults(raw, 2)
    assert len(r) == 2 and r[1]["iterations"] == 6 and r[1]["final"][0, 3] == 112.0 and r[0]["converged"] is False
    assert r.n_eval.tolist() == [1, 2] and r.finals.shape == (2, 4, 4) and r.finals[1][3, 0] == 103.0
    assert [x["score"] for x in r] == [0.0, 1.5] and len(r + r) == 4 and (r + r)[3]["n_eval"] == 2


def test_block_ordering_fill_is_comparable_to_the_references_amd():
    """The fill of the direct solver's minimum-degree block ordering against CSparse's own AMD (the reference's lm_var path,
    run here from the vendored sources): same league on the sphere graphs (measured 0.89 - 1.13 x)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_pgo as P
    if not P.have_csparse():
        pytest.skip("reference CSparse not built (oracle/_ref)")
    from lv_slam_b200.graph_slam import chol_analyze
    from lv_slam_b200.synth import posegraph as G
    for npl, laps in ((20, 10), (50, 20)):
        g = G.sphere(npl, laps, seed=7)
        o = P.OraclePGO()
        o.set_graph(g["poses7"], g["ij"], g["meas7"], g["info21"], g["huber"])
        lin = o.linearize()
        ok, _, _ = o.solve(1e-5 * np.max(np.abs(np.einsum("nii->ni", lin["Hd"]))), P.SOLVER_CSPARSE)
        assert ok
        n = len(g["poses7"])
        off = np.array(sorted({(min(a, b), max(a, b)) for a, b in np.asarray(g["ij"]) if a != b}), np.int32)
        st, _ = chol_analyze(n, off)
        ours = (st["nnz_l_blocks"] - n) * 36 + n * 21          # scalar entries of the block factor
        assert ours <= 1.3 * P.csparse_lnz(o)


def test_keyframe_dump_round_trip(tmp_path):
    """dump_service artefacts (graph.g2o + .kernels, %06d/data + cloud.pcd, special_nodes.csv, ggo_*_odom.txt) written and read
    back; `data` keeps the reference's token layout (keyframe.cpp:48-92) at Eigen's default 6 significant digits."""
    from lv_slam_b200 import keyframe_io as K
    from lv_slam_b200.graph_slam import GraphSLAM, load_kitti_poses
    from lv_slam_b200.synth import posegraph as G
    rng = np.random.default_rng(3)
    g = GraphSLAM("lm_var_cholmod")
    kfs, odoms = [], {}
    T = np.eye(4)
    for i in range(4):
        T = T @ G.matrix(np.concatenate([[2.0 + i, 0.1, 0.0], G.pose7(np.eye(4))[3:]]))
        node = g.add_se3_node(T)
        cloud = rng.normal(size=(50 + i, 4)).astype(np.float32)
        kfs.append(dict(stamp=(100 + i, 5000 * i), seq=3 * i, estimate=T.copy(), odom=T.copy(), accum_distance=2.5 * i, id=node.id(), cloud=cloud))
        if i:
            e = g.add_se3_edge(kfs[i - 1]["node"], node, np.linalg.inv(kfs[i - 1]["estimate"]) @ T, np.eye(6) * (i + 1))
            g.add_robust_kernel(e, "Huber", 1.0)
        kfs[-1]["node"] = node
    for s in range(0, 12):
        a, f = divmod(s, 3)
        a = min(a, 3)
        step = G.matrix(np.concatenate([[0.5 * (s - 3 * a), 0.0, 0.0], G.pose7(np.eye(4))[3:]]))
        odoms[s] = kfs[a]["odom"] @ step
    d = str(tmp_path / "dump")
    K.dump(d, g, kfs, odoms)
    txt = open(os.path.join(d, "000001", "data")).read().split("\n")
    assert txt[0] == "stamp 101 5000" and txt[1] == "estimate" and txt[6] == "odom" and txt[11].startswith("accum_distance 2.5") and txt[12] == "id 1"
    assert open(os.path.join(d, "special_nodes.csv")).read() == "anchor_node -1\nanchor_edge -1\nfloor_node -1\n"
    hdr = open(os.path.join(d, "000002", "cloud.pcd"), "rb").read(200).decode("ascii", "replace")
    assert "FIELDS x y z intensity" in hdr and "POINTS 52" in hdr and "DATA binary" in hdr
    g2 = GraphSLAM("lm_var_cholmod")
    back = K.load_dump(d, g2)
    assert len(back) == 4 and g2.num_vertices() == 4 and g2.num_edges() == 3
    for a, b in zip(kfs, back):
        assert b["stamp"] == a["stamp"] and b["id"] == a["id"] and b["node"].id() == a["id"]
        np.testing.assert_array_equal(b["cloud"], a["cloud"])                     # binary PCD: exact
        np.testing.assert_allclose(b["estimate"], a["estimate"], rtol=1e-5, atol=1e-5)
        assert abs(b["accum_distance"] - a["accum_distance"]) < 1e-5
    assert all(e.kernel == ("Huber", 1.0) for e in g2._edges)
    # the optimisation changed nothing here (estimate == odom), so the per-frame file is the odometry itself, from the first keyframe
    kf_file = load_kitti_poses(os.path.join(d, "ggo_kf_odom.txt"))
    wf_file = load_kitti_poses(os.path.join(d, "ggo_wf_odom.txt"))
    assert len(kf_file) == 4 and len(wf_file) == 12
    np.testing.assert_allclose(kf_file[2], kfs[2]["estimate"], atol=1e-6)
    first = np.linalg.inv(kfs[0]["estimate"])
    for s in range(12):
        np.testing.assert_allclose(wf_file[s], first @ odoms[s], atol=1e-6)
    # ascii PCD written by other tools reads the same
    with open(os.path.join(d, "a.pcd"), "w") as f:
        f.write("VERSION .7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 2\nHEIGHT 1\nPOINTS 2\nDATA ascii\n1 2 3\n4 5 6.5\n")
    np.testing.assert_array_equal(K.load_pcd(os.path.join(d, "a.pcd")), np.array([[1, 2, 3, 0], [4, 5, 6.5, 0]], np.float32))
    # a camera calibration conjugates the written poses (save_pose :1081-1097)
    with open(os.path.join(d, "calib.txt"), "w") as f:
        f.write("P0: 0\nP1: 0\nP2: 0\nP3: 0\nTr: 0 -1 0 0.1 0 0 -1 0.2 1 0 0 0.3\n")
    C4 = K.load_calib(os.path.join(d, "calib.txt"))
    K.save_pose(d, kfs, odoms, C4)
    np.testing.assert_allclose(load_kitti_poses(os.path.join(d, "ggo_kf_odom.txt"))[3], C4 @ kfs[3]["estimate"] @ np.linalg.inv(C4), atol=1e-6)


def test_save_pose_distributes_the_correction_as_the_reference_writes_it(tmp_path):
    """One segment of 4 frames whose optimised end moved by (0.4 m, 8 deg about z) against the odometry: the per-frame file applies
    the translation share 1/4 and — as written at global_graph_nodelet.cpp:1118 — slerp(4, q), i.e. FOUR times the angle."""
    from lv_slam_b200 import keyframe_io as K
    from lv_slam_b200.graph_slam import load_kitti_poses

    def rz(deg, t=(0, 0, 0)):
        a = np.deg2rad(deg)
        T = np.eye(4)
        T[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
        T[:3, 3] = t
        return T
    odoms = {s: rz(0, (1.0 * s, 0, 0)) for s in range(5)}
    corr = rz(8.0, (0.4, 0, 0))
    kfs = [dict(seq=0, estimate=np.eye(4)), dict(seq=4, estimate=odoms[4] @ corr)]
    n = K.save_pose(str(tmp_path), kfs, odoms)
    assert n == 4 + 1                                       # frames 0..3 of the segment, then frame 4 (the last keyframe's own)
    wf = load_kitti_poses(str(tmp_path / "ggo_wf_odom.txt"))
    np.testing.assert_allclose(wf[0], np.eye(4), atol=1e-9)
    np.testing.assert_allclose(wf[2], odoms[2] @ rz(32.0, (0.1, 0, 0)), atol=1e-6)
    np.testing.assert_allclose(wf[4], odoms[4] @ corr, atol=1e-6)


def test_shims_compile_and_link_against_the_abi(tmp_path):
    """The reference-side bindings of shim/  need PCL, Eigen, g2o and boost, none of which is in this image.
    They are compiled here against interface stubs (tests/shim_stubs: the names and signatures the shims touch, nothing more), with
    the registration classes instantiated for the reference's three point types in the three namespaces, and linked against
    liblvslam_b200.so with --no-undefined: every ABI call in the shims type-checks against include/lvslam_b200.h and resolves."""
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    from lv_slam_b200 import build
    build.build()
    stubs = os.path.join(ROOT, "tests", "shim_stubs")
    inc = ["-I" + stubs, "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "shim")]
    units = [("graph_slam", [os.path.join(ROOT, "shim", "graph_slam_b200.cpp")]), ("aux", [os.path.join(ROOT, "shim", "aux_b200.cpp")]),
             ("ndt_omp", [os.path.join(stubs, "instantiate.cpp")]), ("ndt_pca", ["-DLVS_SHIM_PCA", "-Dshim_probe=shim_probe_pca", os.path.join(stubs, "instantiate.cpp")]),
             ("ndt_ground", ["-DLVS_SHIM_GROUND", "-Dshim_probe=shim_probe_ground", os.path.join(stubs, "instantiate.cpp")]),
             ("nodelet_like", [os.path.join(stubs, "nodelet_like.cpp")])]       # the three classes in ONE translation unit
    objs = []
    for name, args in units:
        o = str(tmp_path / (name + ".o"))
        r = subprocess.run(["g++", "-std=c++14", "-Wall", "-Wextra", "-Werror", "-fPIC", "-c"] + inc + args + ["-o", o], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        objs.append(o)
    lib_dir = os.path.join(ROOT, "lv_slam_b200")
    so = str(tmp_path / "libshim.so")
    r = subprocess.run(["g++", "-shared"] + objs + ["-o", so, "-L" + lib_dir, "-llvslam_b200", "-Wl,--no-undefined", "-Wl,-rpath," + lib_dir],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    und = subprocess.run(["nm", "-D", "--undefined-only", so], capture_output=True, text=True).stdout
    used = sorted({ln.split()[-1] for ln in und.splitlines() if ln.split() and ln.split()[-1].startswith("lvs_")})
    header = open(os.path.join(ROOT, "include", "lvslam_b200.h")).read()
    assert used and all(sym + "(" in header for sym in used), used
    # both namespaces carry the three instantiations
    defined = subprocess.run(["nm", "-DC", "--defined-only", so], capture_output=True, text=True).stdout
    for ns in ("pclomp::NormalDistributionsTransform", "pclpca::NormalDistributionsTransform", "pclomp_ground::NormalDistributionsTransformGround"):
        for pt in ("PointXYZ,", "PointXYZI,", "PointXYZRGBL,"):
            assert any(ns + "<pcl::" + pt in ln and "computeTransformation" in ln for ln in defined.splitlines()), (ns, pt)


def test_header_is_plain_c_and_a_c_caller_links(tmp_path):
    """The boundary is a C ABI: the header compiles as pedantic C99 and a C translation unit that takes the address of every
    declared entry point links against the library (what a cgo / JNI / ctypes-free binding would do)."""
    import shutil
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    from lv_slam_b200 import build
    build.build()
    hdr = os.path.join(ROOT, "include", "lvslam_b200.h")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    src = tmp_path / "caller.c"
    names = _declared_functions()
    src.write_text('#include "lvslam_b200.h"\n#include <stdio.h>\ntypedef void (*fn)(void);\nint main(void) {\n  fn table[] = {\n'
                   + "".join("    (fn)%s,\n" % n for n in names) + "  };\n  printf(\"%d\\n\", (int)(sizeof table / sizeof table[0]));\n  return 0;\n}\n")
    exe = str(tmp_path / "caller")
    lib_dir = os.path.join(ROOT, "lv_slam_b200")
    r = subprocess.run(["gcc", "-std=c99", "-I" + os.path.join(ROOT, "include"), str(src), "-o", exe, "-L" + lib_dir, "-llvslam_b200", "-Wl,-rpath," + lib_dir],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and int(out.stdout) == len(names)
