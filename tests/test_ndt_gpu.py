"""Parity of the CUDA NDT path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star / SURVEY.md §8c): voxel keys and point counts bit-exact; per-cell mean / inverse
covariance 1e-9 relative (fp64 on both sides); (score, gradient, Hessian) 1e-9 relative to the largest entry (the float32
per-point terms are bit-identical, only the fp64 summation order differs); per-iteration pose <= 1e-4 m / 1e-5 rad.
"""
import numpy as np
import pytest

import oracle_ndt as O

pytestmark = pytest.mark.gpu


def _mk(variant, search, **kw):
    import lv_slam_b200 as L
    n = L.NormalDistributionsTransform(variant=variant)
    n.setTransformationEpsilon(kw.get("trans_eps", 0.01))
    n.setMaximumIterations(kw.get("max_iter", 64))
    n.setNeighborhoodSearchMethod(search)
    n.setResolution(kw.get("resolution", 1.0))
    n.setStepSize(kw.get("step_size", 0.1))
    o = O.OracleNDT(variant=variant, resolution=kw.get("resolution", 1.0), step_size=kw.get("step_size", 0.1), trans_eps=kw.get("trans_eps", 0.01),
                    max_iter=kw.get("max_iter", 64), search=search, num_threads=kw.get("threads", 8))
    return n, o


def _relmax(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def _rot_angle(Ra, Rb):
    """Angle of Ra^T Rb from its skew part (well conditioned near zero, unlike arccos of the trace)."""
    d = Ra.astype(np.float64).T @ Rb.astype(np.float64)
    v = 0.5 * np.array([d[2, 1] - d[1, 2], d[0, 2] - d[2, 0], d[1, 0] - d[0, 1]])
    return float(np.arcsin(min(1.0, np.linalg.norm(v))))


def _check_cells(n, o):
    gc, oc = n.cells(), o.leaves()
    assert [a.tolist() for a in n.grid()] == [a.tolist() for a in o.grid()]
    assert np.array_equal(gc["keys"], oc["keys"])                       # every occupied cell, ascending key
    assert np.array_equal(gc["nr_points"], oc["nr_points"])            # raw count, -1 where the reference invalidates
    assert np.array_equal(gc["centroid"], oc["centroid"])              # float32 sums in input order: bit-exact
    assert np.array_equal(gc["weight"], oc["weight"])
    np.testing.assert_allclose(gc["mean"], oc["mean"], rtol=1e-12, atol=0)
    valid = oc["nr_points"] >= 6
    np.testing.assert_allclose(gc["evals"][valid], oc["evals"][valid], rtol=1e-9, atol=1e-300)
    scale = np.abs(oc["icov"]).reshape(-1, 9).max(axis=1)[:, None, None] + 1e-300
    assert np.max(np.abs(gc["icov"] - oc["icov"]) / scale) < 1e-9
    return int(valid.sum())


@pytest.mark.parametrize("variant", [O.VAR_OMP, O.VAR_PCA])
def test_voxel_grid_full_scan(scan_pair, variant):
    tgt, src, guess, truth = scan_pair
    n, o = _mk(variant, O.DIRECT7)
    n.setInputTarget(tgt); o.set_target(tgt)
    assert _check_cells(n, o) > 1000


@pytest.mark.parametrize("resolution", [0.5, 2.0, 0.7])
def test_voxel_grid_other_leaf_sizes(small_pair, resolution):
    tgt, src, guess, truth = small_pair
    n, o = _mk(O.VAR_PCA, O.DIRECT7, resolution=resolution)
    n.setInputTarget(tgt); o.set_target(tgt)
    _check_cells(n, o)


def test_voxel_grid_strided_and_nan_points(small_pair):
    tgt = small_pair[0]
    xyzi = np.zeros((tgt.shape[0], 8), np.float32)      # pcl::PointXYZI: 32-byte stride
    xyzi[:, :3] = tgt
    xyzi[::97, 1] = np.nan                               # non-finite points are skipped by applyFilter (:215-217)
    n, o = _mk(O.VAR_OMP, O.DIRECT7)
    n.setInputTarget(xyzi); o.set_target(xyzi)
    _check_cells(n, o)


def test_lookup_keys_bit_exact(scan_pair):
    tgt, src, guess, truth = scan_pair
    n, o = _mk(O.VAR_OMP, O.DIRECT7)
    n.setInputTarget(tgt); o.set_target(tgt)
    n.setInputSource(src); o.set_source(src)
    for T in (guess, truth.astype(np.float32), np.eye(4, dtype=np.float32)):
        gk = n.lookup_keys(T)
        ok = o.lookup_keys(O.transform(src, T))
        assert np.array_equal(gk, ok)
        assert (gk >= 0).sum() > 0.9 * len(gk)


@pytest.mark.parametrize("variant,search", [(O.VAR_OMP, O.DIRECT7), (O.VAR_PCA, O.DIRECT1), (O.VAR_OMP, O.DIRECT1), (O.VAR_PCA, O.DIRECT7),
                                            (O.VAR_OMP, O.DIRECT26), (O.VAR_PCA, O.DIRECT26), (O.VAR_OMP, O.KDTREE), (O.VAR_PCA, O.KDTREE)])
def test_derivatives_match_oracle(small_pair, variant, search):
    tgt, src, guess, truth = small_pair
    n, o = _mk(variant, search)
    n.setInputTarget(tgt); o.set_target(tgt)
    n.setInputSource(src); o.set_source(src)
    rng = np.random.default_rng(3)
    p0 = O.se3_log_from_matrix4f(guess)
    for k in range(3):
        p = p0 + rng.normal(0, [0.05, 0.05, 0.02, 0.004, 0.004, 0.01])
        for hess in (True, False):
            gs, gg, gH = n.eval_derivatives(p, None, hess)
            os_, og, oH = o.eval_derivatives(p, None, hess)
            assert abs(gs - os_) <= 1e-9 * abs(os_)
            assert _relmax(gg, og) < 1e-9
            if hess:
                assert _relmax(gH, oH) < 1e-9
                assert np.max(np.abs(oH - oH.T)) > 0          # the reference's Hessian is NOT symmetric; neither is ours
            else:
                assert not gH.any()
    # explicit transformed cloud that differs from exp(p) (the initial evaluation of computeTransformation)
    gs, gg, gH = n.eval_derivatives(p0, guess, True)
    os_, og, oH = o.eval_derivatives(p0, guess, True)
    assert abs(gs - os_) <= 1e-9 * abs(os_) and _relmax(gg, og) < 1e-9 and _relmax(gH, oH) < 1e-9


def test_derivatives_full_scan(scan_pair):
    tgt, src, guess, truth = scan_pair
    for variant, search in ((O.VAR_OMP, O.DIRECT7), (O.VAR_PCA, O.DIRECT1)):
        n, o = _mk(variant, search)
        n.setInputTarget(tgt); o.set_target(tgt)
        n.setInputSource(src); o.set_source(src)
        p = O.se3_log_from_matrix4f(guess)
        gs, gg, gH = n.eval_derivatives(p, guess, True)
        os_, og, oH = o.eval_derivatives(p, guess, True)
        assert abs(gs - os_) <= 1e-9 * abs(os_) and _relmax(gg, og) < 1e-9 and _relmax(gH, oH) < 1e-9


def test_double_hessian_and_score_match_oracle(small_pair):
    tgt, src, guess, truth = small_pair
    n, o = _mk(O.VAR_PCA, O.DIRECT1)
    n.setInputTarget(tgt); o.set_target(tgt)
    n.setInputSource(src); o.set_source(src)
    p = O.se3_log_from_matrix4f(guess)
    assert _relmax(n.eval_hessian(p), o.eval_hessian(p)) < 1e-10
    for T in (guess, truth.astype(np.float32)):
        a, b = n.calculateScore(T), o.calculate_score(T)
        assert abs(a - b) <= 1e-10 * abs(b)


def _check_align(n, o, src, guess):
    n.setInputSource(src); o.set_source(src)
    cloud = n.align(guess, want_cloud=True)
    g = n.result()
    r = o.align(guess, want_cloud=True)
    assert g["iterations"] == r["iterations"] and g["converged"] == r["converged"]
    assert g["n_eval"] == r["n_eval"] and g["n_hess"] == r["n_hess"]
    gt, rt = g["trace"], r["trace"]
    assert gt.shape == rt.shape
    if gt.shape[0]:
        # per-iteration pose: translation part <= 1e-4 m, rotation part <= 1e-5 rad
        assert np.max(np.abs(gt[:, 14:17] - rt[:, 14:17])) <= 1e-4
        assert np.max(np.abs(gt[:, 17:20] - rt[:, 17:20])) <= 1e-5
        assert np.array_equal(gt[:, 20:22], rt[:, 20:22])         # line-search trials, Hessian recomputation
    assert np.max(np.abs(g["final"][:3, 3] - r["final"][:3, 3])) <= 1e-4
    assert _rot_angle(g["final"][:3, :3], r["final"][:3, :3]) <= 1e-5
    assert abs(g["trans_probability"] - r["trans_probability"]) <= 1e-6 * abs(r["trans_probability"]) + 1e-300
    if cloud.size:
        assert np.max(np.abs(cloud - r["cloud"])) <= 2e-4 * max(1.0, float(np.max(np.abs(r["cloud"]))) / 100.0)
    return g, r


@pytest.mark.parametrize("variant,search", [(O.VAR_OMP, O.DIRECT7), (O.VAR_PCA, O.DIRECT1)])
def test_align_full_scan_matches_oracle(scan_pair, variant, search):
    tgt, src, guess, truth = scan_pair
    n, o = _mk(variant, search)
    n.setInputTarget(tgt); o.set_target(tgt)
    g, r = _check_align(n, o, src, guess)
    if variant == O.VAR_OMP:
        assert np.max(np.abs(g["final"][:3, 3] - truth[:3, 3])) < 0.05    # and it actually registers the pair


@pytest.mark.parametrize("variant,search", [(O.VAR_OMP, O.DIRECT1), (O.VAR_OMP, O.DIRECT26), (O.VAR_OMP, O.KDTREE), (O.VAR_PCA, O.DIRECT7)])
def test_align_small_matches_oracle(small_pair, variant, search):
    tgt, src, guess, truth = small_pair
    n, o = _mk(variant, search, max_iter=30)
    n.setInputTarget(tgt); o.set_target(tgt)
    _check_align(n, o, src, guess)


def test_align_line_search_and_double_hessian_path(small_pair):
    """step_size <= transformation_epsilon / 2 makes the reference's `interval_converged = (step_max - step_min) > 0`
    false, which is the only way its More-Thuente loop (and the serial all-double computeHessian) ever runs.  Starting
    next to the optimum with a forced 0.25 m step overshoots, so the loop actually iterates.  Every trial lands on the same
    clamped step, so the loop's exit hinges on `f_t > f_l` between two evaluations of the SAME point: the oracle runs with one
    thread here because its OpenMP `schedule(guided)` partial sums (like the reference's) are not run-to-run bit-stable."""
    tgt, src, guess, truth = small_pair
    for ss, eps, g0 in ((0.2, 0.5, truth.astype(np.float32)), (0.2, 0.5, guess), (0.1, 0.3, truth.astype(np.float32))):
        n, o = _mk(O.VAR_OMP, O.DIRECT7, step_size=ss, trans_eps=eps, max_iter=6, threads=1)
        n.setInputTarget(tgt); o.set_target(tgt)
        g, r = _check_align(n, o, src, g0)
        assert r["n_hess"] > 0 and r["trace"][:, 20].max() > 0
    n, o = _mk(O.VAR_PCA, O.DIRECT1, step_size=0.2, trans_eps=0.5, max_iter=6, threads=1)
    n.setInputTarget(tgt); o.set_target(tgt)
    g, r = _check_align(n, o, src, truth.astype(np.float32))
    assert r["n_hess"] > 0


@pytest.mark.parametrize("variant,search,acc", [(O.VAR_OMP, O.DIRECT7, 0), (O.VAR_PCA, O.DIRECT1, 0), (O.VAR_OMP, O.DIRECT7, 1), (O.VAR_OMP, O.KDTREE, 0)])
def test_lean_final_evaluation_changes_no_result(small_pair, scan_pair, variant, search, acc):
    """lvs_ndt_params::lean_final_evaluation skips the Hessian of the derivative pass that ends an align - a result the reference
    computes and never reads.  Everything align() exposes must stay bit-identical: final transform, iterations, converged flag,
    evaluation counts, trans_probability, the per-iteration trace, the aligned cloud; and the corner where the line search is alive
    (step_size <= epsilon / 2: the step is not known beforehand) must not be touched by it."""
    for pair, kw in ((small_pair, dict(max_iter=30)), (scan_pair, dict(max_iter=64)), (small_pair, dict(max_iter=3)), (small_pair, dict(step_size=0.2, trans_eps=0.5, max_iter=6))):
        if pair is scan_pair and search == O.KDTREE:
            continue
        tgt, src, guess, truth = pair
        out = []
        for lean in (0, 1):
            n, _ = _mk(variant, search, **kw)
            n.setAccumulation(acc); n.setLeanFinalEvaluation(lean)
            n.setInputTarget(tgt); n.setInputSource(src)
            cloud = n.align(guess, want_cloud=True)
            out.append((n.result(), cloud, n.getFitnessScore(1.0)))
        (a, ca, fa), (b, cb, fb) = out
        assert np.array_equal(a["final"], b["final"]) and a["iterations"] == b["iterations"] and a["converged"] == b["converged"]
        assert a["n_eval"] == b["n_eval"] and a["n_hess"] == b["n_hess"] and a["trans_probability"] == b["trans_probability"]
        assert np.array_equal(a["trace"], b["trace"]) and np.array_equal(ca, cb) and fa == fb
        assert a["iterations"] >= 2


def test_align_identity_guess_and_repeat(small_pair):
    tgt, src, guess, truth = small_pair
    n, o = _mk(O.VAR_OMP, O.DIRECT7, max_iter=20)
    n.setInputTarget(tgt); o.set_target(tgt)
    _check_align(n, o, src, np.eye(4, dtype=np.float32))
    first = n.getFinalTransformation()
    # the reference nodelet aligns scan 1 a second time from the first result (scan_matching_odom_nodelet.cpp:223-227)
    _check_align(n, o, src, first)


def test_align_is_run_to_run_deterministic(small_pair):
    tgt, src, guess, truth = small_pair
    n, o = _mk(O.VAR_OMP, O.DIRECT7, max_iter=20)
    n.setInputTarget(tgt); n.setInputSource(src)
    n.align(guess); a = n.result()
    n.align(guess); b = n.result()
    assert np.array_equal(a["final"], b["final"]) and np.array_equal(a["trace"], b["trace"])


def test_edge_cases(small_pair):
    import lv_slam_b200 as L
    tgt, src, guess, truth = small_pair
    n, o = _mk(O.VAR_OMP, O.DIRECT7, max_iter=10)
    with pytest.raises(L.LvsError) as e:
        n.align(guess)
    assert e.value.status == -5                                   # align before setInputTarget
    n.setInputTarget(tgt); o.set_target(tgt)
    with pytest.raises(L.LvsError) as e:
        n.align(guess)
    assert e.value.status == -6                                   # align before setInputSource
    # source entirely outside the target box: no neighbours, zero gradient -> delta_p_norm == 0 -> converged, 0 iterations
    far = (src + np.float32(5000.0)).astype(np.float32)
    _check_align(n, o, far, np.eye(4, dtype=np.float32))
    assert n.getFinalNumIteration() == 0 and n.hasConverged()
    # empty source
    n.setInputSource(np.zeros((0, 3), np.float32))
    n.align(guess)
    assert n.getFinalNumIteration() == 0
    # a target with fewer than min_points_per_voxel points per cell has no usable cell
    n.setInputTarget(tgt[:5]); o.set_target(tgt[:5])
    _check_align(n, o, src, guess)
    # empty target cloud
    n.setInputTarget(np.zeros((0, 3), np.float32))
    n.setInputSource(src)
    n.align(guess)
    assert n.getFinalNumIteration() == 0
    # ragged sizes that are not multiples of the CTA tile
    for cnt in (1, 31, 257, 1025):
        n.setInputTarget(tgt); o.set_target(tgt)
        _check_align(n, o, src[:cnt], guess)


def test_set_resolution_revoxelises(small_pair):
    tgt, src, guess, truth = small_pair
    n, o = _mk(O.VAR_OMP, O.DIRECT7)
    n.setInputTarget(tgt); o.set_target(tgt)
    n.setResolution(0.5); o.set(resolution=0.5)                   # ndt_omp.h:126-136
    _check_cells(n, o)


def test_batch_matches_single(small_pair, scan_pair):
    import lv_slam_b200 as L
    tgt, src, guess, truth = small_pair
    tgt2, src2 = scan_pair[0][::3].copy(), scan_pair[1][::3].copy()
    b = L.NdtBatch(2, 3, transformation_epsilon=0.01, max_iterations=30, search_method=L.LVS_DIRECT7)
    b.set_target(0, tgt); b.set_target(1, tgt2)
    b.set_source(0, src); b.set_source(1, src2); b.set_source(2, src[:4000])
    g2 = guess.copy(); g2[0, 3] = 1.0
    pairs = [(0, 0, guess), (1, 1, scan_pair[2]), (2, 0, g2), (0, 0, g2), (1, 1, np.eye(4, dtype=np.float32))]
    res = b.align([p[0] for p in pairs], [p[1] for p in pairs], [p[2] for p in pairs])
    srcs, tgts = [src, src2, src[:4000]], [tgt, tgt2]
    for (s, t, g), r in zip(pairs, res):
        o = O.OracleNDT(variant=O.VAR_OMP, trans_eps=0.01, max_iter=30, search=O.DIRECT7, num_threads=8)
        o.set_target(tgts[t]); o.set_source(srcs[s])
        ref = o.align(g)
        assert r["iterations"] == ref["iterations"] and r["converged"] == ref["converged"]
        assert np.max(np.abs(r["final"][:3, 3] - ref["final"][:3, 3])) <= 1e-4
        assert _rot_angle(r["final"][:3, :3], ref["final"][:3, :3]) <= 1e-5


def test_device_resident_inputs(small_pair):
    import torch
    tgt, src, guess, truth = small_pair
    n, o = _mk(O.VAR_OMP, O.DIRECT7, max_iter=20)
    n.setInputTarget(torch.from_numpy(tgt).cuda()); o.set_target(tgt)
    o.set_source(src)
    n.setInputSource(torch.from_numpy(src).cuda())
    n.align(guess)
    r = o.align(guess)
    assert n.getFinalNumIteration() == r["iterations"]
    assert np.max(np.abs(n.getFinalTransformation() - r["final"])) <= 1e-4


def test_plural_setters_and_overlapped_uploads(small_pair, scan_pair):
    """set_targets / set_sources (one call, fused repack) from pinned host, pageable host and resident clouds, with targets
    re-set several times in a row (build lanes, staging ring, slot re-use): every path must give the single-call results bit for bit."""
    import torch
    import lv_slam_b200 as L
    tgt, src, guess, truth = small_pair
    tgt2, src2 = scan_pair[0][::3].copy(), scan_pair[1][::3].copy()
    clouds_t = [tgt, tgt2, tgt[::2].copy(), tgt2[::2].copy(), tgt, tgt2]
    clouds_s = [src, src2, src[:4000].copy(), src2[:9000].copy(), src[::2].copy(), src2[::2].copy(), src, src2]
    pairs = [(i, i % len(clouds_t), guess if i % 2 == 0 else scan_pair[2]) for i in range(len(clouds_s))]
    kw = dict(transformation_epsilon=0.01, max_iterations=12, search_method=L.LVS_DIRECT7)
    ref = L.NdtBatch(len(clouds_t), len(clouds_s), **kw)
    for i, c in enumerate(clouds_t):
        ref.set_target(i, c)
    for i, c in enumerate(clouds_s):
        ref.set_source(i, c)
    want = ref.align([p[0] for p in pairs], [p[1] for p in pairs], [p[2] for p in pairs])

    def as_kind(c, kind):
        t = torch.from_numpy(c)
        return t.pin_memory() if kind == "pinned" else t.cuda() if kind == "resident" else c

    for kind in ("pinned", "pageable", "resident"):
        b = L.NdtBatch(len(clouds_t), len(clouds_s), **kw)
        tt = [as_kind(c, kind) for c in clouds_t]
        ss = [as_kind(c, kind) for c in clouds_s]
        junk_t = [as_kind(np.ascontiguousarray(c[::-1]), kind) for c in clouds_t]
        b.set_targets(list(range(len(tt))), junk_t)            # overwritten below without an align in between
        b.set_sources(list(range(len(ss))), ss)
        b.set_targets(list(range(len(tt))), tt)
        got = []
        for a in range(0, len(pairs), 3):                      # grouped aligns: later uploads overlap earlier aligns
            g = pairs[a:a + 3]
            got += b.align([p[0] for p in g], [p[1] for p in g], [p[2] for p in g])
        b.wait_uploads()
        for w, r in zip(want, got):
            assert r["iterations"] == w["iterations"] and np.array_equal(r["final"], w["final"]), kind
        # second round on the same object: slots re-used with different clouds
        b.set_sources([0, 1], [ss[1], ss[0]])
        r2 = b.align([0, 1], [1, 0], [pairs[1][2], pairs[0][2]])
        assert np.array_equal(r2[0]["final"], want[1]["final"]) and np.array_equal(r2[1]["final"], want[0]["final"]), kind
        b.close()
    ref.close()


def test_point_sharding_self_exchange(small_pair):
    """world = 1: the evaluation kernel sends its 43 sums through its own mailbox (same code path as the multi-GPU exchange, one
    rank) - results must equal the unsharded object bit for bit, on the batch and on the single registration object."""
    import lv_slam_b200 as L
    tgt, src, guess, truth = small_pair
    kw = dict(transformation_epsilon=0.01, max_iterations=20, search_method=L.LVS_DIRECT7)
    a = L.NdtBatch(1, 2, **kw)
    b = L.NdtBatch(1, 2, **kw)
    b.enable_point_sharding(0, 1, 4, lambda blob: [blob])
    for nb in (a, b):
        nb.set_target(0, tgt); nb.set_source(0, src); nb.set_source(1, src[:5000])
    g2 = guess.copy(); g2[1, 3] += 0.2
    ra = a.align([0, 1, 0], [0, 0, 0], [guess, guess, g2])
    rb = b.align([0, 1, 0], [0, 0, 0], [guess, guess, g2])
    for x, y in zip(ra, rb):
        assert x["iterations"] == y["iterations"] and x["n_eval"] == y["n_eval"]
        assert np.array_equal(x["final"], y["final"]) and x["score"] == y["score"]
    n, o = _mk(O.VAR_PCA, O.DIRECT1, max_iter=15)
    n.enable_point_sharding(0, 1, lambda blob: [blob])
    n.setInputTarget(tgt); o.set_target(tgt)
    n.setInputSource(src); o.set_source(src)
    n.align(guess)
    r = o.align(guess)
    assert n.getFinalNumIteration() == r["iterations"]
    assert np.max(np.abs(n.getFinalTransformation() - r["final"])) <= 1e-6


def _shift(T, dx, dy=0.0, dz=0.0):
    S = np.array(T, dtype=np.float64).copy()
    S[:3, 3] += (dx, dy, dz)
    return S.astype(np.float32)


def test_fitness_score_matches_oracle(small_pair):
    """getFitnessScore (loop_detector.hpp:176,255) / calc_fitness_score (information_matrix_calculator.cpp:53-87): the float
    nearest-neighbour distances are bit-identical to the exhaustive CPU scan, so the correspondence COUNT is exact for any
    max_range and the mean agrees to the order of the fp64 summation (1e-12 relative)."""
    tgt, src, guess, truth = small_pair
    n, o = _mk(O.VAR_OMP, O.DIRECT7)
    n.setInputTarget(tgt); n.setInputSource(src)
    o.set_target(tgt); o.set_source(src)
    big = np.finfo(np.float64).max
    cases = [(truth, big), (guess, big), (guess, 0.25), (truth, 0.01), (_shift(truth, 0.0, 0.0, 7.5), big),      # 7.5 m above: ring search gives up
             (_shift(truth, 400.0, -300.0, 0.0), big), (_shift(truth, 400.0), 1.0)]                               # far outside the grid: brute force
    for T, mr in cases:
        gs, gn = n.getFitnessScore(mr, T=T, with_count=True)
        os_, on = o.fitness_score(T, mr)
        assert gn == on, (gn, on, mr)
        if on == 0:
            assert gs == big and os_ == big
        else:
            assert abs(gs - os_) <= 1e-12 * os_, (gs, os_)
    # T = None: the final transformation of the last align
    n.align(guess)
    o.align(guess)
    fin = n.getFinalTransformation()
    assert n.getFitnessScore(0.5) == n.getFitnessScore(0.5, T=fin)
    assert abs(n.getFitnessScore(0.5) - o.fitness_score(fin, 0.5)[0]) <= 1e-12 * o.fitness_score(fin, 0.5)[0]
    # non-finite source points have no neighbour and drop out of the mean; an empty source has no correspondence at all
    s2 = src.copy(); s2[::7, 1] = np.nan
    n.setInputSource(s2); o.set_source(s2)
    gs, gn = n.getFitnessScore(big, T=truth, with_count=True)
    os_, on = o.fitness_score(truth, big)
    assert gn == on == int(np.isfinite(s2[:, :3]).all(axis=1).sum()) and abs(gs - os_) <= 1e-12 * os_
    n.setInputSource(np.zeros((0, 3), np.float32))
    assert n.getFitnessScore(big, T=truth, with_count=True) == (big, 0)


def test_fitness_score_dense_cells_far_queries_and_bad_target_points():
    """The three passes of the search against the exhaustive scan: a cell with thousands of points next to the queries (the
    thread pass runs out of budget -> warp pass), queries several cells away from every point (outer shells), queries far outside
    the grid (brute-force tiles, more than one tile and more than one query group), NaN / inf target points, and a leaf size that
    is not a power of two (the slab order must hold for the binning arithmetic, not for the ideal cell)."""
    rng = np.random.default_rng(11)
    dense = rng.normal([3.4, 2.6, 0.5], [0.25, 0.25, 0.02], size=(6000, 3))              # a slab-like blob inside one or two cells
    wall = np.stack([np.full(3000, 9.3), rng.uniform(-6, 6, 3000), rng.uniform(0, 3, 3000)], axis=1)
    sparse = rng.uniform([-20, -20, -1], [20, 20, 3], size=(1500, 3))
    tgt = np.concatenate([dense, wall, sparse]).astype(np.float32)
    tgt[::97, 0] = np.nan
    tgt[5::211, 2] = np.inf
    near = dense[:1500] + rng.normal(0, [0.6, 0.6, 0.3], size=(1500, 3))                 # around the blob, mostly outside it
    mid = rng.uniform([-40, -40, -6], [40, 40, 9], size=(1500, 3))                        # up to many cells from anything
    far = rng.uniform([150, -300, -5], [400, 300, 60], size=(300, 3))                     # outside the grid
    src = np.concatenate([near, mid, far]).astype(np.float32)
    big = np.finfo(np.float64).max
    for res in (1.0, 0.7, 2.0):
        n, o = _mk(O.VAR_OMP, O.DIRECT7, resolution=res)
        n.setInputTarget(tgt); n.setInputSource(src)
        o.set_target(tgt); o.set_source(src)
        for T, mr in ((np.eye(4, dtype=np.float32), big), (_shift(np.eye(4), 0.3, -0.2, 0.1), 4.0), (_shift(np.eye(4), 0.0, 0.0, 30.0), big)):
            gs, gn = n.getFitnessScore(mr, T=T, with_count=True)
            os_, on = o.fitness_score(T, mr)
            assert gn == on, (res, gn, on)
            assert abs(gs - os_) <= 1e-12 * os_, (res, gs, os_)


def test_fitness_score_full_scan_properties(scan_pair):
    """At BASELINE's full size the exhaustive oracle is too slow; size-independent properties instead."""
    tgt, src, guess, truth = scan_pair
    n, _ = _mk(O.VAR_OMP, O.DIRECT7)
    n.setInputTarget(tgt); n.setInputSource(tgt)
    big = np.finfo(np.float64).max
    assert n.getFitnessScore(big, T=np.eye(4, dtype=np.float32), with_count=True) == (0.0, int(np.isfinite(tgt[:, :3]).all(axis=1).sum()))   # a cloud against itself
    n.setInputSource(src)
    s_truth, c_truth = n.getFitnessScore(big, T=truth, with_count=True)
    s_guess, c_guess = n.getFitnessScore(big, T=guess, with_count=True)
    assert c_truth == c_guess == len(src) and s_truth < s_guess
    # the count is monotone in max_range and the capped mean never exceeds the cap
    prev = 0
    for mr in (1e-4, 1e-2, 0.25, 4.0, big):
        s, c = n.getFitnessScore(mr, T=truth, with_count=True)
        assert c >= prev and (c == 0 or s <= mr)
        prev = c
    # batch entry point: same numbers as the single-object one
    import lv_slam_b200 as L
    b = L.NdtBatch(1, 1)
    b.set_target(0, tgt); b.set_source(0, src)
    assert b.fitness_score(0, 0, truth, 0.25) == n.getFitnessScore(0.25, T=truth, with_count=True)


def test_information_matrix_calculator_mirror(small_pair):
    """InformationMatrixCalculator (information_matrix_calculator.cpp:27-51): constant matrix, and the fitness-weighted one."""
    import lv_slam_b200 as L
    tgt, src, guess, truth = small_pair
    c = L.InformationMatrixCalculator(use_const_inf_matrix=True, const_stddev_x=0.5, const_stddev_q=0.1)
    assert np.array_equal(c.calc_information_matrix(tgt, src, truth), np.diag([2.0, 2, 2, 10, 10, 10]))     # dlo_lfa_ggo_kitti.launch:134-136
    c = L.InformationMatrixCalculator(fitness_score_thresh=2.5)
    o = O.OracleNDT(num_threads=8); o.set_target(tgt); o.set_source(src)
    fs = o.fitness_score(truth)[0]
    wx = np.float32(0.1 ** 2 + (5.0 ** 2 - 0.1 ** 2) * (1 - np.exp(-20.0 * fs)) / (1 - np.exp(-20.0 * 2.5)))
    wq = np.float32(0.05 ** 2 + (0.2 ** 2 - 0.05 ** 2) * (1 - np.exp(-20.0 * fs)) / (1 - np.exp(-20.0 * 2.5)))
    inf = c.calc_information_matrix(tgt, src, truth)
    np.testing.assert_allclose(np.diag(inf), [1 / wx] * 3 + [1 / wq] * 3, rtol=1e-6)
    assert np.count_nonzero(inf - np.diag(np.diag(inf))) == 0


def test_prefilter_matches_oracle(scan_pair, small_pair):
    """PrefilteringNodelet chain (distance filter 0.5-100 m, VoxelGrid 0.1 m; launch/dlo_lfa_ggo_kitti.launch:30-36) on a full 64-beam
    scan with an intensity column: same points, same order, bit for bit (both sides sum a leaf in input order)."""
    import lv_slam_b200 as L
    tgt = scan_pair[0]
    rng = np.random.default_rng(3)
    cloud = np.concatenate([tgt[:, :3], rng.random((len(tgt), 1), dtype=np.float32)], axis=1).astype(np.float32)
    cloud[::997, 2] = np.nan
    cloud[5::1201, :3] *= 0.001                                  # inside the near threshold
    pf = L.Prefilter(distance_near_thresh=0.5, distance_far_thresh=100.0, downsample_resolution=0.1)
    got = pf.filter(cloud)
    exp, fl = O.prefilter(cloud, 0.5, 100.0, True, 0.1)
    assert fl == 0 and pf.last_flags == 0
    assert got.shape == exp.shape and 0.3 * len(cloud) < len(got) < len(cloud)
    assert np.array_equal(got, exp)
    # xyz only, a 32-byte PointXYZI stride, other leaf sizes, and no downsampling at all
    wide = np.zeros((len(cloud), 8), np.float32); wide[:, :3] = cloud[:, :3]
    for leaf in (0.1, 0.25, 1.0):
        pf.downsample_resolution = leaf
        assert np.array_equal(pf.filter(wide[:, :3]), O.prefilter(cloud[:, :3], 0.5, 100.0, True, leaf)[0])
    pf.downsample_method = "NONE"
    assert np.array_equal(pf.filter(cloud), O.prefilter(cloud, 0.5, 100.0, True, 0.0)[0])
    # overflow guard -> pass-through with the flag; empty input; everything filtered away
    pf.downsample_method, pf.downsample_resolution = "VOXELGRID", 1e-4
    got = pf.filter(cloud[:, :3])
    exp, fl = O.prefilter(cloud[:, :3], 0.5, 100.0, True, 1e-4)
    assert fl == 1 and pf.last_flags == 1 and np.array_equal(got, exp)
    pf.downsample_resolution = 0.1
    assert pf.filter(np.zeros((0, 3), np.float32)).shape == (0, 3)
    pf.distance_far_thresh = 0.6
    assert len(pf.filter(small_pair[0])) == len(O.prefilter(small_pair[0], 0.5, 0.6, True, 0.1)[0])
    # device-resident input and output
    import torch
    pf.distance_far_thresh = 100.0
    got = pf.filter(torch.from_numpy(cloud).cuda())
    assert got.is_cuda and np.array_equal(got.cpu().numpy(), O.prefilter(cloud, 0.5, 100.0, True, 0.1)[0])
    # the filtered scan feeds the registration like the nodelet chain does
    n, o = _mk(O.VAR_PCA, O.DIRECT1)
    f_t, f_s = pf.filter(small_pair[0]), pf.filter(small_pair[1])
    n.setInputTarget(f_t); n.setInputSource(f_s); o.set_target(f_t); o.set_source(f_s)
    _check_align(n, o, f_s, small_pair[2])


def test_align_begin_end_overlaps_the_next_batch(small_pair, scan_pair):
    """align_begin / align_end: the same results as the one-call align, and the next batch may be staged into other slots while the
    aligns are in flight (lvs_ndt_batch_align_begin / _end)."""
    import lv_slam_b200 as L
    tgt, src, guess, truth = small_pair
    tgt2, src2, guess2, _ = scan_pair
    b = L.NdtBatch(2, 2, transformation_epsilon=0.01, max_iterations=64)
    b.set_targets([0], [tgt]); b.set_sources([0], [src])
    ref = b.align([0], [0], [guess])[0]
    b.align_begin([0], [0], [guess])
    b.set_targets([1], [tgt2]); b.set_sources([1], [src2])          # staged while pair (0, 0) is being aligned
    with pytest.raises(L.LvsError):
        b.align_begin([1], [1], [guess2])                           # one align in flight per object
    with pytest.raises(L.LvsError):
        b.fitness_score(0, 0, truth)
    with pytest.raises(L.LvsError):
        b.set_sources([0], [src2])                                  # a cloud the align in flight reads cannot be replaced
    with pytest.raises(L.LvsError):
        b.set_targets([1, 0], [tgt2, tgt2])
    got = b.align_end()[0]
    assert np.array_equal(got["final"], ref["final"]) and got["iterations"] == ref["iterations"] and got["n_eval"] == ref["n_eval"]
    with pytest.raises(L.LvsError):
        b.align_end()                                               # nothing in flight
    one = L.NdtBatch(1, 1, transformation_epsilon=0.01, max_iterations=64)
    one.set_target(0, tgt2); one.set_source(0, src2)
    exp = one.align([0], [0], [guess2])[0]
    b.align_begin([1], [1], [guess2])
    b.set_targets([0], [tgt]); b.set_sources([0], [src])            # and back again
    got2 = b.align_end()[0]
    assert np.array_equal(got2["final"], exp["final"]) and got2["iterations"] == exp["iterations"]
    assert np.array_equal(b.align([0], [0], [guess])[0]["final"], ref["final"])


def test_window_map_matches_oracle(small_pair):
    """The window cloud of the global-graph nodelet (global_graph_nodelet.cpp:199-243): scans moved into the window's frame with
    double matrices, concatenated, 0.1 m VoxelGrid - bit for bit against the restatement."""
    import lv_slam_b200 as L
    from lv_slam_b200 import synth
    rng = np.random.default_rng(11)
    scans = [synth.scan(77, f, np.array([1.2 * f, 0.0, 0.0, 0.01 * f, 0.0, 0.0]), 16, 600) for f in range(4)]
    scans = [np.concatenate([s, rng.random((len(s), 1), dtype=np.float32)], axis=1) for s in scans]
    Ts = [None] + [synth.pose_matrix(np.array([1.2 * f, 0.0, 0.0, 0.01 * f, 0.0, 0.0])) for f in range(1, 4)]
    w = L.WindowMap(0.1)
    w.start(scans[0])
    for s, T in zip(scans[1:], Ts[1:]):
        w.add(s, T)
    got = w.flush()
    exp = O.window_map(scans, Ts, 0.1)
    assert got.shape == exp.shape and len(got) > len(scans[0]) and np.array_equal(got, exp)
    # a second window on the same object, xyz only, coarser leaf
    w.leaf = 0.5
    w.start(scans[2][:, :3]); w.add(scans[3][:, :3], Ts[1])
    assert np.array_equal(w.flush(), O.window_map([scans[2][:, :3], scans[3][:, :3]], [None, Ts[1]], 0.5))


def test_edge_cases_of_the_neighbouring_stages(small_pair):
    """Empty and degenerate inputs of getFitnessScore, the prefilter and the window map behave like the restatement."""
    import torch
    import lv_slam_b200 as L
    tgt, src, guess, truth = small_pair
    big = np.finfo(np.float64).max
    n, o = _mk(O.VAR_OMP, O.DIRECT7)
    # a target without a single finite point: no neighbour for anybody
    bad = np.full((50, 3), np.nan, np.float32)
    n.setInputTarget(bad); n.setInputSource(src); o.set_target(bad); o.set_source(src)
    assert n.getFitnessScore(big, T=truth, with_count=True) == (big, 0) and o.fitness_score(truth, big) == (big, 0)
    n.setInputTarget(np.zeros((0, 3), np.float32))
    assert n.getFitnessScore(big, T=truth, with_count=True) == (big, 0)
    # a one-point target: every source point has the same neighbour
    one = np.array([[1.0, 2.0, 0.5]], np.float32)
    n.setInputTarget(one); o.set_target(one)
    gs, gn = n.getFitnessScore(big, T=truth, with_count=True)
    os_, on = o.fitness_score(truth, big)
    assert gn == on == len(src) and abs(gs - os_) <= 1e-12 * os_
    # prefilter: nothing finite, a single point, everything inside one leaf
    pf = L.Prefilter(distance_near_thresh=0.5, distance_far_thresh=100.0, downsample_resolution=0.1)
    assert pf.filter(bad).shape[0] == 0
    assert np.array_equal(pf.filter(one), one)
    clump = (np.array([[5.0, 5.0, 1.0]], np.float32) + 0.01 * np.random.default_rng(1).random((40, 3), dtype=np.float32)).astype(np.float32)
    got, exp = pf.filter(clump), O.prefilter(clump, 0.5, 100.0, True, 0.1)[0]
    assert np.array_equal(got, exp) and 1 <= len(got) <= 8
    # window map fed from device memory, identity transforms
    w = L.WindowMap(0.1)
    w.start(torch.from_numpy(tgt).cuda())
    w.add(torch.from_numpy(src).cuda(), np.eye(4))
    assert np.array_equal(w.flush(), O.window_map([tgt, src], [None, np.eye(4)], 0.1))
