"""Parity of the CUDA path for pclomp_ground::NormalDistributionsTransformGround (LVS_NDT_GROUND; include/ndt_omp/ndt_ground.h,
ndt_ground_impl.hpp) against the CPU oracle, through the C-ABI.  Same tolerances as tests/test_ndt_gpu.py: cell flags exact,
(score, gradient, Hessian) 1e-9 relative, per-iteration pose <= 1e-4 m / 1e-5 rad."""
import numpy as np
import pytest

import oracle_ndt as O
from test_ndt_gpu import _check_align, _mk, _relmax

pytestmark = pytest.mark.gpu


def _ground_pair(small_pair, mot, step=2):
    tgt = small_pair[0]
    motion = O.se3_exp_matrix4f(np.array(mot, dtype=np.float64))
    src = O.transform(tgt[::step], np.linalg.inv(motion.astype(np.float64)).astype(np.float32))
    return tgt, src


@pytest.mark.parametrize("resolution", [1.0, 2.0])
def test_ground_cell_flags_match_oracle(small_pair, scan_pair, resolution):
    for tgt in (small_pair[0], scan_pair[0]):
        n, o = _mk(O.VAR_GROUND, O.DIRECT1, resolution=resolution)
        n.setInputTarget(tgt); o.set_target(tgt)
        ang = o.leaf_angles()
        want = ((ang >= 0) & (ang < 10)).astype(np.int32)
        got = n.cell_horizontal()
        clear = np.abs(ang - 10) > 1e-6                     # a normal within 1e-6 degrees of the threshold may fall either way
        assert np.array_equal(got[clear], want[clear]) and 0 < want.sum() < (ang >= 0).sum()
        assert np.array_equal(n.cells()["keys"], o.leaves()["keys"])
    # the flag is a property of the ground variant only
    n, o = _mk(O.VAR_OMP, O.DIRECT1, resolution=resolution)
    n.setInputTarget(small_pair[0])
    assert not n.cell_horizontal().any()


@pytest.mark.parametrize("search", [O.DIRECT1, O.DIRECT7, O.DIRECT26, O.KDTREE])
def test_ground_derivatives_match_oracle(small_pair, search):
    tgt, src, guess, truth = small_pair
    n, o = _mk(O.VAR_GROUND, search)
    n.setInputTarget(tgt); o.set_target(tgt)
    n.setInputSource(src); o.set_source(src)
    rng = np.random.default_rng(4)
    p0 = O.se3_log_from_matrix4f(guess)
    for k in range(2):
        p = p0 + rng.normal(0, [0.05, 0.05, 0.02, 0.004, 0.004, 0.01])
        for hess in (True, False):
            gs, gg, gH = n.eval_derivatives(p, None, hess)
            os_, og, oH = o.eval_derivatives(p, None, hess)
            assert abs(os_) > 10 and abs(gs - os_) <= 1e-9 * abs(os_)
            assert _relmax(gg, og) < 1e-9
            assert (gg[[0, 1, 5]] == 0).all()                 # x, y, yaw are not solved for (ndt_ground_impl.hpp:554-561)
            if hess:
                assert _relmax(gH, oH) < 1e-9
                assert (gH[[0, 1, 5], :] == 0).all() and (gH[:, [0, 1, 5]] == 0).all() and gH[2:5, 2:5].all()
            else:
                assert not gH.any()
    # the gate and the doubling against the pclomp pass on the same object family: score of ground <= score of omp in magnitude,
    # and the all-double computeHessian of the line search is NOT masked (ndt_ground_impl.hpp calls the plain computeHessian)
    n2, o2 = _mk(O.VAR_OMP, search)
    n2.setInputTarget(tgt); n2.setInputSource(src)
    s_omp = n2.eval_derivatives(p0, guess, False)[0]
    s_gnd = n.eval_derivatives(p0, guess, False)[0]
    assert 0 < abs(s_gnd) < abs(s_omp)
    assert _relmax(n.eval_hessian(p0), o.eval_hessian(p0)) < 1e-10 and n.eval_hessian(p0)[0, 0] != 0


def test_ground_derivatives_full_scan(scan_pair):
    tgt, src, guess, truth = scan_pair
    n, o = _mk(O.VAR_GROUND, O.DIRECT1, resolution=2.0)
    n.setInputTarget(tgt); o.set_target(tgt)
    n.setInputSource(src); o.set_source(src)
    p = O.se3_log_from_matrix4f(guess)
    gs, gg, gH = n.eval_derivatives(p, guess, True)
    os_, og, oH = o.eval_derivatives(p, guess, True)
    assert abs(gs - os_) <= 1e-9 * abs(os_) and _relmax(gg, og) < 1e-9 and _relmax(gH, oH) < 1e-9


@pytest.mark.parametrize("search,resolution,mot", [(O.DIRECT1, 2.0, [0, 0, 0.05, 0, 0, 0]), (O.DIRECT7, 2.0, [0, 0, 0.08, 0, 0, 0]),
                                                   (O.DIRECT1, 4.0, [0, 0, 0.12, 0.006, -0.004, 0]), (O.KDTREE, 2.0, [0, 0, 0.05, 0, 0, 0])])
def test_ground_align_matches_oracle(small_pair, search, resolution, mot):
    tgt, src = _ground_pair(small_pair, mot)
    n, o = _mk(O.VAR_GROUND, search, resolution=resolution, trans_eps=0.001, max_iter=64)
    n.setInputTarget(tgt); o.set_target(tgt)
    g, r = _check_align(n, o, src, np.eye(4, dtype=np.float32))
    assert r["converged"] and 1 <= r["iterations"] < 64
    pf = O.se3_log_from_matrix4f(g["final"])
    assert abs(pf[2] - mot[2]) < 0.02                       # and it does recover the height
    assert (g["trace"][:, [6, 7, 11]] == 0).all()           # unit directions without x, y, yaw


def test_ground_first_iteration_can_end_the_align(small_pair):
    """ndt_ground_impl.hpp:173 drops the `nr_iterations_ &&` of ndt_omp_impl2.hpp:175-176: a first step below epsilon ends the align
    after ONE iteration (two for pclomp), lean_final_evaluation included."""
    tgt = small_pair[0]
    for lean in (0, 1):
        n, o = _mk(O.VAR_GROUND, O.DIRECT1, resolution=2.0, trans_eps=0.05, max_iter=64)
        n.setLeanFinalEvaluation(lean)
        n.setInputTarget(tgt); o.set_target(tgt)
        g, r = _check_align(n, o, tgt[::2], np.eye(4, dtype=np.float32))
        assert g["iterations"] == 1 and g["n_eval"] == 2
    n, o = _mk(O.VAR_OMP, O.DIRECT1, resolution=2.0, trans_eps=0.05, max_iter=64)
    n.setInputTarget(tgt); o.set_target(tgt)
    g, r = _check_align(n, o, tgt[::2], np.eye(4, dtype=np.float32))
    assert g["iterations"] == 2


def test_ground_line_search_path_and_batch(small_pair):
    """Forced More-Thuente loop (step_size <= epsilon / 2): the trials run the gated pass without Hessian and the accepted point gets the
    plain all-double computeHessian; then the same pairs through the batched entry point."""
    import lv_slam_b200 as L
    tgt, src = _ground_pair(small_pair, [0, 0, 0.05, 0, 0, 0])
    n, o = _mk(O.VAR_GROUND, O.DIRECT7, resolution=2.0, step_size=0.2, trans_eps=0.5, max_iter=4, threads=1)
    n.setInputTarget(tgt); o.set_target(tgt)
    g, r = _check_align(n, o, src, np.eye(4, dtype=np.float32))
    assert r["n_hess"] > 0
    # the batched entry point runs the same variant (different CTA split per pair, so the sums agree to rounding only)
    singles = []
    for s in (src, src[::2]):
        n1, _ = _mk(O.VAR_GROUND, O.DIRECT1, resolution=2.0, trans_eps=0.001)
        n1.setInputTarget(tgt); n1.setInputSource(s)
        n1.align(np.eye(4, dtype=np.float32))
        singles.append(n1.result())
    b = L.NdtBatch(1, 2, variant=L.LVS_NDT_GROUND, search_method=L.LVS_DIRECT1, resolution=2.0, transformation_epsilon=0.001, max_iterations=64)
    b.set_target(0, tgt); b.set_source(0, src); b.set_source(1, src[::2])
    res = b.align([0, 1], [0, 0], [np.eye(4, dtype=np.float32)] * 2)
    for r, one in zip(res, singles):
        assert r["iterations"] == one["iterations"] and np.max(np.abs(r["final"] - one["final"])) <= 1e-5


def test_ground_rejects_tolerance_mode():
    import lv_slam_b200 as L
    n = L.NormalDistributionsTransformGround()
    with pytest.raises(L.LvsError):
        n.setAccumulation(L.LVS_ACC_FAST)
