"""Pose-graph timing on BASELINE config 4 (5 000 vertices / 19 599 edges) and, with --big, config 5's graph (50 000 / 198 999);
--no-cpu skips the CPU oracle legs."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import lv_slam_b200 as L
import oracle_pgo as P
from lv_slam_b200.synth import posegraph as G

for (npl, laps, cpu) in ((100, 50, "--no-cpu" not in sys.argv), (250, 200, "--big" in sys.argv and "--no-cpu" not in sys.argv)):
    if npl == 250 and "--big" not in sys.argv:
        continue
    t = time.time(); g = G.sphere(npl, laps, seed=7); tg = time.time() - t
    print("graph %d vertices / %d edges (generated in %.1f s)" % (len(g["poses7"]), len(g["ij"]), tg))
    # the third row answers why the two solver kinds end at different chi2 on the large graph: g2o's PCG stops at a relative residual of
    # 1e-6, LM then works with inexact steps and gives up (ten rejected trials) short of the optimum the exact solves reach; with a tight
    # PCG tolerance the iterative kind lands on the direct solver's chi2
    for solver, name in ((0, "lm_var (exact solve)"), (2, "lm_pcg"), (2, "lm_pcg tolerance 1e-12")):
        pg = L.PoseGraph(solver)
        if name.endswith("1e-12"): pg.set_options(pcg_tolerance=1e-12)
        t = time.time(); pg.set_graph(g["poses7"], g["ij"], g["meas7"], g["info21"], g["huber"]); ts = time.time() - t
        t = time.time(); st = pg.optimize(1024); to = time.time() - t
        if solver == 0: print("  direct solver structure:", pg.chol_info())
        print("  GPU %-22s set_graph %.1f ms, optimize wall %.1f ms (device %.1f ms: linearize %.2f ms, solve %.1f ms), iters %d trials %d pcg %d launches %d chi2 %.4g -> %.6g" % (
            name, ts * 1e3, to * 1e3, st["device_ms"], st["linearize_ms"], st["solve_ms"], st["iterations"], st["lm_trials"], st["pcg_iterations"], st["launches"], st["chi2_before"], st["chi2_after"]))
    if cpu:
        for solver, name in ((P.SOLVER_CSPARSE, "lm_var + CSparse"), (P.SOLVER_PCG, "lm_pcg")):
            o = P.OraclePGO(); o.set_graph(g["poses7"], g["ij"], g["meas7"], g["info21"], g["huber"])
            t = time.time(); r = o.optimize(1024, P.ALG_LM, solver); to = time.time() - t
            print("  CPU %-22s optimize %.1f ms, iters %d trials %d chi2 -> %.6g" % (name, to * 1e3, r["iterations"], r["trials"], r["chi2_after"]))
