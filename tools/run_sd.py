#!/usr/bin/env python
"""Run the regularized SD algorithm on a small SMPS instance with the cut formation on the GPU.

The Python twin of the reference's driver script test/instance_test/sd_single_cut_test.jl:
one epigraph, x0 from the first-stage bounds (or --x0), rho = 0.1, one new scenario per iteration;
second-stage LPs and the master QP by HiGHS on the CPU, everything of the cut-formation path
(scenario deltas, dual-vertex pool, argmax, cuts, cut list, incumbent test, master rows) by
libsqlp_b200.so.  Needs a CUDA device.

  python tools/run_sd.py --instance lands --iterations 200
  python tools/run_sd.py --instance baa99-20 --iterations 1000 --lower-bound -500000
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instance", default="lands", help="a tests/golden/instances/<name>_full.npz fixture")
    ap.add_argument("--smps", default=None, metavar="PREFIX",
                    help="read PREFIX.cor / .tim / .sto with the library's native SMPS reader instead of a fixture "
                         "(e.g. spInput/lands/lands); the device epigraph is built straight from the files")
    ap.add_argument("--fused", action="store_true", help="one library call per iteration (sqlp_cell_sd_step)")
    ap.add_argument("--iterations", type=int, default=200)
    ap.add_argument("--lower-bound", type=float, default=0.0)
    ap.add_argument("--rho", type=float, default=0.1)
    ap.add_argument("--adaptive", action="store_true", help="AdaptiveQuadScalarSchedule instead of a constant rho")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--x0", type=float, nargs="*", default=None)
    args = ap.parse_args()

    from sqlp_b200 import sd, smps, twosd as T
    native = sto = None
    if args.smps:
        native = smps.NativeSmps(args.smps + ".cor", args.smps + ".tim", args.smps + ".sto")
        sto = native.sto()
        zf = smps.full_tables(native.cor(), native.stage2(), sto)
    else:
        zf = dict(np.load(os.path.join(ROOT, "tests", "golden", "instances", f"{args.instance}_full.npz")))
    n1, m2, s = int(zf["n1"]), int(zf["m2"]), len(zf["pos_row"])
    dvs = T.sdDualVertexSet(m2=m2)
    fs = sd.FirstStage(zf["x_cost"], zf["A1"], zf["row_lower"], zf["row_upper"], zf["x_lower"], zf["x_upper"])
    cell = sd.sdCell(fs, dvs, device_cuts=not args.fused, fused_step=args.fused)
    if native is not None:
        epi = T.sdEpigraph.from_smps(native, 1.0, args.lower_bound, dvs)
    else:
        coef = T.sdSubprobCoefficients.from_tables(zf["rbar"], zf["T_colptr"], zf["T_rowval"], zf["T_nzval"],
                                                   zf["pos_row"], zf["pos_col"])
        epi = T.sdEpigraph(coef, 1.0, args.lower_bound, dvs)
    sd.bind_epigraph_(cell, epi)
    Tm = np.zeros((m2, n1))
    for j in range(n1):
        for q in range(zf["T_colptr"][j], zf["T_colptr"][j + 1]):
            Tm[zf["T_rowval"][q], j] = zf["T_nzval"][q]
    lp = sd.Stage2LP(zf["W"], zf["cost"], zf["y_lower"], zf["y_upper"], zf["directions"], zf["rbar"], Tm,
                     zf["pos_row"], zf["pos_col"])
    if args.x0 is not None and len(args.x0) == n1:
        x0 = np.asarray(args.x0, float)
    else:   # a feasible start: the first-stage LP optimum (the reference uses a 10-scenario all_in_one model)
        from scipy.optimize import linprog
        rows = [(zf["A1"][i], zf["row_lower"][i], zf["row_upper"][i]) for i in range(len(zf["A1"]))]
        A_ub = [a for a, lo, up in rows if np.isfinite(up)] + [-a for a, lo, up in rows if np.isfinite(lo)]
        b_ub = [up for a, lo, up in rows if np.isfinite(up)] + [-lo for a, lo, up in rows if np.isfinite(lo)]
        res = linprog(zf["x_cost"], A_ub=np.array(A_ub) if A_ub else None, b_ub=np.array(b_ub) if b_ub else None,
                      bounds=[(None if np.isinf(l) else l, None if np.isinf(u) else u)
                              for l, u in zip(zf["x_lower"], zf["x_upper"])], method="highs")
        x0 = res.x if res.status == 0 else np.where(np.isfinite(zf["x_lower"]), zf["x_lower"], 0.0)
    cell.x_candidate[:] = x0
    cell.x_incumbent[:] = x0
    sched = sd.AdaptiveQuadScalarSchedule() if args.adaptive else sd.ConstantQuadScalarSchedule(args.rho)
    cell.ext["quad_scalar"] = args.rho

    rng = np.random.default_rng(args.seed)
    cdf, vals, cnt = zf["out_cdf"], zf["out_vals"], zf["out_cnt"]

    def draw():                         # rand(sto): INDEP DISCRETE from the tables, NORMAL / UNIFORM when read from files
        u = rng.random(s)
        if sto is not None:
            return smps.sample_values(sto, u[None, :])[0]
        idx = np.minimum((u[:, None] >= cdf).sum(axis=1), cnt - 1)
        return vals[np.arange(s), idx]

    t_lp = t_master = 0.0
    solve_master = sd.solve_master

    def timed_master(*a, **k):
        nonlocal t_master
        t = time.perf_counter()
        out = solve_master(*a, **k)
        t_master += time.perf_counter() - t
        return out
    sd.solve_master = timed_master
    t0 = time.perf_counter()
    for it in range(1, args.iterations + 1):
        def solve(i, x, v):
            nonlocal t_lp
            a = time.perf_counter()
            out = lp.solve(x, v)
            t_lp += time.perf_counter() - a
            return out
        sd.sd_iteration_(cell, [draw()], solve, quad_scalar_schedule=sched)
        if it % max(1, args.iterations // 10) == 0:
            info = cell.improvement_info
            print(f"Iter {it:5d} lb={info.candidate_estimation:14.4f} inc_est={info.incumbent_estimation:14.4f} "
                  f"repl={info.is_improved!s:5s} dual={len(dvs):4d} cuts={len(cell.epi[0].cuts):3d} "
                  f"|x_inc - x_cand|={np.linalg.norm(cell.x_incumbent - cell.x_candidate):.3e}", flush=True)
    wall = time.perf_counter() - t0
    print(f"done: {args.iterations} iterations in {wall:.2f} s (second-stage LPs {t_lp:.2f} s, master QPs "
          f"{t_master:.2f} s, cut formation + bookkeeping through the library {wall - t_lp - t_master:.2f} s); "
          f"kernels launched: {dvs.ctx.launch_count()}")
    print("x_incumbent =", np.array2string(cell.x_incumbent, precision=4, max_line_width=120))


if __name__ == "__main__":
    main()
