#!/usr/bin/env python
"""Generate the committed fixtures under tests/golden/ (run in the BUILD container only).

Reads the SMPS instances that ship with the reference (``/root/reference/spInput``; data
files, not source), re-derives the stage-2 tables with ``sqlp_b200.smps`` and harvests
real dual vertices by solving sampled stage-2 LPs with scipy's HiGHS.  Julia/GLPK are not
in the image, so where the reference's tests quote a GLPK dual vertex the script checks the
quoted NUMBERS (scores, subgradient, objectives) against HiGHS and records both.

Outputs (small, committed):
  tests/golden/instances/<name>.npz   stage-2 tables, outcome tables, a harvested pool
  tests/golden/lands_known_answers.json   the reference's own known answers for the path

Nothing under tests/, bench.py or smoke() reads /root/reference at run time; they read
these fixtures.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
from scipy.optimize import linprog

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from sqlp_b200 import smps  # noqa: E402
from oracle import oracle as O  # noqa: E402

REF = "/root/reference/spInput"
OUT = os.path.join(ROOT, "tests", "golden")


def load(name):
    d = os.path.join(REF, name)
    cor = smps.read_cor(os.path.join(d, f"{name}.cor"))
    tim = smps.read_tim(os.path.join(d, f"{name}.tim"))
    sto = smps.read_sto(os.path.join(d, f"{name}.sto"))
    return cor, tim, sto, smps.stage2_tables(cor, tim, sto)


def solve_stage2(st: smps.Stage2, x, values):
    """min c'y s.t. W y (dir) r_w - T_w x.  Returns (obj, y, dual[m2]) with JuMP's sign
    convention for a MIN problem (<= rows: dual <= 0, >= rows: dual >= 0)."""
    r = st.rbar.copy()
    T = st.T_dense()
    for e, v in enumerate(values):
        if st.pos_col[e] < 0:
            r[st.pos_row[e]] = v
        else:
            T[st.pos_row[e], st.pos_col[e]] = v
    b = r - T @ x
    dirs = np.asarray(st.directions)
    L, G, E = (dirs == "L"), (dirs == "G"), (dirs == "E")
    A_ub = np.vstack([st.W[L], -st.W[G]])
    b_ub = np.concatenate([b[L], -b[G]])
    kw = {}
    if E.any():
        kw.update(A_eq=st.W[E], b_eq=b[E])
    bounds = list(zip(st.y_lower, [None if np.isinf(u) else u for u in st.y_upper]))
    bounds = [(None if np.isinf(l) else l, u) for l, u in bounds]
    res = linprog(st.cost, A_ub=A_ub if len(b_ub) else None, b_ub=b_ub if len(b_ub) else None,
                  bounds=bounds, method="highs-ds", **kw)
    if res.status != 0:
        raise RuntimeError(f"LP failed: {res.message}")
    dual = np.zeros(st.m2)
    nL = int(L.sum())
    if len(b_ub):
        m = res.ineqlin.marginals
        dual[L] = m[:nL]
        dual[G] = -m[nL:]
    if E.any():
        dual[E] = res.eqlin.marginals
    return res.fun, res.x, dual


def expected_value_x(cor, st: smps.Stage2, sto):
    """First-stage part of the expected-value LP (used as the harvest point)."""
    mean = []
    for k, p in zip(sto.kind, sto.params):
        mean.append(float(np.dot(p[0], p[1])) if k == "DISCRETE" else
                    (p[0] if k == "NORMAL" else 0.5 * (p[0] + p[1])))
    ncol = len(cor.col_names)
    nrow = len(cor.row_names)
    A = np.zeros((nrow - 1, ncol))
    c = np.zeros(ncol)
    for (i, j), v in cor.entries.items():
        if i == 0:
            c[j] = v
        else:
            A[i - 1, j] = v
    b = cor.rhs[1:].copy()
    r2 = nrow - st.m2
    for e, v in enumerate(mean):
        if st.pos_col[e] < 0:
            b[r2 - 1 + st.pos_row[e]] = v
        else:
            A[r2 - 1 + st.pos_row[e], st.pos_col[e]] = v
    dirs = np.asarray(cor.directions[1:])
    L, G, E = (dirs == "L"), (dirs == "G"), (dirs == "E")
    kw = {}
    if E.any():
        kw.update(A_eq=A[E], b_eq=b[E])
    A_ub = np.vstack([A[L], -A[G]])
    b_ub = np.concatenate([b[L], -b[G]])
    bounds = [(None if np.isinf(l) else l, None if np.isinf(u) else u)
              for l, u in zip(cor.lower, cor.upper)]
    res = linprog(c, A_ub=A_ub, b_ub=b_ub, bounds=bounds, method="highs", **kw)
    if res.status != 0:
        raise RuntimeError(res.message)
    return res.x[:st.n1], res.fun


def discrete_tables(sto):
    """Pad the per-element outcome tables to a rectangle [s, max_outcomes]."""
    mo = max(len(p[0]) if k == "DISCRETE" else 0 for k, p in zip(sto.kind, sto.params))
    s = len(sto.positions)
    vals = np.zeros((s, max(mo, 1)))
    cdf = np.ones((s, max(mo, 1)))
    cnt = np.zeros(s, dtype=np.int32)
    for e, (k, p) in enumerate(zip(sto.kind, sto.params)):
        if k != "DISCRETE":
            continue
        n = len(p[0])
        vals[e, :n] = p[0]
        vals[e, n:] = p[0][-1]
        cdf[e, :n] = np.cumsum(p[1])
        cnt[e] = n
    return vals, cdf, cnt


def sample_discrete(vals, cdf, cnt, u):
    """values[i, e] = vals[e, #{c : cdf[e, c] <= u}] clipped to the last outcome."""
    idx = (u[:, :, None] >= cdf[None, :, :]).sum(axis=2)
    idx = np.minimum(idx, np.maximum(cnt[None, :] - 1, 0))
    return np.take_along_axis(vals[None, :, :].repeat(len(u), 0), idx[:, :, None], 2)[:, :, 0]


def make_instance(name, n_harvest):
    cor, tim, sto, st = load(name)
    s = len(st.pos_row)
    out = dict(n1=st.n1, m2=st.m2, n2=st.n2, rbar=st.rbar, T_colptr=st.T_colptr,
               T_rowval=st.T_rowval, T_nzval=st.T_nzval, pos_row=st.pos_row,
               pos_col=st.pos_col, x_lower=st.x_lower, x_upper=st.x_upper)
    if all(k == "DISCRETE" for k in sto.kind):
        vals, cdf, cnt = discrete_tables(sto)
        out.update(out_vals=vals, out_cdf=cdf, out_cnt=cnt)
        x0, obj = expected_value_x(cor, st, sto)
        # two harvest points: the expected-value solution and a perturbed one
        u = O.u01(7, np.arange(n_harvest * s)).reshape(n_harvest, s)
        values = sample_discrete(vals, cdf, cnt, u)
        span = np.where(np.isfinite(st.x_upper), st.x_upper, np.maximum(1.0, np.abs(x0))) \
            - st.x_lower
        x1 = np.clip(x0 + 0.05 * span * (O.u01(8, np.arange(st.n1)) - 0.5), st.x_lower,
                     np.where(np.isfinite(st.x_upper), st.x_upper, np.inf))
        dv = O.DualVertexSet()
        worst = 0.0
        for i in range(n_harvest):
            xx = x0 if i % 2 == 0 else x1
            try:
                objv, _, dual = solve_stage2(st, xx, values[i])
            except RuntimeError:
                continue            # perturbed point without complete recourse
            chk = O.eval_dual(problem_of(st), values[i], xx, dual)
            worst = max(worst, abs(chk - objv) / max(1.0, abs(objv)))
            dv.push(dual)
        print(f"{name}: n1={st.n1} m2={st.m2} s={s} EV obj={obj:.6f} harvested {len(dv)} "
              f"distinct of {n_harvest}; max |pi.(r-Tx)-obj| rel = {worst:.2e}")
        out.update(pool=dv.matrix(), x_ev=x0, x_alt=x1, ev_obj=obj)
    os.makedirs(os.path.join(OUT, "instances"), exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "instances", f"{name}.npz"), **out)
    return st, sto


def problem_of(st):
    return O.Problem(st.m2, st.n1, st.rbar, st.T_colptr, st.T_rowval, st.T_nzval,
                     st.pos_row, st.pos_col)


def lands_known_answers():
    """Known answers of test/sd_test.jl, test/sgd_example.jl on lands, re-derived with HiGHS."""
    cor, tim, sto, st = load("lands")
    ka = {}
    # sd_test.jl:17-23  lookups (1-based in Julia)
    ka["row_lookup_S2C5_1based"] = st.row_names.index("S2C5") + 1
    ka["col_lookup_X2_1based"] = st.x_names.index("X2") + 1
    # sd_test.jl:36-41  template at RHS=3, scenario RHS=5 -> delta_rhs[S2C5] == 2
    ka["delta_case"] = dict(template_rhs=3.0, scenario_rhs=5.0, expect=2.0)
    # sd_test.jl:45-65  eval_dual == LP objective at x=[3,3,3,3], RHS 5 and 3
    x1 = np.array([3.0, 3.0, 3.0, 3.0])
    o5, _, d5 = solve_stage2(st, x1, [5.0])
    o3, _, d3 = solve_stage2(st, x1, [3.0])
    o7, _, d7 = solve_stage2(st, x1, [7.0])
    ka["eval_dual"] = dict(x=x1.tolist(), rhs=[5.0, 3.0], obj=[o5, o3],
                           dual=[d5.tolist(), d3.tolist()])
    # sd_test.jl:69-94  argmax values at x2 equal re-solved LP objectives
    x2 = np.array([2.0, 4.0, 2.0, 6.0])
    scen = [5.0, 5.0, 3.0, 7.0]
    pool = O.DualVertexSet()
    for r in scen:
        pool.push(solve_stage2(st, x1, [r])[2])
    ka["argmax"] = dict(x1=x1.tolist(), x2=x2.tolist(), scen_rhs=scen,
                        pool=[v.tolist() for v in pool], pool_size_expected=3,
                        lp_obj_at_x2=[solve_stage2(st, x2, [r])[0] for r in scen])
    # sd_test.jl:97-103 and sgd_example.jl:22-28  subgradient at x=[2,3,4,5], RHS=7
    x = np.array([2.0, 3.0, 4.0, 5.0])
    _, _, p7 = solve_stage2(st, x, [7.0])
    ka["subgradient"] = dict(x=x.tolist(), rhs=7.0, dual=p7.tolist(),
                             expect=[-11.0, -6.0, -19.0, 0.0])
    # sd_test.jl:207-235  build_sasa_cut with weights 1.5 / 0.5.  The comments there quote
    # GLPK's vertices through their scores: scenario RHS=3: 168 vs 169 (second vertex wins),
    # scenario RHS=7: 344 vs 331 (first wins).  my_dual (RHS=5 at x1) reproduces 168/344
    # with HiGHS.  GLPK's my_dual_2 (degenerate RHS=3 LP) is reconstructed as the optimal
    # dual vertex of that LP whose S2C5 multiplier is (331-169)/4 = 40.5.
    ka["sasa"] = dict(x=x.tolist(), weights=[1.5, 0.5], scen_rhs=[3.0, 7.0],
                      my_dual=d5.tolist(), my_dual_2_highs=d3.tolist(),
                      quoted_scores=dict(scen3=[168.0, 169.0], scen7=[344.0, 331.0]),
                      weight_mark=2.0)
    d3g = glpk_like_vertex(st, x1, 3.0, 40.5)
    if d3g is not None:
        ka["sasa"]["my_dual_2_glpk_reconstructed"] = d3g.tolist()
    with open(os.path.join(OUT, "lands_known_answers.json"), "w") as fh:
        json.dump(ka, fh, indent=1)
    print("lands known answers:", json.dumps({k: ka[k] for k in ("subgradient",)}, indent=None))
    print(" d5", d5, "o5", o5, "\n d3", d3, "o3", o3, "\n d3 glpk-like", d3g)


def glpk_like_vertex(st, x, rhs, s2c5):
    """Optimal dual vertex of the stage-2 LP with the S2C5 multiplier pinned."""
    r = st.rbar.copy()
    r[st.pos_row[0]] = rhs
    b = r - st.T_dense() @ x
    dirs = np.asarray(st.directions)
    # dual LP: max b'pi  s.t. W'pi <= c, pi<=0 on L rows, pi>=0 on G rows
    bounds = [(None, 0.0) if d == "L" else ((0.0, None) if d == "G" else (None, None))
              for d in dirs]
    row = st.pos_row[0]
    bounds[row] = (s2c5, s2c5)
    res = linprog(-b, A_ub=st.W.T, b_ub=st.cost, bounds=bounds, method="highs-ds")
    if res.status != 0:
        return None
    primal, _, _ = solve_stage2(st, x, [rhs])
    if abs(-res.fun - primal) > 1e-9:
        return None
    return res.x


def make_full_instance(name):
    """Everything a host SD loop needs (first-stage rows, second-stage LP data) for a small
    instance: tests/golden/instances/<name>_full.npz (config C1, tests/test_sd_loop*.py)."""
    cor, tim, sto, st = load(name)
    r2 = len(cor.row_names) - st.m2
    A1 = np.zeros((r2 - 1, st.n1))
    for (i, j), v in cor.entries.items():
        if 1 <= i < r2 and j < st.n1:
            A1[i - 1, j] = v
        assert not (1 <= i < r2 and j >= st.n1), "first-stage row touches a second-stage column"
    dirs1 = cor.directions[1:r2]
    b1 = cor.rhs[1:r2]
    lo = np.array([b if d in "GE" else -np.inf for d, b in zip(dirs1, b1)])
    up = np.array([b if d in "LE" else np.inf for d, b in zip(dirs1, b1)])
    vals, cdf, cnt = discrete_tables(sto)
    out = dict(n1=st.n1, m2=st.m2, n2=st.n2, rbar=st.rbar, T_colptr=st.T_colptr, T_rowval=st.T_rowval,
               T_nzval=st.T_nzval, pos_row=st.pos_row, pos_col=st.pos_col, x_lower=st.x_lower,
               x_upper=st.x_upper, x_cost=st.x_cost, A1=A1, row_lower=lo, row_upper=up, W=st.W,
               cost=st.cost, y_lower=st.y_lower, y_upper=st.y_upper,
               directions=np.array(st.directions), out_vals=vals, out_cdf=cdf, out_cnt=cnt,
               probs=np.array([p[1] for p in sto.params]))
    np.savez_compressed(os.path.join(OUT, "instances", f"{name}_full.npz"), **out)
    print(f"{name}_full: first stage {A1.shape}, second stage W {st.W.shape}")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 2 and sys.argv[1] == "--full":
        for nm in sys.argv[2:]:
            make_full_instance(nm)
        sys.exit(0)
    lands_known_answers()
    make_instance("lands", 12)
    make_instance("baa99-20", 96)
    make_instance("ssn", 96)
    make_instance("storm", 96)
