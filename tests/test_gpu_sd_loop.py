"""BASELINE.json config C1 on the device: a whole TwoSD run on lands (~200 SD iterations, x0 =
[3, 3, 3, 3], rho = 0.1) driven by ``sqlp_b200.sd.sd_iteration_`` with the cut formation done by
the CUDA library, checked IN LOCK STEP against the CPU oracle fed the same scenarios and the same
dual vertices: identical dedup decisions, argmax within the north-star rule, cuts within 1e-10."""
import numpy as np
import pytest

from oracle import oracle as O
from sqlp_b200 import sd
from tests.helpers import load_full_instance, load_instance, make_cell, sample_instance_values, \
    check_argmax_parity
from tests.test_sd_host_cpu import lands_true_objective

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(600)
@pytest.mark.parametrize("n_epi,schedule,device_cuts,fused", [(1, "constant", False, False), (2, "adaptive", False, False),
                                                              (2, "constant", True, False), (2, "adaptive", False, True)])
def test_lands_sd_run_matches_oracle_every_iteration(n_epi, schedule, device_cuts, fused):
    from sqlp_b200 import twosd as T
    zf = load_full_instance("lands")
    P, z = load_instance("lands")
    coef = T.sdSubprobCoefficients.from_tables(P.rbar, P.T_colptr, P.T_rowval, P.T_nzval, P.pos_row, P.pos_col)
    dvs = T.sdDualVertexSet(m2=P.m2)
    cell, lp = make_cell(zf, dvs, lambda w, lb: T.sdEpigraph(coef, w, lb, dvs), np.full(4, 3.0), n_epi=n_epi)
    cell.device_cuts = device_cuts             # cut lists, incumbent test and master rows on the device
    cell.fused_step = fused                    # one library call per iteration (sqlp_cell_sd_step)
    shadow = O.DualVertexSet()                 # the oracle's pool, fed the same vertices
    seen = [[] for _ in range(n_epi)]          # scenarios of each epigraph so far
    vals = sample_instance_values(zf, 200 * n_epi, seed=42)
    checked = {"cuts": 0, "exempt": 0}

    def solve(i, x, v):
        obj, y, dual = lp.solve(x, v)
        before = len(shadow)
        shadow.push(dual)
        solve.expect.append(len(shadow) > before)
        return obj, y, dual
    solve.expect = []

    def on_cuts(i, cand, inc):
        pool = shadow.matrix()
        assert len(dvs) == len(pool)
        V = np.asarray(seen[i])
        w = np.ones(len(V))
        for x, cut in ((cell.x_candidate, cand), (cell.x_incumbent, inc)):
            mv, mi = cell.epi[i].argmax(x)
            checked["exempt"] += check_argmax_parity(P, V, x, pool, mv, mi)
            ref = O.build_sasa_cut(P, V, w, x, pool, forced_idx=mi)
            assert abs(cut.alpha - ref["alpha"]) <= 1e-10 * (abs(ref["alpha"]) + 1.0)
            assert np.max(np.abs(cut.beta - ref["beta"])) <= 1e-10 * (np.abs(ref["beta"]).max() + 1.0)
            assert cut.weight_mark == ref["weight_mark"] == float(len(V))
            checked["cuts"] += 1

    sched = sd.ConstantQuadScalarSchedule(0.1) if schedule == "constant" else sd.AdaptiveQuadScalarSchedule()
    cell.ext["quad_scalar"] = 0.1
    iters = 200 if n_epi == 1 else 100
    for it in range(iters):
        scen = [vals[it * n_epi + i] for i in range(n_epi)]
        for i in range(n_epi):
            seen[i].append(scen[i])
        k0 = len(dvs)
        sd.sd_iteration_(cell, scen, solve, quad_scalar_schedule=sched, on_cuts=on_cuts)
        assert len(dvs) == len(shadow)         # same dedup decisions, vertex for vertex
        if device_cuts:                        # the device's master rows are the host formula's bits
            for i, epi in enumerate(cell.epi):
                tw = epi.total_scenario_weight
                ref = O.cut_master_rows([(c.alpha, c.beta, c.weight_mark) for c in epi.cuts],
                                        (epi.incumbent_cut.alpha, epi.incumbent_cut.beta, epi.incumbent_cut.weight_mark),
                                        tw, epi.lower_bound)
                assert np.array_equal(cell.cut_rows[i], ref)
                assert epi.cuts_count() == (len(epi.cuts), True)
    for k in range(len(shadow)):
        assert (dvs[k] == shadow.matrix()[k]).all()
    assert checked["cuts"] == 2 * n_epi * iters
    x = cell.x_incumbent
    assert (zf["A1"] @ x >= zf["row_lower"] - 1e-7).all() and (zf["A1"] @ x <= zf["row_upper"] + 1e-7).all()
    assert abs(lands_true_objective(zf, lp, x) - 381.8533) / 381.8533 < 5e-3
