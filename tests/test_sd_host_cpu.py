"""Host side of the SD loop (sqlp_b200/sd.py): the reference's known answers for the cut
master rows and the epigraph evaluation (test/sd_test.jl:150-194), incumbent selection, the
proximal schedules, and a whole lands run with the cut formation answered by the CPU oracle
(BASELINE.json config C1 on the CPU; tests/test_gpu_sd_loop.py is the same run on the device)."""
import numpy as np
import pytest

from sqlp_b200 import sd
from sqlp_b200.twosd import sdCut
from tests.helpers import load_full_instance, load_instance, OracleEpigraph, OraclePool, make_cell, \
    sample_instance_values


class _Epi:
    def __init__(self, w, lb, tw):
        self.objective_weight, self.lower_bound, self.total_scenario_weight = w, lb, tw
        self.cuts, self.incumbent_cut = [], None


def _demo_cell():
    """The cell of test/sd_test.jl:105-176: two epigraphs of weight 0.5 (lower bounds 0 and
    100), two scenarios of weight 1 each, demo cuts pushed by hand."""
    fs = sd.FirstStage(np.array([10.0, 7, 16, 6]), np.zeros((0, 4)), np.zeros(0), np.zeros(0),
                       np.zeros(4), np.full(4, np.inf))
    cell = sd.sdCell(fs, None)
    e1, e2 = _Epi(0.5, 0.0, 2.0), _Epi(0.5, 100.0, 2.0)
    sd.bind_epigraph_(cell, e1)
    sd.bind_epigraph_(cell, e2)
    cut1 = sdCut(1.0, np.array([2.0, 3, 4, 5]), 1.0)
    cut2 = sdCut(6.0, np.array([7.0, 8, 9, 10]), 2.0)
    inc = sdCut(11.0, np.array([12.0, 13, 14, 15]), 1.0)
    e1.cuts += [cut1, cut2]
    e1.incumbent_cut = inc
    e2.cuts.append(cut1)
    return cell, e1, e2


def test_add_cut_to_master_row():                       # sd_test.jl:153-157
    a, b = sd.add_cut_to_master(sdCut(1.0, np.array([2.0, 3, 4, 5]), 0.1), 1.0, 0.0)
    assert a == 1.0 and (b == [2, 3, 4, 5]).all()


def test_sync_cuts_rows_and_discounted_lower_bound():   # sd_test.jl:166-188
    cell, e1, e2 = _demo_cell()
    sd.sync_cuts(cell)
    assert cell.cut_rows[0].shape == (3, 5)              # two cuts + the incumbent cut
    sd.sync_cuts(cell)
    assert cell.cut_rows[0].shape == (3, 5)              # no redundant rows
    assert cell.cut_rows[1][0, 0] == 50.5                # 100 * 0.5 + 1.0 * 0.5
    assert (cell.cut_rows[1][0, 1:] == 0.5 * np.array([2.0, 3, 4, 5])).all()
    assert (cell.cut_rows[0][2] == [11, 12, 13, 14, 15]).all()     # incumbent: no discount


def test_evaluate_epigraph_known_answers():             # sd_test.jl:190-194
    cell, e1, e2 = _demo_cell()
    x10 = np.full(4, 10.0)
    assert sd.evaluate_epigraph(e1, x10) == 551.0 * 0.5
    assert sd.evaluate_epigraph(e2, x10) == (141 / 2 + 100 / 2) * 0.5
    assert sd.evaluate_epigraph(e2, -np.ones(4)) == 100.0 * 0.5
    assert sd.evaluate_multi_epigraph(cell.epi, x10) == 551.0 * 0.5 + (141 / 2 + 100 / 2) * 0.5


def test_epigraph_info_is_a_snapshot():                 # sd_test.jl:199-205
    _, e1, _ = _demo_cell()
    info = sd.sdEpigraphInfo.of(e1)
    e1.cuts.clear()
    assert len(info.cuts) == 2


def test_check_improvement_rule():                      # improvement.jl:19-49
    cell, e1, e2 = _demo_cell()
    last = [sd.sdEpigraphInfo.of(e) for e in cell.epi]
    e1.cuts.append(sdCut(700.0, np.zeros(4), 2.0))      # a new cut that lifts the estimate everywhere
    xc, xi = np.full(4, 1.0), np.full(4, 2.0)
    info = sd.check_improvement(last, cell.epi, xc, xi, cell.objf_original)
    f = lambda es, x: sd.evaluate_multi_epigraph(es, x) + 39.0 * x[0]
    assert info.candidate_estimation == f(cell.epi, xc)
    assert info.incumbent_estimation == f(cell.epi, xi)
    assert info.required_improvement == 0.2 * (f(last, xc) - f(last, xi))
    assert info.is_improved == (info.candidate_estimation < info.incumbent_estimation + info.required_improvement)


def test_adaptive_quad_scalar_schedule():               # quad_scalar.jl:16-75
    cell, _, _ = _demo_cell()
    g = sd.AdaptiveQuadScalarSchedule()
    with pytest.raises(AssertionError):
        g(cell)
    cell.ext["quad_scalar"] = 1.0
    cell.x_incumbent[:] = 0.0
    cell.x_candidate[:] = 0.0
    cell.improvement_info = sd.sdImprovementInfo(0, 0, 0, False)
    assert g(cell) == 1.0 and "normDk_1" not in cell.ext         # no movement yet
    cell.x_candidate[:] = 1.0                                     # |d|^2 = 4
    assert g(cell) == 1.0 / 0.95 and cell.ext["normDk_1"] == 4.0  # not improved: divide by R2
    cell.improvement_info = sd.sdImprovementInfo(0, 0, 0, True)
    cell.x_candidate[:] = 2.0                                     # |d|^2 = 16 >= 2 * 4
    assert g(cell) == pytest.approx((1.0 / 0.95) * 0.95 * 2.0 * 4.0 / 16.0)
    assert sd.ConstantQuadScalarSchedule(0.1)(cell) == 0.1


def test_master_qp_small_case():
    """min x + eta + rho/2 (x - 1)^2  s.t.  eta >= 2 - x, eta >= x / 2, 0 <= x <= 10."""
    fs = sd.FirstStage(np.array([1.0]), np.zeros((0, 1)), np.zeros(0), np.zeros(0), np.zeros(1), np.array([10.0]))
    cell = sd.sdCell(fs, None)
    e = _Epi(1.0, -50.0, 1.0)
    e.cuts += [sdCut(2.0, np.array([-1.0]), 1.0), sdCut(0.0, np.array([0.5]), 1.0)]
    sd.bind_epigraph_(cell, e)
    sd.sync_cuts(cell)
    x, eta = sd.solve_master(cell, np.array([1.0]), 0.1)
    assert x[0] == pytest.approx(1.0, abs=1e-7) and eta[0] == pytest.approx(1.0, abs=1e-7)
    assert abs(cell.cut_duals[0][0]) > 0.5 and abs(cell.cut_duals[0][1]) < 1e-6


def lands_true_objective(zf, lp, x):
    """cost.x + sum_w p_w Q(x, w) over the three outcomes of the demand (the deterministic
    equivalent whose optimum test/crash_test.jl:37 quotes as 381.8533...)."""
    return float(zf["x_cost"] @ x + sum(p * lp.solve(x, [v])[0]
                                         for v, p in zip(zf["out_vals"][0], zf["probs"][0])))


def lands_sampler(zf, seed):
    vals = sample_instance_values(zf, 400, seed=seed)
    return lambda it: [vals[it]]


@pytest.mark.timeout(300)
def test_lands_sd_run_with_oracle_cut_formation():
    """~120 SD iterations on lands (x0 = [3, 3, 3, 3], rho = 0.1, lb = 0): the incumbent stays
    first-stage feasible and its true cost approaches the optimum of the 3-scenario
    deterministic equivalent (381.8533..., test/crash_test.jl:37)."""
    zf = load_full_instance("lands")
    P, _ = load_instance("lands")
    pool = OraclePool()
    cell, lp = make_cell(zf, pool, lambda w, lb: OracleEpigraph(P, w, lb, pool), np.full(4, 3.0))
    draw = lands_sampler(zf, 42)
    for it in range(120):
        sd.sd_iteration_(cell, draw(it), lambda i, x, v: lp.solve(x, v))
    x = cell.x_incumbent
    assert (zf["A1"] @ x >= zf["row_lower"] - 1e-7).all() and (zf["A1"] @ x <= zf["row_upper"] + 1e-7).all()
    assert (x >= -1e-9).all()
    assert abs(lands_true_objective(zf, lp, x) - 381.8533) / 381.8533 < 2e-3
    # the estimate is the sample-average value at the incumbent: off by the sampling error only
    assert abs(cell.improvement_info.incumbent_estimation - 381.8533) / 381.8533 < 0.06
    assert 3 <= len(pool) <= 30 and len(cell.epi[0].cuts) <= 8
