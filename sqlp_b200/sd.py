"""Host side of the regularized SD loop: the CALLER of the cut-formation path.

A Python mirror of the reference's cell / iteration code, keeping its names and argument
meaning, so that a whole ``TwoSD`` run can be driven with the device library doing the cut
formation (BASELINE.json config C1):

* ``sdCell``, ``bind_epigraph_``, ``add_cut_to_master``, ``sync_cuts``
                                   -- ``src/sd_algorithm/cell.jl:4-202``, ``epigraph.jl:101-117``
* ``sdEpigraphInfo``, ``evaluate_epigraph``, ``evaluate_multi_epigraph``
                                   -- ``src/sd_algorithm/epigraph.jl:148-228``
* ``sdImprovementInfo``, ``check_improvement``   -- ``src/sd_algorithm/improvement.jl:1-49``
* ``ConstantQuadScalarSchedule``, ``AdaptiveQuadScalarSchedule``
                                   -- ``src/sd_algorithm/quad_scalar.jl:1-75``
* ``sd_iteration_``                -- ``src/sd_algorithm/algorithm.jl:39-115``

As in the reference, the master QP and the second-stage LPs are solved on the CPU (the
reference calls JuMP with GLPK / CPLEX; here HiGHS through scipy).  Nothing in this file forms
cuts: ``sd_iteration_`` calls ``add_scenario_``, ``push`` and the candidate / incumbent
``build_sasa_cut`` pair of the epigraph objects it is given (``sqlp_b200.twosd.sdEpigraph``,
i.e. the CUDA path).  The master keeps the reference's wire format: one dense row
``(discount * alpha + (1 - discount) * lb, discount * beta)`` per cut, rebuilt every iteration
(``sync_cuts!`` deletes and re-adds every constraint too).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

CUT_REMOVE_TOLERANCE = 0.001          # algorithm.jl:24
INCUMBENT_SELECTION_Q = 0.2           # improvement.jl:1


# ---------------------------------------------------------------- epigraph evaluation ---

@dataclass
class sdEpigraphInfo:
    """What is needed to evaluate the piecewise approximation (epigraph.jl:152-171)."""
    objective_weight: float
    cuts: list
    incumbent_cut: object
    total_scenario_weight: float
    lower_bound: float

    @classmethod
    def of(cls, epi):
        return cls(epi.objective_weight, list(epi.cuts), epi.incumbent_cut,
                   epi.total_scenario_weight, epi.lower_bound)


def _dot(a, b):
    return float(np.dot(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)))


def evaluate_cuts(cuts, incumbent_cut, x, total_scenario_weight, lower_bound):
    """Pointwise max of the discounted cuts, the undiscounted incumbent cut and the lower
    bound at x, without the epigraph's weight (epigraph.jl:177-203, MIN sense)."""
    best_val = lower_bound
    for cut in cuts:
        discount = cut.weight_mark / total_scenario_weight
        val = discount * (cut.alpha + _dot(cut.beta, x)) + (1 - discount) * lower_bound
        if val > best_val:
            best_val = val
    if incumbent_cut is not None:
        val = incumbent_cut.alpha + _dot(incumbent_cut.beta, x)
        if val > best_val:
            best_val = val
    return best_val


def evaluate_epigraph(epi, x):
    """``evaluate_epigraph(epi | info, x)`` including the epigraph's weight (epigraph.jl:208-220)."""
    info = epi if isinstance(epi, sdEpigraphInfo) else sdEpigraphInfo.of(epi)
    return info.objective_weight * evaluate_cuts(info.cuts, info.incumbent_cut, x,
                                                 info.total_scenario_weight, info.lower_bound)


def evaluate_multi_epigraph(v_epi, x):
    """Weighted sum over the epigraphs (epigraph.jl:225-228)."""
    return sum(evaluate_epigraph(epi, x) for epi in v_epi)


def add_cut_to_master(cut, discount, lower_bound):
    """The master row of one cut, ``eta >= new_alpha + new_beta . x`` (epigraph.jl:101-117)."""
    return discount * cut.alpha + (1 - discount) * lower_bound, discount * np.asarray(cut.beta)


# ---------------------------------------------------------------- incumbent selection ---

@dataclass
class sdImprovementInfo:
    candidate_estimation: float
    incumbent_estimation: float
    required_improvement: float
    is_improved: bool


def check_improvement(f_last, f_current, x_candidate, x_incumbent, first_stage_cost):
    """Incumbent selection (improvement.jl:19-49, MIN sense).  ``first_stage_cost(x)`` is the
    master objective without epigraph and proximal terms (``cell.objf_original``)."""
    f_cand = first_stage_cost(x_candidate)
    f_inc = first_stage_cost(x_incumbent)
    candidate_estimation = evaluate_multi_epigraph(f_current, x_candidate) + f_cand
    incumbent_estimation = evaluate_multi_epigraph(f_current, x_incumbent) + f_inc
    last_candidate_estimation = evaluate_multi_epigraph(f_last, x_candidate) + f_cand
    last_incumbent_estimation = evaluate_multi_epigraph(f_last, x_incumbent) + f_inc
    required_improvement = INCUMBENT_SELECTION_Q * (last_candidate_estimation - last_incumbent_estimation)
    req = incumbent_estimation + required_improvement
    return sdImprovementInfo(candidate_estimation, incumbent_estimation, required_improvement,
                             candidate_estimation < req)


# ---------------------------------------------------------------- proximal schedules ----

def ConstantQuadScalarSchedule(reg: float):
    def g(cell) -> float:
        return reg
    return g


def AdaptiveQuadScalarSchedule(min_quad_scalar=1e-3, max_quad_scalar=1e4, R2=0.95, R3=2.0,
                               tolerance=1e-3):
    """Trust-region-like schedule; ``cell.ext['quad_scalar']`` must be initialised and
    ``cell.ext['normDk_1']`` keeps the last step's squared length (quad_scalar.jl:16-75)."""
    def g(cell) -> float:
        if "quad_scalar" not in cell.ext:
            raise AssertionError("Quad_scalar not initialized. To use AdaptiveQuadScalarSchedule, "
                                 "set up cell.ext['quad_scalar'] first!")
        normDk = 0.0
        for xi, xc in zip(cell.x_incumbent, cell.x_candidate):
            d = xi - xc
            normDk += d * d
        if "normDk_1" not in cell.ext:
            if normDk > tolerance:
                cell.ext["normDk_1"] = normDk
            else:
                return cell.ext["quad_scalar"]
        normDk_1 = cell.ext["normDk_1"]
        if cell.improvement_info.is_improved:
            if normDk > tolerance and normDk >= R3 * normDk_1:
                cell.ext["quad_scalar"] *= (R2 * R3 * normDk_1 / normDk)
        else:
            cell.ext["quad_scalar"] /= R2
        cell.ext["quad_scalar"] = min(cell.ext["quad_scalar"], max_quad_scalar)
        cell.ext["quad_scalar"] = max(cell.ext["quad_scalar"], min_quad_scalar)
        cell.ext["normDk_1"] = normDk
        return cell.ext["quad_scalar"]
    return g


# ---------------------------------------------------------------- LP / QP back ends -----

def _highs():
    try:
        from scipy.optimize._highspy import _core as hc
    except Exception as exc:                      # pragma: no cover
        raise RuntimeError("the master QP needs HiGHS (scipy.optimize._highspy)") from exc
    return hc


_highs_scheduler_reset = False


def _single_thread_highs(h):
    """These LPs and QPs have tens of variables; on a many-core host HiGHS's task scheduler costs
    ~50x the solve (80 ms against 1.6 ms per master QP on the 100+-core GPU box).  The scheduler is
    process-global and refuses a thread count that differs from the one it was started with, so it
    is reset once and every model of this module then asks for one thread (scipy's linprog, which
    leaves the option at 0 = "whatever is running", follows)."""
    global _highs_scheduler_reset
    if not _highs_scheduler_reset:
        h.resetGlobalScheduler(True)
        _highs_scheduler_reset = True
    h.setOptionValue("threads", 1)


@dataclass
class FirstStage:
    """Root-stage template: min cost.x  s.t.  row_lower <= A x <= row_upper, x_lower <= x <= x_upper."""
    cost: np.ndarray
    A: np.ndarray
    row_lower: np.ndarray
    row_upper: np.ndarray
    x_lower: np.ndarray
    x_upper: np.ndarray


class Stage2LP:
    """``solve_problem!(prob, x, scenario)`` (smps_routines.jl:40-62): min c.y  s.t.
    W y (dir) r_w - T_w x, returning (objective, y, duals of the m2 stage rows) in JuMP's
    sign convention for a MIN problem."""

    def __init__(self, W, cost, y_lower, y_upper, directions, rbar, T_dense, pos_row, pos_col):
        self.W, self.cost = np.asarray(W, float), np.asarray(cost, float)
        self.y_lower, self.y_upper = np.asarray(y_lower, float), np.asarray(y_upper, float)
        self.dirs = np.asarray([str(d) for d in directions])
        self.rbar, self.T = np.asarray(rbar, float), np.asarray(T_dense, float)
        self.pos_row, self.pos_col = np.asarray(pos_row), np.asarray(pos_col)

    def solve(self, x, values):
        from scipy.optimize import linprog
        r, T = self.rbar.copy(), self.T
        if (self.pos_col >= 0).any():
            T = T.copy()
        for e, v in enumerate(values):
            if self.pos_col[e] < 0:
                r[self.pos_row[e]] = v
            else:
                T[self.pos_row[e], self.pos_col[e]] = v
        b = r - T @ np.asarray(x, float)
        L, G, E = (self.dirs == "L"), (self.dirs == "G"), (self.dirs == "E")
        kw = {}
        if E.any():
            kw.update(A_eq=self.W[E], b_eq=b[E])
        A_ub = np.vstack([self.W[L], -self.W[G]])
        b_ub = np.concatenate([b[L], -b[G]])
        bounds = [(None if np.isinf(lo) else lo, None if np.isinf(up) else up)
                  for lo, up in zip(self.y_lower, self.y_upper)]
        res = linprog(self.cost, A_ub=A_ub if len(b_ub) else None, b_ub=b_ub if len(b_ub) else None,
                      bounds=bounds, method="highs-ds", **kw)
        if res.status != 0:
            raise RuntimeError(f"second-stage LP failed: {res.message}")
        dual = np.zeros(len(self.dirs))
        nL = int(L.sum())
        if len(b_ub):
            m = res.ineqlin.marginals
            dual[L] = m[:nL]
            dual[G] = -m[nL:]
        if E.any():
            dual[E] = res.eqlin.marginals
        return res.fun, res.x, dual


# ---------------------------------------------------------------- the cell --------------

@dataclass
class sdCell:
    """Master problem and solver state (cell.jl:4-73).  The JuMP model is replaced by the
    first-stage tables plus the dense cut rows ``sync_cuts`` produces."""
    first_stage: FirstStage
    dual_vertices: object
    epi: list = field(default_factory=list)
    x_candidate: np.ndarray = None
    x_incumbent: np.ndarray = None
    improvement_info: sdImprovementInfo | None = None
    ext: dict = field(default_factory=dict)
    # state of the last master solve
    master_solved: bool = False
    cut_duals: list = field(default_factory=list)          # per epigraph: duals of its cut rows
    incumbent_cut_dual: list = field(default_factory=list)
    cut_rows: list = field(default_factory=list)           # per epigraph: [n_rows, 1 + n1] block
    master_objective: float = float("nan")
    # keep the cut lists on the device as well (sqlp_epi_cuts_*): the incumbent test and the master
    # rows then come from the library (rows N1 / N3); the host lists are still maintained
    device_cuts: bool = False
    fused_step: bool = False       # one library call per iteration (sqlp_cell_sd_step) instead of add / push / push / cuts

    def __post_init__(self):
        n1 = len(self.first_stage.cost)
        if self.x_candidate is None:
            self.x_candidate = np.zeros(n1)
        if self.x_incumbent is None:
            self.x_incumbent = np.zeros(n1)

    def objf_original(self, x):
        return _dot(self.first_stage.cost, x)


def bind_epigraph_(cell: sdCell, epi):
    """``bind_epigraph!(cell, epi)`` (cell.jl:97-114): one epigraph variable per epigraph, entering
    the master objective with the epigraph's weight."""
    cell.epi.append(epi)
    cell.cut_duals.append(np.zeros(0))
    cell.incumbent_cut_dual.append(None)
    cell.cut_rows.append(np.zeros((0, 1 + len(cell.first_stage.cost))))


def sync_cuts(cell: sdCell):
    """``sync_cuts!(cell)`` (cell.jl:163-202): every cut row is rebuilt with its current discount
    ``weight_mark / total_scenario_weight``; the incumbent cut, if any, is the last row and is
    not discounted."""
    n1 = len(cell.first_stage.cost)
    for i, epi in enumerate(cell.epi):
        if cell.device_cuts:
            cell.cut_rows[i] = epi.master_rows()
            continue
        rows = []
        tw = epi.total_scenario_weight
        for cut in epi.cuts:
            a, b = add_cut_to_master(cut, cut.weight_mark / tw, epi.lower_bound)
            rows.append(np.concatenate([[a], b]))
        if epi.incumbent_cut is not None:
            a, b = add_cut_to_master(epi.incumbent_cut, 1.0, epi.lower_bound)
            rows.append(np.concatenate([[a], b]))
        cell.cut_rows[i] = np.asarray(rows).reshape(len(rows), 1 + n1)


def solve_master(cell: sdCell, x0, rho: float):
    """``add_regularization!`` + ``optimize!`` (cell.jl:128-132, algorithm.jl:100-112):
    min cost.x + sum_e w_e eta_e + rho/2 |x - x0|^2 over the first-stage rows and the cut rows.

    Solved in coordinates centred on the proximal point, x = x0 + d and eta_e = eta0_e + zeta_e with
    eta0_e the value of epigraph e's rows at x0, so that the right-hand sides handed to the QP solver
    are differences of cut values instead of the cuts' absolute level (1e5..1e6 on baa99-20, where the
    uncentred model makes HiGHS's active-set QP solver fail).  Multipliers are unchanged by the shift."""
    hc = _highs()
    fs = cell.first_stage
    n1, E = len(fs.cost), len(cell.epi)
    x0 = np.asarray(x0, float)
    ncol = n1 + E
    inf = hc.kHighsInf
    rows_A, lo, up = [], [], []
    if fs.A.size:
        Ax0 = fs.A @ x0
        rows_A.append(np.hstack([fs.A, np.zeros((fs.A.shape[0], E))]))
        lo += list(fs.row_lower - Ax0)
        up += list(fs.row_upper - Ax0)
    eta0 = np.zeros(E)
    for i, block in enumerate(cell.cut_rows):              # zeta_i - beta.d >= (alpha + beta.x0) - eta0_i
        if len(block) == 0:
            continue
        at_x0 = block[:, 0] + block[:, 1:] @ x0
        eta0[i] = at_x0.max()
        R = np.zeros((len(block), ncol))
        R[:, :n1] = -block[:, 1:]
        R[:, n1 + i] = 1.0
        rows_A.append(R)
        lo += list(at_x0 - eta0[i])
        up += [np.inf] * len(block)
    A = np.vstack(rows_A) if rows_A else np.zeros((0, ncol))
    lo, up = np.asarray(lo, float), np.asarray(up, float)
    model = hc.HighsModel()
    lp = model.lp_
    lp.num_col_, lp.num_row_ = ncol, A.shape[0]
    lp.col_cost_ = np.concatenate([fs.cost, [e.objective_weight for e in cell.epi]])
    lp.col_lower_ = np.concatenate([np.where(np.isinf(fs.x_lower), -inf, fs.x_lower - x0), np.full(E, -inf)])
    lp.col_upper_ = np.concatenate([np.where(np.isinf(fs.x_upper), inf, fs.x_upper - x0), np.full(E, inf)])
    lp.row_lower_ = np.where(np.isinf(lo), -inf, lo) if len(lo) else np.zeros(0)
    lp.row_upper_ = np.where(np.isinf(up), inf, up) if len(up) else np.zeros(0)
    M = lp.a_matrix_
    M.format_ = hc.MatrixFormat.kRowwise
    M.num_col_, M.num_row_ = ncol, A.shape[0]
    A = np.where(np.abs(A) <= 1e-9, 0.0, A)                # what HiGHS would drop with a warning
    nzr, nzc = np.nonzero(A)
    M.start_ = np.concatenate([[0], np.cumsum(np.bincount(nzr, minlength=A.shape[0]))]).astype(np.int32)
    M.index_ = nzc.astype(np.int32)
    M.value_ = A[nzr, nzc].astype(np.float64)
    H = model.hessian_
    H.dim_ = ncol
    H.format_ = hc.HessianFormat.kTriangular
    H.start_ = np.concatenate([np.arange(n1 + 1), np.full(E, n1)]).astype(np.int32)
    H.index_ = np.arange(n1, dtype=np.int32)
    H.value_ = np.full(n1, float(rho))
    h = hc._Highs()
    h.setOptionValue("output_flag", False)
    _single_thread_highs(h)
    ok = (hc.HighsStatus.kOk, hc.HighsStatus.kWarning)
    if h.passModel(model) not in ok or h.run() not in ok \
            or h.getModelStatus() != hc.HighsModelStatus.kOptimal:
        cell.master_solved = False
        raise RuntimeError(f"master QP not solved: {h.getModelStatus()}")
    sol = h.getSolution()
    dv = np.asarray(sol.col_value)
    rd = np.asarray(sol.row_dual)
    at = fs.A.shape[0] if fs.A.size else 0
    for i, epi in enumerate(cell.epi):
        nb = len(cell.cut_rows[i])
        d = rd[at:at + nb]
        at += nb
        ninc = 1 if epi.incumbent_cut is not None else 0
        cell.cut_duals[i] = d[:nb - ninc].copy()
        cell.incumbent_cut_dual[i] = float(d[-1]) if ninc else None
    cell.master_solved = True
    eta = eta0 + dv[n1:]
    cell.master_objective = float(h.getInfo().objective_function_value) + _dot(fs.cost, x0) + \
        float(np.dot([e.objective_weight for e in cell.epi], eta0))
    return x0 + dv[:n1], eta


# ---------------------------------------------------------------- one SD iteration ------

def sd_iteration_(cell: sdCell, scenario_list, solve_subproblem, update_incumbent_cut=True,
                  quad_scalar_schedule=None, on_cuts=None):
    """``sd_iteration!(cell, scenario_list; ...)`` (algorithm.jl:39-115).

    ``scenario_list[i]`` is the new scenario of epigraph i (a value vector in position-table
    order, or the reference's ``[((col, row), value), ...]`` form); ``solve_subproblem(i, x, values)``
    returns ``(obj, y, dual)`` for epigraph i.  ``on_cuts(i, candidate_cut, incumbent_cut)`` is
    called after each epigraph's cut pair has been formed (tests use it to check parity).
    """
    if quad_scalar_schedule is None:
        quad_scalar_schedule = ConstantQuadScalarSchedule(0.1)
    assert len(scenario_list) == len(cell.epi)
    # solve the new scenario's subproblem at candidate and incumbent (algorithm.jl:45-55)
    # With ``cell.fused_step`` the scenarios and the dual vertices of the iteration are only collected here
    # and reach the library in ONE call further down (``sqlp_cell_sd_step``): nothing between this loop
    # and the cut formation reads the scenario store or the pool, so the order of effects is the reference's.
    fused = bool(getattr(cell, "fused_step", False)) and update_incumbent_cut and not cell.device_cuts
    step_values, step_duals = [], []
    for i, scen in enumerate(scenario_list):
        epi = cell.epi[i]
        values = epi.subproblem_coef.scenario_values(scen)
        if fused:
            step_values.append(values)
        else:
            epi.add_scenarios(values.reshape(1, -1), [1.0])
        for x in (cell.x_candidate, cell.x_incumbent):
            _, _, dual_opt = solve_subproblem(i, x, values)
            if fused:
                step_duals.append(np.asarray(dual_opt, dtype=np.float64))
            else:
                cell.dual_vertices.push(dual_opt)
    # drop the cuts whose master multiplier is (numerically) zero (algorithm.jl:57-72)
    if cell.master_solved:
        for i, epi in enumerate(cell.epi):
            duals = cell.cut_duals[i]
            assert len(duals) == len(epi.cuts)
            keep = [j for j in range(len(epi.cuts)) if not abs(duals[j]) < CUT_REMOVE_TOLERANCE]
            if cell.device_cuts and len(keep) < len(epi.cuts):
                epi.cuts_delete([j for j in range(len(epi.cuts)) if j not in set(keep)])
            epi.cuts[:] = [epi.cuts[j] for j in keep]
    epi_info_last = [sdEpigraphInfo.of(epi) for epi in cell.epi]
    # the hot path: candidate cut + regenerated incumbent cut (algorithm.jl:79-85)
    if fused:
        from .twosd import sd_step
        _, _, pairs = sd_step(cell.epi, step_values, None, np.stack(step_duals), cell.x_candidate, cell.x_incumbent)
        if on_cuts is not None:
            for i, (new_cut, inc_cut) in enumerate(pairs):
                on_cuts(i, new_cut, inc_cut)
    for i, epi in enumerate(cell.epi if not fused else ()):
        if update_incumbent_cut:
            new_cut, inc_cut = epi.build_cuts2(cell.x_candidate, cell.x_incumbent)
            epi.cuts.append(new_cut)
            epi.incumbent_cut = inc_cut
        else:
            new_cut = epi.build_cut(cell.x_candidate)
            epi.cuts.append(new_cut)
        if cell.device_cuts:
            epi.cuts_commit(update_incumbent_cut)
        if on_cuts is not None:
            on_cuts(i, new_cut, epi.incumbent_cut)
    if cell.device_cuts:
        from .twosd import check_improvement_device
        cell.improvement_info = sdImprovementInfo(*check_improvement_device(
            cell.epi, cell.x_candidate, cell.x_incumbent, cell.first_stage.cost, INCUMBENT_SELECTION_Q))
    else:
        cell.improvement_info = check_improvement(epi_info_last, cell.epi, cell.x_candidate,
                                                  cell.x_incumbent, cell.objf_original)
    rho = quad_scalar_schedule(cell)
    if cell.improvement_info.is_improved:
        cell.x_incumbent[:] = cell.x_candidate
    sync_cuts(cell)
    x_new, _ = solve_master(cell, cell.x_incumbent, rho)
    cell.x_candidate[:] = x_new
