"""Host-side mirror of the reference's cut-formation interface (``TwoSD`` module).

Julia is not in this image, so the host side above the C ABI is Python; it keeps the
reference's names, argument meaning and error behaviour for the path (trailing ``!`` is
spelled ``_``), so the parity tests read like the reference's own tests:

=====================================  ===================================================
this module                            reference (yhz0/SQLP)
=====================================  ===================================================
``sdDualVertexSet`` / ``push_``        ``src/sd_algorithm/dual_set.jl:69-93``
``sdSubprobCoefficients``              ``src/sd_algorithm/subprob.jl:4-12``
``delta_coefficients``                 ``src/sd_algorithm/subprob.jl:104-121``
``eval_dual``                          ``src/sd_algorithm/subprob.jl:128-131``
``argmax_procedure``                   ``src/sd_algorithm/subprob.jl:141-169``
``sdCut``, ``sdEpigraph``              ``src/sd_algorithm/epigraph.jl:5-61``
``add_scenario_``                      ``src/sd_algorithm/epigraph.jl:81-96``
``build_sasa_cut``                     ``src/sd_algorithm/epigraph.jl:125-146``
``build_cuts_at_candidate_and_incumbent``  the loop of ``src/sd_algorithm/algorithm.jl:79-85``
=====================================  ===================================================

Every computation happens in ``libsqlp_b200.so`` on the GPU.  The ready-to-``include``
Julia twin of this file is ``julia/TwoSDB200.jl``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import MIN_SENSE, MAX_SENSE, NoArgmaxError, SqlpError, check  # noqa: F401


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------- context ----------------

class Context:
    """One GPU (and, in a scenario-sharded job, this process's rank)."""

    def __init__(self, device: int = 0, rank: int = 0, world: int = 1, nccl_id: bytes | None = None,
                 devices=None):
        self._h = C.c_void_p()
        self.device, self.rank, self.world = device, rank, world
        L = _lib.lib()
        if devices is not None:
            # ONE process (this thread) driving several GPUs: ``sqlp_ctx_create_multi``
            dv = np.ascontiguousarray(list(devices), dtype=np.int32)
            self.device, self.rank, self.world = int(dv[0]), 0, 1
            self.n_gpus = len(dv)
            check(L.sqlp_ctx_create_multi(len(dv), _ptr(dv), C.byref(self._h)))
        elif world > 1:
            if nccl_id is None or len(nccl_id) != 128:
                raise ValueError("a 128-byte ncclUniqueId is required when world > 1")
            buf = C.create_string_buffer(nccl_id, 128)
            check(L.sqlp_ctx_create_dist(device, rank, world, buf, C.byref(self._h)))
        else:
            check(L.sqlp_ctx_create(device, C.byref(self._h)))

    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(_lib.lib().sqlp_nccl_unique_id(buf))
        return buf.raw

    def set_stream(self, cuda_stream: int | None):
        check(_lib.lib().sqlp_ctx_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def synchronize(self):
        check(_lib.lib().sqlp_ctx_synchronize(self._h))

    def launch_count(self) -> int:
        n = C.c_int64()
        check(_lib.lib().sqlp_ctx_launch_count(self._h, C.byref(n)))
        return n.value

    def timer_start(self):
        check(_lib.lib().sqlp_ctx_timer_start(self._h))

    def timer_stop(self):
        check(_lib.lib().sqlp_ctx_timer_stop(self._h))

    def timer_elapsed_ms(self) -> float:
        ms = C.c_double()
        check(_lib.lib().sqlp_ctx_timer_elapsed_ms(self._h, C.byref(ms)))
        return ms.value

    def profile(self, enable):
        """False / True: every kernel class off / on; an int > 1 is the class mask of ``sqlp_ctx_profile``
        (2 = the contraction alone)."""
        check(_lib.lib().sqlp_ctx_profile(self._h, int(enable)))

    def profile_read(self, reset=True):
        ms, n, fl = C.c_double(), C.c_int64(), C.c_double()
        check(_lib.lib().sqlp_ctx_profile_read(self._h, int(reset), C.byref(ms), C.byref(n), C.byref(fl)))
        return ms.value, n.value, fl.value

    PROFILE_CLASSES = ("contract", "delta", "reduce", "pool", "bias", "screen", "resolve", "fallback")

    def profile_classes(self, reset=True):
        """{class: (device ms, event scopes, work)}; work = flops for the contraction, algorithmic bytes else."""
        import numpy as np
        ms, n, wk = np.zeros(8), np.zeros(8, dtype=np.int64), np.zeros(8)
        check(_lib.lib().sqlp_ctx_profile_classes(self._h, int(reset), ms.ctypes.data_as(C.c_void_p),
                                                  n.ctypes.data_as(C.c_void_p), wk.ctypes.data_as(C.c_void_p)))
        return {k: (float(ms[i]), int(n[i]), float(wk[i])) for i, k in enumerate(self.PROFILE_CLASSES)}

    def set_screen(self, mode: int):
        """0: every score in FP64; 1: screening pass on the tensor cores when it pays (default); 2: always."""
        check(_lib.lib().sqlp_ctx_set_screen(self._h, int(mode)))

    def close(self):
        if self._h:
            _lib.lib().sqlp_ctx_destroy(self._h)
            self._h = C.c_void_p()


_default_ctx: dict[int, Context] = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


# ---------------------------------------------------------------- dual vertex set -------

class sdDualVertexSet:
    """Device-resident dual-vertex pool with the reference's dedup rule.

    The reference allows vectors of different lengths in one set: they never compare equal
    (dual_set.jl:26), so they only count and iterate (test/dual_set_test.jl:23-26).  The device pool
    holds vertices of ONE length ``m2`` -- the stage-2 row count, fixed by the first push or by the
    epigraph the set is bound to; a vector of another length goes to a side pool of its own length
    (same device dedup rule among its peers) and keeps its place in the insertion order.  Only
    vertices of length ``m2`` are ever scored: a vector of another length cannot be a dual of the
    epigraph's subproblem (``dot`` would throw in the reference, subprob.jl:155).
    """

    def __init__(self, data=(), ctx: Context | None = None, m2: int | None = None):
        self.ctx = ctx or default_context()
        self._h = C.c_void_p()
        self.m2 = None
        self._side: dict[int, "sdDualVertexSet"] = {}      # other lengths -> their own pool
        self._side_at: list[tuple[int, int, int]] = []     # (main-pool size when inserted, length, side slot)
        if m2 is not None:
            self._create(int(m2))
        for d in data:                       # dual_set.jl:98-104
            push_(self, d)

    def _create(self, m2):
        check(_lib.lib().sqlp_pool_create(self.ctx._h, m2, C.byref(self._h)))
        self.m2 = m2

    def push(self, v):
        """Returns (inserted, 0-based slot of v or of its stored duplicate)."""
        v = _f64(v)
        if v.ndim != 1:
            raise ValueError("a dual vertex is a vector")
        if self.m2 is None:
            self._create(len(v))
        if len(v) != self.m2:
            # dual_set.jl:26: never equal to a vertex of another length -- dedup among its own kind only
            side = self._side.get(len(v))
            if side is None:
                side = self._side[len(v)] = sdDualVertexSet(ctx=self.ctx, m2=len(v))
            ins, slot = side.push(v)
            if ins:
                self._side_at.append((self._main_len(), len(v), slot))
            pos = [q for q, (_, ln, sl) in enumerate(self._side_at) if ln == len(v) and sl == slot][0]
            return ins, self._global_index_of_side(pos)
        ins, idx = C.c_int32(), C.c_int64()
        check(_lib.lib().sqlp_pool_push(self._h, _ptr(v), C.byref(ins), C.byref(idx)))
        return bool(ins.value), self._global_index_of_main(idx.value)

    # -- insertion order over the main pool and the side pools ------------------------------
    def _main_len(self):
        K = C.c_int64()
        check(_lib.lib().sqlp_pool_size(self._h, C.byref(K)))
        return K.value

    def _global_index_of_main(self, k):
        return k + sum(1 for at, _, _ in self._side_at if at <= k)

    def _global_index_of_side(self, pos):
        return self._side_at[pos][0] + pos

    def _order(self):
        """[(length, slot)] in insertion order: side entry q sits after `at` main vertices."""
        out, q = [], 0
        for k in range(self._main_len()):
            while q < len(self._side_at) and self._side_at[q][0] <= k:
                out.append(self._side_at[q][1:])
                q += 1
            out.append((self.m2, k))
        out.extend(t[1:] for t in self._side_at[q:])
        return out

    def push_many(self, V):
        V = _f64(V)
        if self.m2 is None:
            self._create(V.shape[1])
        if V.ndim != 2 or V.shape[1] != self.m2:
            raise ValueError("expected an [n x m2] array")
        n = len(V)
        ins = np.zeros(n, dtype=np.int32)
        idx = np.zeros(n, dtype=np.int64)
        check(_lib.lib().sqlp_pool_push_batch(self._h, n, _ptr(V), _ptr(ins), _ptr(idx)))
        return ins.astype(bool), idx

    def __len__(self):                       # dual_set.jl:109-111
        if self.m2 is None:
            return 0
        return self._main_len() + len(self._side_at)

    def _main_get(self, k):
        out = np.zeros(self.m2)
        check(_lib.lib().sqlp_pool_get(self._h, int(k), _ptr(out)))
        return out

    def __getitem__(self, k):
        """Vertex ``k`` in insertion order.  Without vectors of another length (every SD run) this is
        the device slot the argmax returns."""
        if not self._side_at:
            if not 0 <= k < self._main_len():
                raise IndexError(k)
            return self._main_get(k)
        order = self._order()
        if not 0 <= k < len(order):
            raise IndexError(k)
        ln, slot = order[k]
        return self._main_get(slot) if ln == self.m2 else self._side[ln]._main_get(slot)

    def __iter__(self):                      # dual_set.jl:116-122, insertion order
        if not self._side_at:
            for k in range(self._main_len()):
                yield self._main_get(k)
            return
        for ln, slot in self._order():
            yield self._main_get(slot) if ln == self.m2 else self._side[ln]._main_get(slot)

    def hash(self, v) -> int:
        v = _f64(v)
        h = C.c_uint64()
        check(_lib.lib().sqlp_pool_hash(self._h, _ptr(v), C.byref(h)))
        return h.value

    def close(self):
        for side in self._side.values():
            side.close()
        self._side = {}
        if self._h:
            check(_lib.lib().sqlp_pool_destroy(self._h))
            self._h = C.c_void_p()


def push_(dvs: sdDualVertexSet, new_vec) -> sdDualVertexSet:
    """``Base.push!(dvs, new_vec)`` -- returns the set, like the reference (dual_set.jl:84-93)."""
    dvs.push(new_vec)
    return dvs


# ---------------------------------------------------------------- coefficients ----------

@dataclass
class sdSubprobCoefficients:
    """r, T of the second-stage template plus name -> index lookups (subprob.jl:4-12).

    ``rhs`` dense [m2]; ``transfer`` as 0-based CSC (colptr, rowval, nzval) with shape
    (m2, n1); lookups are 0-based here (1-based in Julia).  ``position_table`` resolves the
    instance's random elements once: a list of (col_name, row_name) in the order scenario
    value vectors use.
    """
    rhs: np.ndarray
    T_colptr: np.ndarray
    T_rowval: np.ndarray
    T_nzval: np.ndarray
    n1: int
    row_lookup: dict = field(default_factory=dict)
    col_lookup: dict = field(default_factory=dict)
    position_table: list = field(default_factory=list)

    def __post_init__(self):
        self.rhs = _f64(self.rhs)
        self.T_colptr = np.ascontiguousarray(self.T_colptr, dtype=np.int64)
        self.T_rowval = np.ascontiguousarray(self.T_rowval, dtype=np.int64)
        self.T_nzval = _f64(self.T_nzval)

    @property
    def m2(self):
        return len(self.rhs)

    def positions(self):
        """(pos_row, pos_col) int32 arrays of the position table; col -1 = RHS."""
        rows, cols = [], []
        for col_name, row_name in self.position_table:
            rows.append(self.row_lookup[row_name])            # KeyError like the reference
            cols.append(-1 if col_name in ("RHS", "rhs") else self.col_lookup[col_name])
        return np.asarray(rows, dtype=np.int32), np.asarray(cols, dtype=np.int32)

    @classmethod
    def from_tables(cls, rbar, T_colptr, T_rowval, T_nzval, pos_row, pos_col):
        """Index-only construction (no names): rows are 'r<i>', columns 'c<j>'."""
        m2, n1 = len(rbar), len(T_colptr) - 1
        table = [("RHS" if c < 0 else f"c{int(c)}", f"r{int(r)}") for r, c in zip(pos_row, pos_col)]
        return cls(rbar, T_colptr, T_rowval, T_nzval, n1,
                   {f"r{i}": i for i in range(m2)}, {f"c{j}": j for j in range(n1)}, table)

    def scenario_values(self, scenario):
        """Flatten a ``spSmpsScenario`` ([((col_name, row_name), value), ...]) into the value
        vector in position-table order; a plain sequence of floats passes through."""
        if len(scenario) and isinstance(scenario[0], (tuple, list)) and len(scenario[0]) == 2 \
                and isinstance(scenario[0][0], (tuple, list)):
            index = {tuple(p): e for e, p in enumerate(self.position_table)}
            vals = np.full(len(self.position_table), np.nan)
            for pos, val in scenario:
                pos = tuple(pos)
                if pos not in index:
                    # resolve through the lookups so unknown names raise KeyError
                    self.row_lookup[pos[1]]
                    if pos[0] not in ("RHS", "rhs"):
                        self.col_lookup[pos[0]]
                    raise KeyError(pos)
                vals[index[pos]] = val
            if np.isnan(vals).any():
                raise ValueError("scenario does not realise every element of the position table")
            return vals
        return _f64(scenario)


@dataclass
class sdDeltaCoefficients:
    """delta_rhs dense [m2]; delta_transfer as {(row, col): value} (subprob.jl:95-98)."""
    delta_rhs: np.ndarray
    delta_transfer: dict


@dataclass
class sdCut:
    """eta >= alpha + beta.x, never scaled (epigraph.jl:5-12)."""
    alpha: float
    beta: np.ndarray
    weight_mark: float


# ---------------------------------------------------------------- epigraph ---------------

class sdEpigraph:
    """Device half of ``sdEpigraph`` (epigraph.jl:17-61): scenario store, weights, deltas.

    The JuMP subproblem, the cut list and ``add_cut_to_master!`` stay on the host and are
    out of scope; ``cuts`` / ``incumbent_cut`` are plain Python lists kept for API parity.
    """

    def __init__(self, coef: sdSubprobCoefficients, objective_weight: float, lower_bound: float,
                 dual_vertices: sdDualVertexSet):
        self.subproblem_coef = coef
        self.objective_weight = float(objective_weight)
        self.lower_bound = float(lower_bound)
        self.dual_vertices = dual_vertices
        self.ctx = dual_vertices.ctx
        self.cuts: list[sdCut] = []
        self.incumbent_cut: sdCut | None = None
        if dual_vertices.m2 is None:
            dual_vertices._create(coef.m2)
        pos_row, pos_col = coef.positions()
        self.s = len(pos_row)
        nz = np.nonzero(coef.rhs)[0].astype(np.int64)
        rv = coef.rhs[nz].copy()
        self._h = C.c_void_p()
        check(_lib.lib().sqlp_epi_create(
            self.ctx._h, dual_vertices._h, coef.m2, coef.n1, len(nz), _ptr(nz), _ptr(rv),
            _ptr(coef.T_colptr), _ptr(coef.T_rowval), _ptr(coef.T_nzval), self.s,
            _ptr(pos_row), _ptr(pos_col), C.byref(self._h)))
        self.scenario_delta = _DeviceDeltaSet(self)
        check(_lib.lib().sqlp_epi_set_weights(self._h, self.objective_weight, self.lower_bound))

    @classmethod
    def from_smps(cls, native, objective_weight: float, lower_bound: float, dual_vertices: sdDualVertexSet):
        """``sdEpigraph(prob, w, lb)`` built by the library straight from parsed SMPS files
        (``smps.NativeSmps``; ``sqlp_epi_create_smps``): coefficient tables, position table, outcome
        tables and distributions never pass through the host language."""
        self = cls.__new__(cls)
        st = native.stage2()
        sto = native.sto()
        self.subproblem_coef = sdSubprobCoefficients(
            st.rbar, st.T_colptr, st.T_rowval, st.T_nzval, st.n1, {r: i for i, r in enumerate(st.row_names)},
            {c: j for j, c in enumerate(st.x_names)}, list(sto.positions))
        self.objective_weight, self.lower_bound = float(objective_weight), float(lower_bound)
        self.dual_vertices, self.ctx = dual_vertices, dual_vertices.ctx
        self.cuts, self.incumbent_cut = [], None
        if dual_vertices.m2 is None:
            dual_vertices._create(st.m2)
        self.s = len(st.pos_row)
        self._h = C.c_void_p()
        check(_lib.lib().sqlp_epi_create_smps(self.ctx._h, dual_vertices._h, native._h, C.byref(self._h)))
        self.scenario_delta = _DeviceDeltaSet(self)
        check(_lib.lib().sqlp_epi_set_weights(self._h, self.objective_weight, self.lower_bound))
        return self

    # -- scenario store ---------------------------------------------------------------
    def add_scenarios(self, values, weights=None):
        values = _f64(values).reshape(-1, self.s) if self.s else np.zeros((len(values), 0))
        w = None if weights is None else _f64(weights)
        if w is not None and len(w) != len(values):
            raise ValueError("one weight per scenario")
        check(_lib.lib().sqlp_epi_add_scenarios(self._h, len(values), _ptr(values), _ptr(w)))

    def set_outcomes(self, vals, cdf, cnt):
        vals, cdf = _f64(vals), _f64(cdf)
        cnt = np.ascontiguousarray(cnt, dtype=np.int32)
        check(_lib.lib().sqlp_epi_set_outcomes(self._h, vals.shape[1], _ptr(vals), _ptr(cdf), _ptr(cnt)))

    def set_distributions(self, kind, par_a, par_b):
        """kind[e]: 0 DISCRETE (outcome tables), 1 NORMAL(mean, variance), 2 UNIFORM(left, right)."""
        kind = np.ascontiguousarray(kind, dtype=np.int32)
        check(_lib.lib().sqlp_epi_set_distributions(self._h, _ptr(kind), _ptr(_f64(par_a)), _ptr(_f64(par_b))))

    def sample_scenarios(self, n_new, seed, weight_seed=0):
        check(_lib.lib().sqlp_epi_sample_scenarios(self._h, int(n_new), int(seed), int(weight_seed)))

    def counts(self):
        ng, nl, tw = C.c_int64(), C.c_int64(), C.c_double()
        check(_lib.lib().sqlp_epi_counts(self._h, C.byref(ng), C.byref(nl), C.byref(tw)))
        return ng.value, nl.value, tw.value

    @property
    def total_scenario_weight(self):
        return self.counts()[2]

    def __len__(self):
        return self.counts()[0]

    def delta(self, local_scen: int) -> sdDeltaCoefficients:
        drhs = np.zeros(self.subproblem_coef.m2)
        dT = np.zeros(max(self.s, 1))
        check(_lib.lib().sqlp_epi_delta(self._h, int(local_scen), _ptr(drhs), _ptr(dT)))
        pr, pc = self.subproblem_coef.positions()
        return sdDeltaCoefficients(drhs, {(int(r), int(c)): float(dT[e])
                                          for e, (r, c) in enumerate(zip(pr, pc)) if c >= 0})

    # -- hot path -------------------------------------------------------------------------
    def argmax(self, x, sense=MIN_SENSE):
        x = _f64(x)
        self._check_x(x)
        n = self.counts()[1]
        mv = np.zeros(n)
        mi = np.zeros(n, dtype=np.int64)
        check(_lib.lib().sqlp_epi_argmax(self._h, _ptr(x), int(sense), _ptr(mv), _ptr(mi)))
        return mv, mi

    def build_cut(self, x, with_val=False):
        x = _f64(x)
        self._check_x(x)
        alpha, wm, val = C.c_double(), C.c_double(), C.c_double()
        beta = np.zeros(self.subproblem_coef.n1)
        check(_lib.lib().sqlp_epi_build_cut(self._h, _ptr(x), C.byref(alpha), _ptr(beta),
                                            C.byref(wm), C.byref(val)))
        cut = sdCut(alpha.value, beta, wm.value)
        return (cut, val.value) if with_val else cut

    def build_cuts2(self, x_cand, x_inc, with_val=False):
        xc, xi = _f64(x_cand), _f64(x_inc)
        self._check_x(xc); self._check_x(xi)
        n1 = self.subproblem_coef.n1
        alpha = np.zeros(2); beta = np.zeros((2, n1)); val = np.zeros(2)
        wm = C.c_double()
        check(_lib.lib().sqlp_epi_build_cuts2(self._h, _ptr(xc), _ptr(xi), _ptr(alpha), _ptr(beta),
                                              C.byref(wm), _ptr(val)))
        cuts = (sdCut(alpha[0], beta[0].copy(), wm.value), sdCut(alpha[1], beta[1].copy(), wm.value))
        return (cuts, val) if with_val else cuts

    # -- the cut list on the device -----------------------------------
    def cuts_push(self, cut: sdCut):
        check(_lib.lib().sqlp_epi_cuts_push(self._h, float(cut.alpha), _ptr(_f64(cut.beta)), float(cut.weight_mark)))

    def cuts_set_incumbent(self, cut: sdCut | None):
        if cut is None:
            check(_lib.lib().sqlp_epi_cuts_set_incumbent(self._h, 0.0, None, 0.0))
        else:
            check(_lib.lib().sqlp_epi_cuts_set_incumbent(self._h, float(cut.alpha), _ptr(_f64(cut.beta)),
                                                         float(cut.weight_mark)))

    def cuts_commit(self, with_incumbent=True):
        """After a cut formation: snapshot the list, push the candidate cut and (optionally) replace
        the incumbent cut, all on the device (algorithm.jl:76-84)."""
        check(_lib.lib().sqlp_epi_cuts_commit(self._h, int(bool(with_incumbent))))

    def cuts_delete(self, idx):
        idx = np.ascontiguousarray(sorted(int(i) for i in idx), dtype=np.int64)
        check(_lib.lib().sqlp_epi_cuts_delete(self._h, len(idx), _ptr(idx)))

    def cuts_count(self):
        n, inc = C.c_int64(), C.c_int32()
        check(_lib.lib().sqlp_epi_cuts_count(self._h, C.byref(n), C.byref(inc)))
        return n.value, bool(inc.value)

    def cuts_get(self, index: int) -> sdCut:
        """Cut ``index`` of the device list; -1 is the incumbent cut."""
        a, wm = C.c_double(), C.c_double()
        beta = np.zeros(self.subproblem_coef.n1)
        check(_lib.lib().sqlp_epi_cuts_get(self._h, int(index), C.byref(a), _ptr(beta), C.byref(wm)))
        return sdCut(a.value, beta, wm.value)

    def evaluate(self, x, snapshot=False) -> float:
        """``evaluate_epigraph(epi, x)`` on the device list (``snapshot``: the list as it was before
        the last commit, i.e. the reference's ``sdEpigraphInfo`` f_{k-1})."""
        x = _f64(x)
        self._check_x(x)
        out = C.c_double()
        check(_lib.lib().sqlp_epi_evaluate(self._h, _ptr(x), int(bool(snapshot)), C.byref(out)))
        return out.value

    def master_rows(self) -> np.ndarray:
        """The dense ``[n_rows, 1 + n1]`` block ``sync_cuts!`` adds to the master."""
        n = C.c_int64()
        check(_lib.lib().sqlp_epi_master_rows(self._h, None, C.byref(n)))
        rows = np.zeros((n.value, self.subproblem_coef.n1 + 1))
        if n.value:
            check(_lib.lib().sqlp_epi_master_rows(self._h, _ptr(rows), C.byref(n)))
        return rows

    def view_columns(self):
        """(classes of score-equivalent vertices the sweep visits, relevant rows) -- ``sqlp_epi_view_columns``."""
        kv, nr = C.c_int64(), C.c_int64()
        check(_lib.lib().sqlp_epi_view_columns(self._h, C.byref(kv), C.byref(nr)))
        return kv.value, nr.value

    def screen_stats(self) -> dict:
        out = np.zeros(8, dtype=np.int64)
        check(_lib.lib().sqlp_epi_screen_stats(self._h, _ptr(out)))
        keys = ("passes", "fallbacks", "emitted", "evaluated", "overflowed_lists", "bad_operands", "live_0", "live_1")
        d = {k: int(v) for k, v in zip(keys, out)}
        d["unprofitable_passes"] = d["bad_operands"] // 2       # succeeded, but slower than the FP64 sweep would be
        d["bad_operands"] &= 1
        return d

    def eval_dual(self, local_scen, vertex, x):
        x = _f64(x)
        self._check_x(x)
        out = C.c_double()
        check(_lib.lib().sqlp_eval_dual(self._h, int(local_scen), int(vertex), _ptr(x), C.byref(out)))
        return out.value

    def _check_x(self, x):
        if x.shape != (self.subproblem_coef.n1,):
            raise ValueError(f"x has shape {x.shape}, expected ({self.subproblem_coef.n1},)")

    def close(self):
        if self._h:
            _lib.lib().sqlp_epi_destroy(self._h)
            self._h = C.c_void_p()


class _DeviceDeltaSet:
    """Stands in for ``epi.scenario_delta::Vector{sdDeltaCoefficients}`` (epigraph.jl:40):
    the deltas live on the device; indexing reads one back."""

    def __init__(self, epi: sdEpigraph):
        self.epi = epi

    def __len__(self):
        return self.epi.counts()[1]

    def __getitem__(self, i):
        return self.epi.delta(i)


# ---------------------------------------------------------------- reference functions ---

def add_scenario_(epi: sdEpigraph, scenario, weight: float = 1.0):
    """``add_scenario!(epi, scenario, weight)`` (epigraph.jl:81-96)."""
    vals = epi.subproblem_coef.scenario_values(scenario)
    epi.add_scenarios(vals.reshape(1, -1), [float(weight)])


def delta_coefficients(coef: sdSubprobCoefficients, scenario, ctx: Context | None = None) -> sdDeltaCoefficients:
    """``delta_coefficients(coef, scenario)`` (subprob.jl:104-121), computed on the device
    through a scratch epigraph."""
    dvs = sdDualVertexSet(ctx=ctx, m2=coef.m2)
    epi = sdEpigraph(coef, 1.0, 0.0, dvs)
    try:
        add_scenario_(epi, scenario)
        return epi.delta(0)
    finally:
        epi.close()
        dvs.close()


def eval_dual(coef: sdSubprobCoefficients, scenario, x, dual, ctx: Context | None = None) -> float:
    """``eval_dual(coef, delta, x, dual)`` (subprob.jl:128-131) for one scenario and one
    vertex, evaluated on the device in the reference's operation order."""
    dvs = sdDualVertexSet(ctx=ctx, m2=coef.m2)
    epi = sdEpigraph(coef, 1.0, 0.0, dvs)
    try:
        add_scenario_(epi, scenario)
        _, k = dvs.push(dual)
        return epi.eval_dual(0, k, x)
    finally:
        epi.close()
        dvs.close()


def argmax_procedure(coef, delta_set, x, dual_vertices: sdDualVertexSet, sense=MIN_SENSE):
    """``argmax_procedure(coef, delta_set, x, dual_vertices; sense)`` (subprob.jl:141-169).

    ``delta_set`` is ``epi.scenario_delta``.  Returns ``(max_val, max_arg)`` where
    ``max_arg[i]`` is the winning vertex (the reference returns Refs aliasing pool vectors;
    here they are rebuilt from the returned pool slots) and ``max_arg.index`` keeps the
    0-based slots."""
    if not isinstance(delta_set, _DeviceDeltaSet):
        raise TypeError("delta_set must be an epigraph's scenario_delta")
    epi = delta_set.epi
    if epi.subproblem_coef is not coef or epi.dual_vertices is not dual_vertices:
        raise ValueError("coef / dual_vertices do not belong to this delta set's epigraph")
    mv, mi = epi.argmax(x, sense)
    if (mi < 0).any():
        raise NoArgmaxError(_lib.E_NO_ARGMAX, "undefined reference: no vertex beat -Inf")
    cache = {}
    args = _ArgList(cache.setdefault(int(k), dual_vertices[int(k)]) for k in mi)
    args.index = mi
    return mv, args


class _ArgList(list):
    index = None


def build_sasa_cut(epi: sdEpigraph, x, dual_vertices: sdDualVertexSet | None = None) -> sdCut:
    """``build_sasa_cut(epi, x, dual_vertices)::sdCut`` (epigraph.jl:125-146)."""
    if dual_vertices is not None and dual_vertices is not epi.dual_vertices:
        raise ValueError("dual_vertices is not the pool this epigraph was bound to")
    return epi.build_cut(x)


def build_cuts_at_candidate_and_incumbent(epis, x_candidate, x_incumbent):
    """The cut-generation loop of ``sd_iteration!`` (algorithm.jl:79-85) for every epigraph
    of a cell in one library call: pushes the candidate cut to ``epi.cuts`` and replaces
    ``epi.incumbent_cut``."""
    epis = list(epis)
    if not epis:
        return []
    xc, xi = _f64(x_candidate), _f64(x_incumbent)
    n1 = epis[0].subproblem_coef.n1
    for e in epis:
        e._check_x(xc); e._check_x(xi)
        if e.subproblem_coef.n1 != n1:
            raise ValueError("epigraphs of one cell share the first stage")
    E = len(epis)
    handles = (C.c_void_p * E)(*[e._h for e in epis])
    alpha = np.zeros((E, 2)); beta = np.zeros((E, 2, n1)); wm = np.zeros(E); val = np.zeros((E, 2))
    check(_lib.lib().sqlp_cell_build_cuts2(E, handles, _ptr(xc), _ptr(xi), _ptr(alpha), _ptr(beta),
                                           _ptr(wm), _ptr(val)))
    out = []
    for i, e in enumerate(epis):
        cand = sdCut(alpha[i, 0], beta[i, 0].copy(), wm[i])
        inc = sdCut(alpha[i, 1], beta[i, 1].copy(), wm[i])
        e.cuts.append(cand)
        e.incumbent_cut = inc
        out.append((cand, inc))
    return out


def sd_step(epis, scenarios, weights, vertices, x_candidate, x_incumbent):
    """The cut formation of one ``sd_iteration!`` (algorithm.jl:45-55, 79-85) in one library call and one
    synchronisation (``sqlp_cell_sd_step``): ``add_scenario!`` of one scenario per epigraph, ``push!`` of
    the dual vertices found at the candidate and the incumbent (rows of ``vertices``, in order), then both
    cuts of every epigraph, which also go to ``epi.cuts`` / ``epi.incumbent_cut``.
    Returns (inserted [n_vertices] bool, index [n_vertices], [(candidate cut, incumbent cut), ...])."""
    epis = list(epis)
    E = len(epis)
    if E == 0:
        raise ValueError("a cell has at least one epigraph")
    if len(scenarios) != E:
        raise ValueError("one scenario per epigraph")
    dvs = epis[0].dual_vertices
    xc, xi = _f64(x_candidate), _f64(x_incumbent)
    n1 = epis[0].subproblem_coef.n1
    for e in epis:
        e._check_x(xc); e._check_x(xi)
        if e.subproblem_coef.n1 != n1 or e.dual_vertices is not dvs:
            raise ValueError("epigraphs of one cell share the first stage and the dual-vertex set")
    vals = [e.subproblem_coef.scenario_values(sc) for e, sc in zip(epis, scenarios)]
    for e, v in zip(epis, vals):
        if len(v) != e.s:
            raise ValueError("scenario does not match the epigraph's position table")
    flat = _f64(np.concatenate(vals)) if vals and sum(len(v) for v in vals) else np.zeros(1)
    w = None if weights is None else _f64(weights)
    if w is not None and len(w) != E:
        raise ValueError("one weight per epigraph")
    V = _f64(vertices).reshape(-1, dvs.m2)
    nv = len(V)
    ins = np.zeros(max(nv, 1), dtype=np.int32); idx = np.zeros(max(nv, 1), dtype=np.int64)
    handles = (C.c_void_p * E)(*[e._h for e in epis])
    alpha = np.zeros((E, 2)); beta = np.zeros((E, 2, n1)); wm = np.zeros(E); val = np.zeros((E, 2))
    check(_lib.lib().sqlp_cell_sd_step(E, handles, _ptr(flat), _ptr(w), nv, _ptr(V), _ptr(ins), _ptr(idx),
                                       _ptr(xc), _ptr(xi), _ptr(alpha), _ptr(beta), _ptr(wm), _ptr(val)))
    out = []
    for i, e in enumerate(epis):
        cand = sdCut(alpha[i, 0], beta[i, 0].copy(), wm[i])
        inc = sdCut(alpha[i, 1], beta[i, 1].copy(), wm[i])
        e.cuts.append(cand)
        e.incumbent_cut = inc
        out.append((cand, inc))
    return ins[:nv].astype(bool), idx[:nv], out


def check_improvement_device(epis, x_candidate, x_incumbent, cost, q_factor=0.2):
    """``check_improvement`` (improvement.jl:19-49) on the device cut lists of a cell's epigraphs.
    Returns (candidate_estimation, incumbent_estimation, required_improvement, is_improved)."""
    epis = list(epis)
    E = len(epis)
    handles = (C.c_void_p * E)(*[e._h for e in epis])
    out = np.zeros(4)
    check(_lib.lib().sqlp_cell_check_improvement(E, handles, _ptr(_f64(x_candidate)), _ptr(_f64(x_incumbent)),
                                                 _ptr(_f64(cost)), float(q_factor), _ptr(out)))
    return float(out[0]), float(out[1]), float(out[2]), bool(out[3])
