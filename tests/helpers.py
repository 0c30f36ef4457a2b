"""Shared test helpers: fixture loading and counter-based synthetic workloads.

Tests may use the oracle (``oracle/``) as the checker.  Nothing here is product code.
"""
from __future__ import annotations

import json
import os

import numpy as np

from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def known_answers():
    with open(os.path.join(GOLDEN, "lands_known_answers.json")) as fh:
        return json.load(fh)


def load_instance(name):
    """Returns (oracle Problem, dict of raw arrays) for a committed real instance."""
    z = dict(np.load(os.path.join(GOLDEN, "instances", f"{name}.npz")))
    P = O.Problem(int(z["m2"]), int(z["n1"]), z["rbar"], z["T_colptr"], z["T_rowval"],
                  z["T_nzval"], z["pos_row"], z["pos_col"])
    return P, z


def sample_instance_values(z, N, seed=1):
    """Inverse-CDF sampling of the instance's discrete outcome tables."""
    s = len(z["pos_row"])
    u = O.u01(seed, np.arange(N * s, dtype=np.uint64)).reshape(N, s)
    idx = (u[:, :, None] >= z["out_cdf"][None, :, :]).sum(axis=2)
    idx = np.minimum(idx, np.maximum(z["out_cnt"][None, :] - 1, 0))
    return np.take_along_axis(np.broadcast_to(z["out_vals"], (N,) + z["out_vals"].shape),
                              idx[:, :, None], 2)[:, :, 0].copy()


def synthetic_problem(m2=64, n1=16, s=24, n_T=0, seed=6, first_stoch_row=0):
    """Storm/lands-like synthetic template: Tbar = one -1 per first-stage
    column, rbar in [100, 500) on the stochastic rows.  ``n_T`` of the ``s`` random
    elements perturb Tbar entries (a dT path no shipped instance exercises); they land on
    existing nonzeros for even e and on structural zeros for odd e."""
    rng_u = lambda sd, n: O.u01(sd, np.arange(n, dtype=np.uint64))
    rbar = np.zeros(m2)
    s_rhs = s - n_T
    rows = first_stoch_row + np.arange(s_rhs)
    rbar[rows] = 100.0 + 400.0 * rng_u(seed, s_rhs)
    # a few deterministic rows carry rhs too
    extra = np.arange(first_stoch_row + s_rhs, min(m2, first_stoch_row + s_rhs + 5))
    rbar[extra] = 10.0 + rng_u(seed + 1, len(extra))
    t_rows = (np.arange(n1) * 3 + 1) % m2
    order = np.arange(n1)
    colptr = np.arange(n1 + 1, dtype=np.int64)
    rowval = t_rows.astype(np.int64)
    nzval = -1.0 - 0.25 * rng_u(seed + 2, n1)
    pos_row = list(rows)
    pos_col = [-1] * s_rhs
    for e in range(n_T):
        col = (5 * e + 2) % n1
        if e % 2 == 0:
            row = int(t_rows[col])            # perturb a stored Tbar entry
        else:
            row = int((t_rows[col] + 7 + e) % m2)  # a structural zero of Tbar
            if row == t_rows[col]:
                row = (row + 1) % m2
        pos_row.append(row)
        pos_col.append(col)
    P = O.Problem(m2, n1, rbar, colptr, rowval, nzval, np.asarray(pos_row, dtype=np.int32),
                  np.asarray(pos_col, dtype=np.int32))
    return P


def synthetic_values(P, N, seed=1):
    """Outcome tables of 5 values = base*{0.8,...,1.2} (RHS) or Tbar-ish*(0.9+0.2u)."""
    s = P.s
    u = O.u01(seed, np.arange(N * s, dtype=np.uint64)).reshape(N, s)
    vals = np.empty((N, s))
    T = P.T_dense()
    for e in range(s):
        if P.pos_col[e] < 0:
            base = P.rbar[P.pos_row[e]]
            vals[:, e] = base * (0.8 + 0.1 * np.floor(5 * u[:, e]))
        else:
            base = T[P.pos_row[e], P.pos_col[e]]
            if base == 0.0:
                base = -0.5
            vals[:, e] = base * (0.9 + 0.2 * u[:, e])
    return vals


def synthetic_pool(m2, K, seed=2, scale=1000.0):
    """pi_kj = scale*(2u-1), distinct under the dedup rule w.p. ~1."""
    return scale * (2.0 * O.u01(seed, np.arange(K * m2, dtype=np.uint64)).reshape(K, m2) - 1.0)


def check_argmax_parity(P, values, x, pool, got_val, got_idx, rel_gap=1e-12, val_rtol=1e-10):
    """North-star parity rule: indices identical to the oracle except where the oracle's
    score of the device's pick is within rel_gap*max(|best|,1) of the oracle's best."""
    ov, oi = O.argmax_procedure(P, values, x, pool)
    got_idx = np.asarray(got_idx)
    bad = np.nonzero(oi != got_idx)[0]
    exempt = 0
    for i in bad:
        assert 0 <= got_idx[i] < len(pool), f"scenario {i}: index {got_idx[i]} out of range"
        sc, _ = O.score_pair(P, values[i], x, pool[got_idx[i]])
        tol = rel_gap * max(abs(ov[i]), 1.0)
        assert ov[i] - sc <= tol, (
            f"scenario {i}: device picked {got_idx[i]} (oracle score {sc!r}) but oracle "
            f"picked {oi[i]} (score {ov[i]!r}); gap exceeds {tol:g}")
        exempt += 1
    scale = np.maximum(np.abs(ov), 1.0)
    err = np.max(np.abs(np.asarray(got_val) - ov) / scale) if len(ov) else 0.0
    assert err <= val_rtol, f"max relative max_val error {err:g}"
    return exempt


# ---------------------------------------------------------------- SD-loop harness --------

def load_full_instance(name):
    """First-stage rows + second-stage LP data of a small instance (tools/make_golden.py --full)."""
    return dict(np.load(os.path.join(GOLDEN, "instances", f"{name}_full.npz")))


class OracleEpigraph:
    """The epigraph interface ``sqlp_b200.sd.sd_iteration_`` drives, answered by the CPU oracle:
    test infrastructure for the host logic (``-m "not gpu"``) and the lock-step checker of
    the GPU run.  Shares an ``O.DualVertexSet`` with the cell."""

    class _Coef:
        def __init__(self, s):
            self.s = s

        def scenario_values(self, scen):
            return np.asarray(scen, dtype=np.float64).reshape(self.s)

    def __init__(self, P, objective_weight, lower_bound, dual_vertices):
        self.P, self.objective_weight, self.lower_bound = P, float(objective_weight), float(lower_bound)
        self.dual_vertices = dual_vertices
        self.subproblem_coef = self._Coef(P.s)
        self.values, self.weights = [], []
        self.cuts, self.incumbent_cut = [], None

    @property
    def total_scenario_weight(self):
        tw = 0.0
        for w in self.weights:       # epigraph.jl:89, in scenario order
            tw += w
        return tw

    def add_scenarios(self, values, weights=None):
        values = np.asarray(values, dtype=np.float64).reshape(-1, self.P.s)
        for i, v in enumerate(values):
            self.values.append(v.copy())
            self.weights.append(1.0 if weights is None else float(weights[i]))

    def build_cut(self, x, forced_idx=None):
        from sqlp_b200.twosd import sdCut
        r = O.build_sasa_cut(self.P, np.asarray(self.values), np.asarray(self.weights), x,
                             self.dual_vertices.matrix(), forced_idx=forced_idx)
        return sdCut(r["alpha"], r["beta"], r["weight_mark"])

    def build_cuts2(self, x_cand, x_inc):
        return self.build_cut(x_cand), self.build_cut(x_inc)


class OraclePool(O.DualVertexSet):
    """``push`` with the product pool's return convention (inserted, slot)."""

    def push(self, v):
        before = len(self)
        super().push(np.asarray(v, dtype=np.float64))
        return len(self) > before, None


def make_cell(zf, dual_vertices, make_epigraph, x0, n_epi=1, lower_bound=0.0):
    """A cell over the full-instance fixture ``zf`` with ``n_epi`` equally weighted epigraphs."""
    from sqlp_b200 import sd
    fs = sd.FirstStage(zf["x_cost"], zf["A1"], zf["row_lower"], zf["row_upper"], zf["x_lower"], zf["x_upper"])
    cell = sd.sdCell(fs, dual_vertices)
    for _ in range(n_epi):
        sd.bind_epigraph_(cell, make_epigraph(1.0 / n_epi, lower_bound))
    cell.x_candidate[:] = x0
    cell.x_incumbent[:] = x0
    T = np.zeros((int(zf["m2"]), int(zf["n1"])))
    for j in range(int(zf["n1"])):
        for q in range(zf["T_colptr"][j], zf["T_colptr"][j + 1]):
            T[zf["T_rowval"][q], j] = zf["T_nzval"][q]
    lp = sd.Stage2LP(zf["W"], zf["cost"], zf["y_lower"], zf["y_upper"], zf["directions"], zf["rbar"], T,
                     zf["pos_row"], zf["pos_col"])
    return cell, lp
