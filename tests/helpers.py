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


def load_pool(name, K):
    """A REAL dual-vertex pool of K vertices (tools/harvest_pool.py; committed under tests/golden/pools).  If the
    cache is missing it is rebuilt from the committed LP fixture -- the same recipe, the same image, so the
    same vertices."""
    d = os.path.join(GOLDEN, "pools")
    have = sorted(int(f[len(name) + 2:-4]) for f in os.listdir(d) if f.startswith(name + "_K") and f.endswith(".npz"))
    have = [k for k in have if k >= K]
    if have:          # vertices are kept in LP order: the first K of a larger harvest ARE the harvest of K
        return np.ascontiguousarray(np.load(os.path.join(d, f"{name}_K{have[0]}.npz"))["pool"][:K])
    from tools.harvest_pool import harvest
    return np.ascontiguousarray(harvest(name, K, verbose=False))


def sampled_values_at(z, seed, g):
    """Host twin of the device sampler (sqlp_epi_sample_scenarios) for the scenario ordinals ``g``:
    u = u01(seed, g * s + e), value = vals[e][min(#{c : cdf[e][c] <= u}, cnt[e] - 1)]."""
    s = len(z["pos_row"])
    g = np.asarray(g, dtype=np.uint64)
    out = np.empty((len(g), s))
    for a in range(0, len(g), 65536):                      # bounded temporaries
        gg = g[a:a + 65536]
        u = O.u01(seed, gg[:, None] * np.uint64(s) + np.arange(s, dtype=np.uint64)[None, :])
        idx = (u[:, :, None] >= z["out_cdf"][None, :, :]).sum(axis=2)
        idx = np.minimum(idx, np.maximum(z["out_cnt"][None, :] - 1, 0))
        out[a:a + 65536] = np.take_along_axis(np.broadcast_to(z["out_vals"], (len(gg),) + z["out_vals"].shape),
                                              idx[:, :, None], 2)[:, :, 0]
    return out


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


def balanced_pool(P, x, mean_values, K, seed=2, scale=50.0):
    """A synthetic pool whose vertices all score the same at (x, mean scenario) -- like LP duals of
    neighbouring bases -- so that the winner changes from scenario to scenario: random entries, then one
    entry on a deterministic row with a non-zero (rbar - Tbar x) is set to level the scores."""
    pool = synthetic_pool(P.m2, K, seed=seed, scale=scale)
    r = P.rbar.copy()
    T = P.T_dense()
    for e in range(P.s):
        if P.pos_col[e] < 0:
            r[P.pos_row[e]] = mean_values[e]
        else:
            T[P.pos_row[e], P.pos_col[e]] = mean_values[e]
    h = r - T @ x
    det = [j for j in range(P.m2) if j not in set(P.pos_row.tolist()) and abs(h[j]) > 1e-3]
    j0 = det[0]
    sc = pool @ h
    pool[:, j0] -= (sc - sc.mean()) / h[j0]
    return pool


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


# ---------------------------------------------------------------- SMPS files on the fly --

def write_smps(dirpath, name="synth", n1=5, n2=6, m1=2, m2=7, seed=11, n_rhs_elems=3, n_T_elems=2,
               continuous=False, quirks=True):
    """Write a small two-stage problem as .cor/.tim/.sto text (free format, the dialect the reference's
    readers accept) and return (paths, expected) -- ``expected`` holds the tables a correct reader must
    produce.  With ``quirks`` the files carry comment lines, an overwritten COLUMNS entry, an explicit zero
    in the Tbar block, two pairs per line, every bound type and exponent-format numbers."""
    rng = np.random.default_rng(seed)
    rows = ["OBJ"] + [f"A{i}" for i in range(m1)] + [f"S2R{i}" for i in range(m2)]
    dirs = ["N"] + list(rng.choice(list("GLE"), m1)) + list(rng.choice(list("GLE"), m2))
    cols = [f"X{j}" for j in range(n1)] + [f"Y{j}" for j in range(n2)]
    nrow, ncol, r2 = len(rows), len(cols), 1 + m1
    M = np.zeros((nrow, ncol))
    M[0, :] = np.round(rng.uniform(1, 9, ncol), 3)
    for j in range(n1):
        M[1 + rng.integers(m1), j] = np.round(rng.uniform(-3, 3), 2) or 1.0
        for i in rng.choice(m2, 2, replace=False):
            M[r2 + i, j] = np.round(rng.uniform(-2, 2), 3) or -1.0
    for j in range(n2):
        for i in rng.choice(m2, 3, replace=False):
            M[r2 + i, n1 + j] = np.round(rng.uniform(-5, 5), 3) or 2.0
    rhs = np.zeros(nrow)
    rhs[1:] = np.where(rng.random(nrow - 1) < 0.7, np.round(rng.uniform(-50, 50, nrow - 1), 1), 0.0)
    lower, upper = np.zeros(ncol), np.full(ncol, np.inf)
    lines = ["* synthetic core file", f"NAME          {name}", "ROWS"]
    lines += [f" {d}  {r}" for d, r in zip(dirs, rows)]
    lines.append("COLUMNS")
    explicit_zero = None
    for j, c in enumerate(cols):
        nz = [(i, M[i, j]) for i in range(nrow) if M[i, j] != 0.0]
        if quirks and j == 1:                      # first a wrong value, overwritten further down
            lines.append(f"    {c}    {rows[nz[0][0]]}    123.456")
        if quirks and j == 2:                      # an explicit zero where Tbar has none
            zi = next(i for i in range(r2, nrow) if M[i, j] == 0.0)
            explicit_zero = (zi, j)
            lines.append(f"    {c}\t{rows[zi]}\t0.0")
        k = 0
        while k < len(nz):
            if quirks and k + 1 < len(nz) and (j + k) % 2 == 0:
                (i0, v0), (i1, v1) = nz[k], nz[k + 1]
                lines.append(f"    {c}  {rows[i0]}  {float(v0)!r}   {rows[i1]}  {v1:.6E}")
                M[i1, j] = float(f"{v1:.6E}")
                k += 2
            else:
                lines.append(f"    {c}  {rows[nz[k][0]]}  {float(nz[k][1])!r}")
                k += 1
        if quirks and j == 3:
            lines.append("* a comment between columns")
    lines.append("RHS")
    for i in range(1, nrow):
        if rhs[i] != 0.0:
            lines.append(f"    RHS  {rows[i]}  {float(rhs[i])!r}")
    if quirks:
        lines.append("BOUNDS")
        spec = [("UP", 0, 217.0), ("LO", 1, -4.5), ("FX", 2, 3.25), ("FR", 3, None), ("MI", 4, None),
                ("PL", n1, None), ("UP", n1 + 1, 1e3)]
        for bt, j, v in spec:
            lines.append(f" {bt} BND  {cols[j]}" + ("" if v is None else f"  {float(v)!r}"))
            if bt == "UP": upper[j] = v
            elif bt == "LO": lower[j] = v
            elif bt == "FX": lower[j] = upper[j] = v
            elif bt == "FR": lower[j], upper[j] = -np.inf, np.inf
            elif bt == "MI": lower[j] = -np.inf
    lines.append("ENDATA")
    os.makedirs(dirpath, exist_ok=True)
    paths = {k: os.path.join(dirpath, f"{name}.{k}") for k in ("cor", "tim", "sto")}
    with open(paths["cor"], "w") as fh:
        fh.write("\n".join(lines) + "\n")
    with open(paths["tim"], "w") as fh:
        fh.write(f"TIME          {name}\nPERIODS       IMPLICIT\n    {cols[0]}  {rows[0]}  TIME1\n"
                 f"    {cols[n1]}  {rows[r2]}  TIME2\nENDATA\n")
    # random elements: RHS rows first, then Tbar entries (a stored one and a structural zero)
    elems = [("RHS", int(i)) for i in rng.choice(m2, n_rhs_elems, replace=False)]
    taken = set()
    for e in range(n_T_elems):
        j = int(rng.integers(n1))
        stored = [i for i in range(m2) if M[r2 + i, j] != 0.0 and (i, j) not in taken]
        empty = [i for i in range(m2) if M[r2 + i, j] == 0.0 and (i, j) not in taken]
        i = (stored if e % 2 == 0 and stored else empty)[0]
        taken.add((i, j))
        elems.append((cols[j], i))
    sto = [f"STOCH         {name}"]
    kinds, pars, tables = [], [], []
    for e, (cname, i) in enumerate(elems):
        base = rhs[r2 + i] if cname == "RHS" else M[r2 + i, cols.index(cname)]
        base = base if base != 0.0 else 1.5
        if continuous and e % 2 == 1:
            kind = "NORMAL" if e % 4 == 1 else "UNIFORM"
            a, b = (base, 0.25 * abs(base)) if kind == "NORMAL" else (base - 1.0, base + 2.0)
            sto += [f"INDEP         {kind}", f"    {cname}  S2R{i}  {float(a)!r}  {float(b)!r}"]
            kinds.append(kind); pars.append((a, b)); tables.append(None)
        else:
            n_out = 2 + e % 4
            vals = [float(np.round(base * (0.7 + 0.15 * o), 4)) for o in range(n_out)]
            pr = np.full(n_out, 1.0 / n_out)
            sto.append("INDEP         DISCRETE")
            if quirks and e == 0:
                sto.append("* outcomes of the first element")
            sto += [f"    {cname}  S2R{i}  {float(v)!r}  {float(p)!r}" for v, p in zip(vals, pr)]
            kinds.append("DISCRETE"); pars.append((0.0, 0.0)); tables.append((vals, list(pr)))
    sto.append("ENDATA")
    with open(paths["sto"], "w") as fh:
        fh.write("\n".join(sto) + "\n")
    T = M[r2:, :n1]
    expected = dict(name=name, rows=rows, cols=cols, dirs=dirs, M=M, rhs=rhs, lower=lower, upper=upper, n1=n1, n2=n2,
                    m2=m2, r2=r2, T=T, W=M[r2:, n1:], rbar=rhs[r2:], cost=M[0, n1:], x_cost=M[0, :n1],
                    pos_row=np.array([i for _, i in elems], dtype=np.int32),
                    pos_col=np.array([-1 if c == "RHS" else cols.index(c) for c, _ in elems], dtype=np.int32),
                    positions=[(c, f"S2R{i}") for c, i in elems], kinds=kinds, pars=pars, tables=tables,
                    explicit_zero=explicit_zero)
    return paths, expected


def write_smps_from_full(zf, dirpath, name):
    """Write a ``<name>_full.npz`` fixture back as .cor/.tim/.sto text, so that file-based paths (the native
    SMPS reader, ``tools/run_sd.py --smps``) can run on real instances where the reference checkout is absent."""
    n1, n2, m2 = int(zf["n1"]), int(zf["n2"]), int(zf["m2"])
    A1, W = zf["A1"], zf["W"]
    m1 = len(A1)
    T = np.zeros((m2, n1))
    for j in range(n1):
        for q in range(zf["T_colptr"][j], zf["T_colptr"][j + 1]):
            T[zf["T_rowval"][q], j] = zf["T_nzval"][q]
    rows = ["OBJ"] + [f"R1_{i}" for i in range(m1)] + [f"R2_{i}" for i in range(m2)]
    dirs1, rhs1 = [], []
    for lo, up in zip(zf["row_lower"], zf["row_upper"]):
        d = "E" if lo == up else ("G" if np.isfinite(lo) else "L")
        dirs1.append(d)
        rhs1.append(lo if d in "GE" else up)
    dirs = ["N"] + dirs1 + [str(d) for d in zf["directions"]]
    rhs = np.concatenate([[0.0], rhs1, zf["rbar"]])
    cols = [f"X{j}" for j in range(n1)] + [f"Y{j}" for j in range(n2)]
    M = np.zeros((len(rows), n1 + n2))
    M[0, :n1], M[0, n1:] = zf["x_cost"], zf["cost"]
    M[1:1 + m1, :n1] = A1
    M[1 + m1:, :n1], M[1 + m1:, n1:] = T, W
    lines = [f"NAME          {name}", "ROWS"] + [f" {d}  {r}" for d, r in zip(dirs, rows)] + ["COLUMNS"]
    for j, c in enumerate(cols):
        lines += [f"    {c}  {rows[i]}  {float(M[i, j])!r}" for i in range(len(rows)) if M[i, j] != 0.0]
    lines.append("RHS")
    lines += [f"    RHS  {rows[i]}  {float(rhs[i])!r}" for i in range(1, len(rows)) if rhs[i] != 0.0]
    lower = np.concatenate([zf["x_lower"], zf["y_lower"]])
    upper = np.concatenate([zf["x_upper"], zf["y_upper"]])
    bnd = []
    for c, lo, up in zip(cols, lower, upper):
        if np.isneginf(lo) and np.isposinf(up):
            bnd.append(f" FR BND  {c}")
            continue
        if lo != 0.0:
            bnd.append(f" MI BND  {c}" if np.isneginf(lo) else f" LO BND  {c}  {float(lo)!r}")
        if np.isfinite(up):
            bnd.append(f" UP BND  {c}  {float(up)!r}")
    if bnd:
        lines += ["BOUNDS"] + bnd
    lines.append("ENDATA")
    os.makedirs(dirpath, exist_ok=True)
    prefix = os.path.join(dirpath, name)
    with open(prefix + ".cor", "w") as fh:
        fh.write("\n".join(lines) + "\n")
    with open(prefix + ".tim", "w") as fh:
        fh.write(f"TIME          {name}\nPERIODS       IMPLICIT\n    {cols[0]}  {rows[0]}  TIME1\n"
                 f"    {cols[n1]}  {rows[1 + m1]}  TIME2\nENDATA\n")
    sto = [f"STOCH         {name}", "INDEP         DISCRETE"]
    for e in range(len(zf["pos_row"])):
        cname = "RHS" if zf["pos_col"][e] < 0 else cols[int(zf["pos_col"][e])]
        cdf = zf["out_cdf"][e]
        for o in range(int(zf["out_cnt"][e])):
            prob = zf["probs"][e][o] if "probs" in zf else cdf[o] - (cdf[o - 1] if o else 0.0)
            sto.append(f"    {cname}  R2_{int(zf['pos_row'][e])}  {float(zf['out_vals'][e, o])!r}  {float(prob)!r}")
    sto.append("ENDATA")
    with open(prefix + ".sto", "w") as fh:
        fh.write("\n".join(sto) + "\n")
    return prefix
