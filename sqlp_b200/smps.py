"""SMPS (.cor/.tim/.sto) reader producing the flat tables the device path consumes.

Host-side mirror of the reference's readers, restricted to what the cut-formation path
needs:

* ``read_cor``  -- reference ``src/smps/smps_cor.jl:26-194`` (sections NAME/ROWS/COLUMNS/
  RHS/BOUNDS, ``*`` comments, first row must be ``N``).
* ``read_tim``  -- ``src/smps/smps_tim.jl:30-64`` (implicit PERIODS).
* ``read_sto``  -- ``src/smps/smps_sto.jl:41-111`` (INDEP DISCRETE/NORMAL/UNIFORM).
* ``stage2_tables`` -- the stage split of ``src/smps/smps_prob.jl:14-102`` followed by the
  coefficient extraction of ``src/sd_algorithm/subprob.jl:15-69``: rbar, Tbar (CSC), W,
  second-stage costs/bounds and the stochastic-position table (row, col | -1 for RHS).

The reference keeps ``sto.indep`` in a ``Dict`` whose order is hash order
(``smps_sto.jl:35,140-149``); here the position table is fixed once, in order of first
appearance in the ``.sto`` file.

Everything here is host logic in numpy; nothing in this file computes cuts.

``NativeSmps`` binds the same reader written in C++ inside ``libsqlp_b200.so``
(``sqlp_smps_*``, ``csrc/host_smps.cuh``): the library builds its device tables straight from the
three files (``sqlp_epi_create_smps``), so a Julia host needs no JuMP coefficient extraction.  The
numpy reader above stays as the independent restatement the native one is tested against.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


def _data_lines(path):
    """(is_header, tokens) for every non-empty, non-comment line."""
    with open(path, "r") as fh:
        for raw in fh.read().splitlines():
            if not raw.strip() or raw[0] == "*":
                continue
            yield (not raw[0].isspace()), raw.split()


@dataclass
class Cor:
    name: str
    directions: list
    row_names: list
    col_names: list
    entries: dict            # (row_idx, col_idx) -> value  (0-based)
    rhs: np.ndarray
    lower: np.ndarray
    upper: np.ndarray
    row_index: dict = field(default_factory=dict)
    col_index: dict = field(default_factory=dict)


def read_cor(path) -> Cor:
    section = ""
    name = ""
    rows, cols_tok, rhs_tok, bnd_tok = [], [], [], []
    for header, tok in _data_lines(path):
        if header:
            section = tok[0]
            if section not in ("NAME", "ROWS", "COLUMNS", "RHS", "BOUNDS", "ENDATA"):
                raise ValueError(f"unsupported cor section {section}")
            if section == "NAME":
                name = tok[1]
        elif section == "ROWS":
            rows.append(tok)
        elif section == "COLUMNS":
            cols_tok.append(tok)
        elif section == "RHS":
            rhs_tok.append(tok)
        elif section == "BOUNDS":
            bnd_tok.append(tok)
    directions = [t[0][0] for t in rows]
    row_names = [t[1] for t in rows]
    if directions[0] != "N":
        raise ValueError("first row of cor file is not the objective")
    col_names = list(dict.fromkeys(t[0] for t in cols_tok))
    row_index = {r: i for i, r in enumerate(row_names)}
    col_index = {c: i for i, c in enumerate(col_names)}
    entries = {}
    for tok in cols_tok:
        j = col_index[tok[0]]
        for rname, v in zip(tok[1::2], tok[2::2]):
            entries[(row_index[rname], j)] = float(v)
    rhs = np.zeros(len(row_names))
    for tok in rhs_tok:
        for rname, v in zip(tok[1::2], tok[2::2]):
            rhs[row_index[rname]] = float(v)
    lower = np.zeros(len(col_names))
    upper = np.full(len(col_names), np.inf)
    for tok in bnd_tok:
        bt, j = tok[0], col_index[tok[2]]
        if bt == "LO":
            lower[j] = float(tok[3])
        elif bt == "UP":
            upper[j] = float(tok[3])
        elif bt == "FX":
            lower[j] = upper[j] = float(tok[3])
        elif bt == "FR":
            lower[j], upper[j] = -np.inf, np.inf
        elif bt == "MI":
            lower[j] = -np.inf
        elif bt == "PL":
            upper[j] = np.inf
        else:
            raise ValueError(f"unsupported bound type {bt}")
    return Cor(name, directions, row_names, col_names, entries, rhs, lower, upper,
               row_index, col_index)


@dataclass
class Tim:
    name: str
    periods: list            # [(period_name, col_name, row_name)]


def read_tim(path) -> Tim:
    name, periods, section = "", [], ""
    for header, tok in _data_lines(path):
        if header:
            section = tok[0]
            if section == "TIME":
                name = tok[1]
        elif section == "PERIODS":
            periods.append((tok[2], tok[0], tok[1]))
    return Tim(name, periods)


@dataclass
class Sto:
    name: str
    positions: list          # [(col_name, row_name)] in order of first appearance
    kind: list               # "DISCRETE" | "NORMAL" | "UNIFORM" per position
    params: list             # DISCRETE: (values, probs); NORMAL: (mean, var); UNIFORM: (a, b)


def read_sto(path) -> Sto:
    name, section, keys = "", "", []
    index, positions, kind, params = {}, [], [], []
    for header, tok in _data_lines(path):
        if header:
            section, keys = tok[0], tok[1:]
            if section == "STOCH":
                name = keys[0]
            continue
        if section != "INDEP":
            continue
        pos = (tok[0], tok[1])
        k = keys[0]
        if k == "DISCRETE":
            if pos not in index:
                index[pos] = len(positions)
                positions.append(pos); kind.append(k); params.append(([], []))
            vals, probs = params[index[pos]]
            vals.append(float(tok[2])); probs.append(float(tok[3]))
        elif k in ("NORMAL", "UNIFORM"):
            if pos not in index:
                index[pos] = len(positions)
                positions.append(pos); kind.append(k); params.append(None)
            kind[index[pos]] = k                     # a later line replaces the element (indep[pos] = ...)
            params[index[pos]] = (float(tok[2]), float(tok[3]))
        else:
            raise ValueError(f"unsupported INDEP keyword {k}")
    return Sto(name, positions, kind, params)


@dataclass
class Stage2:
    """Flat second-stage tables (0-based)."""
    n1: int
    m2: int
    n2: int
    row_names: list
    x_names: list
    y_names: list
    directions: list          # 'G' | 'L' | 'E' per stage-2 row
    rbar: np.ndarray          # [m2]
    T_colptr: np.ndarray      # CSC of Tbar, rows ascending in each column
    T_rowval: np.ndarray
    T_nzval: np.ndarray
    W: np.ndarray             # dense [m2, n2] (host LP solves only)
    cost: np.ndarray          # [n2]
    y_lower: np.ndarray
    y_upper: np.ndarray
    pos_row: np.ndarray       # int32 [s]
    pos_col: np.ndarray       # int32 [s], -1 = RHS
    x_lower: np.ndarray
    x_upper: np.ndarray
    x_cost: np.ndarray

    def T_dense(self):
        T = np.zeros((self.m2, self.n1))
        for j in range(self.n1):
            for k in range(self.T_colptr[j], self.T_colptr[j + 1]):
                T[self.T_rowval[k], j] = self.T_nzval[k]
        return T


def stage2_tables(cor: Cor, tim: Tim, sto: Sto | None = None) -> Stage2:
    if len(tim.periods) != 2:
        raise ValueError("two-stage problems only")
    c2 = cor.col_index[tim.periods[1][1]]
    r2 = cor.row_index[tim.periods[1][2]]
    ncol, nrow = len(cor.col_names), len(cor.row_names)
    n1, n2, m2 = c2, ncol - c2, nrow - r2
    cols = [[] for _ in range(n1)]
    W = np.zeros((m2, n2))
    cost = np.zeros(n2)
    x_cost = np.zeros(n1)
    for (i, j), v in cor.entries.items():
        if i == 0:
            if j >= c2:
                cost[j - c2] = v
            else:
                x_cost[j] = v
        elif i >= r2:
            if v == 0.0:
                continue
            if j < c2:
                cols[j].append((i - r2, v))
            else:
                W[i - r2, j - c2] = v
    colptr = np.zeros(n1 + 1, dtype=np.int64)
    rowval, nzval = [], []
    for j in range(n1):
        cols[j].sort()
        for r, v in cols[j]:
            rowval.append(r); nzval.append(v)
        colptr[j + 1] = len(rowval)
    pos_row, pos_col = [], []
    if sto is not None:
        for cname, rname in sto.positions:
            pos_row.append(cor.row_index[rname] - r2)
            pos_col.append(-1 if cname in ("RHS", "rhs") else cor.col_index[cname])
            if pos_row[-1] < 0 or pos_col[-1] >= n1:
                raise ValueError(f"random element ({cname},{rname}) is not in stage 2 / Tbar")
    return Stage2(
        n1=n1, m2=m2, n2=n2,
        row_names=cor.row_names[r2:], x_names=cor.col_names[:c2], y_names=cor.col_names[c2:],
        directions=cor.directions[r2:], rbar=cor.rhs[r2:].copy(),
        T_colptr=colptr, T_rowval=np.asarray(rowval, dtype=np.int64),
        T_nzval=np.asarray(nzval, dtype=np.float64), W=W, cost=cost,
        y_lower=cor.lower[c2:].copy(), y_upper=cor.upper[c2:].copy(),
        pos_row=np.asarray(pos_row, dtype=np.int32), pos_col=np.asarray(pos_col, dtype=np.int32),
        x_lower=cor.lower[:c2].copy(), x_upper=cor.upper[:c2].copy(), x_cost=x_cost)


def discrete_tables(sto: Sto):
    """Rectangular outcome tables of the DISCRETE elements: (vals [s, mo], cdf [s, mo], cnt [s]); the cdf of a
    row is the running sum of its probabilities, padded with 1.0 (continuous elements: one dummy outcome)."""
    s = len(sto.positions)
    mo = max([len(p[0]) for k, p in zip(sto.kind, sto.params) if k == "DISCRETE"] + [1])
    vals, cdf, cnt = np.zeros((s, mo)), np.ones((s, mo)), np.ones(s, dtype=np.int32)
    for e, (k, p) in enumerate(zip(sto.kind, sto.params)):
        if k == "DISCRETE":
            n = len(p[0])
            cnt[e] = n
            vals[e, :n] = p[0]
            cdf[e, :n] = np.cumsum(p[1])
    return vals, cdf, cnt


def full_tables(cor: Cor, st: Stage2, sto: Sto) -> dict:
    """Everything a host SD loop needs besides the cut formation: the first-stage rows
    ``row_lower <= A1 x <= row_upper`` (rows 1 .. r2-1 of the cor file restricted to the first-stage columns),
    bounds and costs of both stages, W, and the outcome tables -- the layout of
    ``tests/golden/instances/<name>_full.npz`` and of ``tools/run_sd.py``."""
    r2 = len(cor.row_names) - st.m2
    A1 = np.zeros((r2 - 1, st.n1))
    for (i, j), v in cor.entries.items():
        if 1 <= i < r2:
            if j >= st.n1:
                raise ValueError("a first-stage row touches a second-stage column")
            A1[i - 1, j] = v
    dirs1, b1 = cor.directions[1:r2], cor.rhs[1:r2]
    lo = np.array([b if d in "GE" else -np.inf for d, b in zip(dirs1, b1)])
    up = np.array([b if d in "LE" else np.inf for d, b in zip(dirs1, b1)])
    vals, cdf, cnt = discrete_tables(sto)
    return dict(n1=st.n1, m2=st.m2, n2=st.n2, rbar=st.rbar, T_colptr=st.T_colptr, T_rowval=st.T_rowval,
                T_nzval=st.T_nzval, pos_row=st.pos_row, pos_col=st.pos_col, x_lower=st.x_lower, x_upper=st.x_upper,
                x_cost=st.x_cost, A1=A1, row_lower=lo, row_upper=up, W=st.W, cost=st.cost, y_lower=st.y_lower,
                y_upper=st.y_upper, directions=np.array(st.directions), out_vals=vals, out_cdf=cdf, out_cnt=cnt)


def sample_values(sto: Sto, u: np.ndarray) -> np.ndarray:
    """Realised values [N, s] from uniforms ``u`` [N, s] (inverse CDF for DISCRETE and
    UNIFORM; NORMAL by the inverse normal CDF).  Mirrors ``rand(sto)``
    (``smps_sto.jl:113-149``) with an explicit uniform stream instead of a global RNG."""
    u = np.asarray(u, dtype=np.float64)
    out = np.empty_like(u)
    for e, (k, p) in enumerate(zip(sto.kind, sto.params)):
        if k == "DISCRETE":
            vals, probs = np.asarray(p[0]), np.asarray(p[1])
            cdf = np.cumsum(probs)
            idx = np.minimum(np.searchsorted(cdf, u[:, e], side="right"), len(vals) - 1)
            out[:, e] = vals[idx]
        elif k == "UNIFORM":
            out[:, e] = p[0] + (p[1] - p[0]) * u[:, e]
        else:
            from scipy.special import ndtri
            out[:, e] = p[0] + np.sqrt(p[1]) * ndtri(np.clip(u[:, e], 1e-300, 1 - 1e-16))
    return out


class NativeSmps:
    """``sqlp_smps`` handle: cor + tim + sto parsed and split by the C++ reader of the library."""

    _DIMS = ("rows", "cols", "cor_nnz", "n1", "n2", "m2", "T_nnz", "W_nnz", "r_nnz", "s", "max_outcomes",
             "periods")
    # `what` of sqlp_smps_name (include/sqlp_b200.h)
    COR_NAME, TIM_NAME, STO_NAME, ROW_NAME, COL_NAME, PERIOD_NAME, PERIOD_COL, PERIOD_ROW, ELEM_COL, ELEM_ROW = range(10)

    def __init__(self, cor_path, tim_path, sto_path=None):
        import ctypes as C
        from . import _lib
        self._C, self._lib = C, _lib
        self._h = C.c_void_p()
        enc = lambda p: None if p is None else str(p).encode()
        _lib.check(_lib.lib().sqlp_smps_load(enc(cor_path), enc(tim_path), enc(sto_path), C.byref(self._h)))
        d = np.zeros(len(self._DIMS), dtype=np.int64)
        _lib.check(_lib.lib().sqlp_smps_dims(self._h, d.ctypes.data))
        self.dims = dict(zip(self._DIMS, (int(v) for v in d)))

    def close(self):
        if self._h:
            self._lib.lib().sqlp_smps_destroy(self._h)
            self._h = self._C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _name(self, what, index=0):
        buf = self._C.create_string_buffer(256)
        self._lib.check(self._lib.lib().sqlp_smps_name(self._h, what, index, buf, 256))
        return buf.value.decode()

    @staticmethod
    def _p(a):
        return a.ctypes.data

    def cor(self) -> Cor:
        d = self.dims
        dirs = np.zeros(d["rows"], dtype="S1")
        rhs, lo, up = np.zeros(d["rows"]), np.zeros(d["cols"]), np.zeros(d["cols"])
        colptr = np.zeros(d["cols"] + 1, dtype=np.int64)
        rowval, nzval = np.zeros(d["cor_nnz"], dtype=np.int64), np.zeros(d["cor_nnz"])
        self._lib.check(self._lib.lib().sqlp_smps_cor(self._h, self._p(dirs), self._p(rhs), self._p(lo), self._p(up),
                                                      self._p(colptr), self._p(rowval), self._p(nzval)))
        rows = [self._name(self.ROW_NAME, i) for i in range(d["rows"])]
        cols = [self._name(self.COL_NAME, j) for j in range(d["cols"])]
        entries = {(int(rowval[k]), j): float(nzval[k]) for j in range(d["cols"])
                   for k in range(colptr[j], colptr[j + 1])}
        return Cor(self._name(self.COR_NAME), [c.decode() for c in dirs], rows, cols, entries, rhs, lo, up,
                   {r: i for i, r in enumerate(rows)}, {c: j for j, c in enumerate(cols)})

    def tim(self) -> Tim:
        return Tim(self._name(self.TIM_NAME), [(self._name(self.PERIOD_NAME, i), self._name(self.PERIOD_COL, i), self._name(self.PERIOD_ROW, i))
                                   for i in range(self.dims["periods"])])

    def elements(self):
        """(pos_row, pos_col, kind, par_a, par_b, cnt, vals[s, mo], probs[s, mo])"""
        s, mo = self.dims["s"], self.dims["max_outcomes"]
        pr, pc, kind, cnt = (np.zeros(s, dtype=np.int32) for _ in range(4))
        a, b = np.zeros(s), np.zeros(s)
        vals, probs = np.zeros((s, mo)), np.zeros((s, mo))
        self._lib.check(self._lib.lib().sqlp_smps_elements(self._h, self._p(pr), self._p(pc), self._p(kind), self._p(a),
                                                           self._p(b), self._p(cnt), self._p(vals), self._p(probs)))
        return pr, pc, kind, a, b, cnt, vals, probs

    def sto(self) -> Sto:
        pr, pc, kind, a, b, cnt, vals, probs = self.elements()
        names = ("DISCRETE", "NORMAL", "UNIFORM")
        params = [(list(vals[e, :cnt[e]]), list(probs[e, :cnt[e]])) if kind[e] == 0 else (float(a[e]), float(b[e]))
                  for e in range(len(pr))]
        return Sto(self._name(self.STO_NAME), [(self._name(self.ELEM_COL, e), self._name(self.ELEM_ROW, e)) for e in range(len(pr))],
                   [names[k] for k in kind], params)

    def stage2(self) -> Stage2:
        d = self.dims
        n1, n2, m2 = d["n1"], d["n2"], d["m2"]
        rbar, cost, x_cost = np.zeros(m2), np.zeros(n2), np.zeros(n1)
        Tc, Tr, Tv = np.zeros(n1 + 1, dtype=np.int64), np.zeros(d["T_nnz"], dtype=np.int64), np.zeros(d["T_nnz"])
        Wc, Wr, Wv = np.zeros(n2 + 1, dtype=np.int64), np.zeros(d["W_nnz"], dtype=np.int64), np.zeros(d["W_nnz"])
        self._lib.check(self._lib.lib().sqlp_smps_stage2(self._h, self._p(rbar), self._p(Tc), self._p(Tr), self._p(Tv),
                                                         self._p(Wc), self._p(Wr), self._p(Wv), self._p(cost),
                                                         self._p(x_cost)))
        W = np.zeros((m2, n2))
        for j in range(n2):
            W[Wr[Wc[j]:Wc[j + 1]], j] = Wv[Wc[j]:Wc[j + 1]]
        cor = self.cor()
        r2, c2 = d["rows"] - m2, n1
        pr, pc = self.elements()[:2]
        return Stage2(n1=n1, m2=m2, n2=n2, row_names=cor.row_names[r2:], x_names=cor.col_names[:c2],
                      y_names=cor.col_names[c2:], directions=cor.directions[r2:], rbar=rbar, T_colptr=Tc,
                      T_rowval=Tr, T_nzval=Tv, W=W, cost=cost, y_lower=cor.lower[c2:].copy(),
                      y_upper=cor.upper[c2:].copy(), pos_row=pr, pos_col=pc, x_lower=cor.lower[:c2].copy(),
                      x_upper=cor.upper[:c2].copy(), x_cost=x_cost)
