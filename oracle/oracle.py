"""ctypes front-end of ``oracle/sqlp_oracle.c`` (TEST INFRASTRUCTURE ONLY).

The C file restates the reference's Julia hot path line by line (citations in its
header).  This module marshals numpy arrays into it and adds three plain-Python
restatements that need no C: the cut bookkeeping next to the path (``cut_evaluate``,
``cut_master_rows``, ``cut_check_improvement``: epigraph.jl:101-117,177-220, cell.jl:163-202,
improvement.jl:19-49) and the host twin of the device sampler (``sample_twin``,
smps_sto.jl:113-149).  Parity status: pinned on the reference's own lands-sized test
vectors (tests/test_oracle_golden.py), unpinned beyond (no Julia in the image).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsqlp_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (make -C oracle)."""
    src = os.path.join(_HERE, "sqlp_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "--no-print-directory"], check=True,
                       capture_output=True)
    return _SO


class _Problem(C.Structure):
    _fields_ = [
        ("m2", C.c_int64), ("n1", C.c_int64),
        ("rbar", C.c_void_p),
        ("T_colptr", C.c_void_p), ("T_rowval", C.c_void_p), ("T_nzval", C.c_void_p),
        ("s", C.c_int64),
        ("pos_row", C.c_void_p), ("pos_col", C.c_void_p),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_round_sig.restype = C.c_double
        L.orc_round_sig.argtypes = [C.c_double]
        L.orc_hash_dual_vector.restype = C.c_uint64
        L.orc_hash_dual_vector.argtypes = [C.c_void_p, C.c_int64]
        L.orc_isequal.restype = C.c_int
        L.orc_isequal.argtypes = [C.c_void_p, C.c_int64, C.c_uint64,
                                  C.c_void_p, C.c_int64, C.c_uint64]
        L.orc_pool_push.restype = C.c_int64
        L.orc_pool_push.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                    C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_delta_coefficients.restype = None
        L.orc_delta_coefficients.argtypes = [C.c_void_p] * 4
        L.orc_eval_dual.restype = C.c_double
        L.orc_eval_dual.argtypes = [C.c_void_p] * 4
        L.orc_argmax_procedure.restype = None
        L.orc_argmax_procedure.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                           C.c_void_p]
        L.orc_score_pair.restype = None
        L.orc_score_pair.argtypes = [C.c_void_p] * 6
        L.orc_build_sasa_cut.restype = C.c_int32
        L.orc_build_sasa_cut.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                         C.c_double, C.c_void_p, C.c_void_p, C.c_int64,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_bench_argmax.restype = C.c_int32
        L.orc_bench_argmax.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                       C.c_int32, C.c_int32]
        L.orc_max_threads.restype = C.c_int32
        L.orc_u01.restype = C.c_double
        L.orc_u01.argtypes = [C.c_uint64, C.c_uint64]
        _lib = L
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@dataclass
class Problem:
    """Template coefficients + stochastic-position table of one epigraph
    (``sdSubprobCoefficients``, subprob.jl:4-12, flattened)."""
    m2: int
    n1: int
    rbar: np.ndarray            # dense [m2]
    T_colptr: np.ndarray        # int64 [n1+1], 0-based CSC
    T_rowval: np.ndarray        # int64 [nnz]
    T_nzval: np.ndarray         # f64 [nnz]
    pos_row: np.ndarray         # int32 [s]
    pos_col: np.ndarray         # int32 [s]; -1 = RHS

    def __post_init__(self):
        self.rbar = _f64(self.rbar)
        self.T_colptr = np.ascontiguousarray(self.T_colptr, dtype=np.int64)
        self.T_rowval = np.ascontiguousarray(self.T_rowval, dtype=np.int64)
        self.T_nzval = _f64(self.T_nzval)
        self.pos_row = np.ascontiguousarray(self.pos_row, dtype=np.int32)
        self.pos_col = np.ascontiguousarray(self.pos_col, dtype=np.int32)
        assert self.rbar.shape == (self.m2,)
        assert self.T_colptr.shape == (self.n1 + 1,)
        self._c = _Problem(self.m2, self.n1, self.rbar.ctypes.data,
                           self.T_colptr.ctypes.data, self.T_rowval.ctypes.data,
                           self.T_nzval.ctypes.data, len(self.pos_row),
                           self.pos_row.ctypes.data, self.pos_col.ctypes.data)

    @property
    def s(self) -> int:
        return len(self.pos_row)

    @property
    def ref(self):
        return C.byref(self._c)

    def T_dense(self) -> np.ndarray:
        T = np.zeros((self.m2, self.n1))
        for j in range(self.n1):
            for k in range(self.T_colptr[j], self.T_colptr[j + 1]):
                T[self.T_rowval[k], j] = self.T_nzval[k]
        return T


# ---- A5 ---------------------------------------------------------------------------

def round_sig(x: float) -> float:
    return lib().orc_round_sig(float(x))


def hash_dual_vector(v) -> int:
    v = _f64(v)
    return int(lib().orc_hash_dual_vector(_p(v), len(v)))


def isequal(a, b) -> bool:
    a, b = _f64(a), _f64(b)
    return bool(lib().orc_isequal(_p(a), len(a), hash_dual_vector(a),
                                  _p(b), len(b), hash_dual_vector(b)))


class DualVertexSet:
    """``sdDualVertexSet`` (dual_set.jl:69-127) over vectors of arbitrary length."""

    def __init__(self, data=()):
        self.data: list[np.ndarray] = []
        for d in data:
            self.push(d)

    def push(self, v):
        """Returns (inserted, 0-based index of v or of its first duplicate)."""
        v = _f64(v).copy()
        for k, w in enumerate(self.data):
            if isequal(v, w):
                return False, k
        self.data.append(v)
        return True, len(self.data) - 1

    def __len__(self):
        return len(self.data)

    def __iter__(self):
        return iter(self.data)

    def matrix(self) -> np.ndarray:
        return np.stack(self.data) if self.data else np.zeros((0, 0))


def pool_push_many(m2: int, vectors: np.ndarray):
    """Fixed-length pool: push every row of ``vectors``; returns (pool, inserted, index)."""
    vectors = _f64(vectors).reshape(-1, m2)
    n = len(vectors)
    pool = np.zeros((max(n, 1), m2))
    hashes = np.zeros(max(n, 1), dtype=np.uint64)
    K = C.c_int64(0)
    ins = np.zeros(n, dtype=np.int32)
    idx = np.zeros(n, dtype=np.int64)
    one = C.c_int32(0)
    for i in range(n):
        idx[i] = lib().orc_pool_push(_p(pool), _p(hashes), C.byref(K), len(pool), m2,
                                     _p(vectors[i]), C.byref(one))
        ins[i] = one.value
    return pool[:K.value].copy(), ins, idx


# ---- A1 / A7 / A2 / A3 --------------------------------------------------------------

def delta_coefficients(P: Problem, values):
    values = _f64(values)
    drhs = np.zeros(P.m2)
    dT = np.zeros(max(P.s, 1))
    lib().orc_delta_coefficients(P.ref, _p(values), _p(drhs), _p(dT))
    return drhs, dT[:P.s]


def eval_dual(P: Problem, values, x, dual) -> float:
    values, x, dual = _f64(values), _f64(x), _f64(dual)
    return lib().orc_eval_dual(P.ref, _p(values), _p(x), _p(dual))


def argmax_procedure(P: Problem, values, x, pool, want_second=False):
    values = _f64(values).reshape(-1, max(P.s, 1)) if P.s else _f64(values)
    N = len(values) if P.s else int(np.asarray(values).shape[0])
    x, pool = _f64(x), _f64(pool).reshape(-1, P.m2)
    mv = np.zeros(N)
    mi = np.zeros(N, dtype=np.int64)
    sec = np.zeros(N) if want_second else None
    lib().orc_argmax_procedure(P.ref, N, _p(values), _p(x), _p(pool), len(pool),
                               _p(mv), _p(mi), _p(sec) if want_second else None)
    return (mv, mi, sec) if want_second else (mv, mi)


def score_pair(P: Problem, values, x, vertex):
    values, x, vertex = _f64(values), _f64(x), _f64(vertex)
    s = C.c_double(0)
    sl = C.c_longdouble(0)
    lib().orc_score_pair(P.ref, _p(values), _p(x), _p(vertex), C.byref(s), C.byref(sl))
    return s.value, sl.value


def build_sasa_cut(P: Problem, values, weights, x, pool, total_weight=None, forced_idx=None):
    """Returns dict(alpha, beta, weight_mark, val, max_val, max_idx, status)."""
    weights = _f64(weights)
    if P.s:
        values = _f64(values).reshape(-1, P.s)
        N = len(values)
    else:                      # no random element: the scenarios differ by their weights only
        values, N = np.zeros(1), len(weights)
    x, pool = _f64(x), _f64(pool).reshape(-1, P.m2)
    if total_weight is None:
        total_weight = 0.0
        for w in weights:          # epigraph.jl:89 sequential accumulation
            total_weight += float(w)
    alpha = C.c_double(0)
    wm = C.c_double(0)
    val = C.c_double(0)
    beta = np.zeros(P.n1)
    mv = np.zeros(N)
    mi = np.zeros(N, dtype=np.int64)
    fi = None if forced_idx is None else np.ascontiguousarray(forced_idx, dtype=np.int64)
    st = lib().orc_build_sasa_cut(P.ref, N, _p(values), _p(weights), float(total_weight),
                                  _p(x), _p(pool), len(pool), C.byref(alpha), _p(beta),
                                  C.byref(wm), C.byref(val), _p(mv), _p(mi),
                                  None if fi is None else _p(fi))
    return dict(alpha=alpha.value, beta=beta, weight_mark=wm.value, val=val.value,
                max_val=mv, max_idx=mi, status=st)


def bench_argmax(P: Problem, values, x, pool, threads=0, dot_kind=1):
    values = _f64(values).reshape(-1, max(P.s, 1))
    N = len(values)
    x, pool = _f64(x), _f64(pool).reshape(-1, P.m2)
    mv = np.zeros(N)
    mi = np.zeros(N, dtype=np.int64)
    used = lib().orc_bench_argmax(P.ref, N, _p(values), _p(x), _p(pool), len(pool),
                                  _p(mv), _p(mi), threads, dot_kind)
    return mv, mi, used


def max_threads() -> int:
    return lib().orc_max_threads()


def u01(seed: int, idx) -> np.ndarray:
    """Vectorised numpy twin of orc_u01 (splitmix64 counter generator)."""
    idx = np.asarray(idx, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) ^ (idx * np.uint64(0x9E3779B97F4A7C15))
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def u01_open(seed: int, idx) -> np.ndarray:
    """Open-interval twin of the device's u01_open: (top 53 bits + 1/2) / 2^53."""
    idx = np.asarray(idx, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) ^ (idx * np.uint64(0x9E3779B97F4A7C15))
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def sample_twin(seed: int, g0: int, n: int, kind, par_a, par_b, out_vals=None, out_cdf=None, out_cnt=None):
    """Host twin of sqlp_epi_sample_scenarios for scenarios g0 .. g0+n-1 (values [n, s]): the
    reference's rand(sto) (smps_sto.jl:113-149) with the counter generator in place of the global
    RNG.  DISCRETE: inverse CDF on u01; NORMAL(mean, variance): mean + sqrt(variance) * Phi^-1(u);
    UNIFORM(left, right): left + (right - left) * u, u from u01_open.  Phi^-1 is scipy's ndtri, so
    NORMAL values agree with the device's normcdfinv to rounding, not bit for bit."""
    from scipy.special import ndtri
    kind = np.asarray(kind)
    s = len(kind)
    g = np.arange(g0, g0 + n, dtype=np.uint64)
    ctr = g[:, None] * np.uint64(s) + np.arange(s, dtype=np.uint64)[None, :]
    u, uo = u01(seed, ctr), u01_open(seed, ctr)
    out = np.empty((n, s))
    for e in range(s):
        if kind[e] == 0:
            idx = (u[:, e, None] >= np.asarray(out_cdf)[e][None, :]).sum(axis=1)
            idx = np.minimum(idx, max(int(out_cnt[e]) - 1, 0))
            out[:, e] = np.asarray(out_vals)[e][idx]
        elif kind[e] == 1:
            out[:, e] = par_a[e] + np.sqrt(par_b[e]) * ndtri(uo[:, e])
        else:
            out[:, e] = par_a[e] + (par_b[e] - par_a[e]) * uo[:, e]
    return out


# ---------------------------------------------------------------- cut list (N1 / N3) -----
# Plain-Python restatement of the reference's cut bookkeeping; a cut is (alpha, beta, weight_mark).

def cut_evaluate(cuts, incumbent, x, total_weight, lower_bound, objective_weight=1.0):
    """evaluate_epigraph, src/sd_algorithm/epigraph.jl:177-220 (MIN sense): pointwise max of the
    lower bound, the discounted cuts and the undiscounted incumbent cut, times the epigraph's weight."""
    x = np.asarray(x, dtype=np.float64)
    best = lower_bound
    for alpha, beta, wm in cuts:
        discount = wm / total_weight
        val = discount * (alpha + float(np.dot(np.asarray(beta, dtype=np.float64), x))) + (1 - discount) * lower_bound
        if val > best:
            best = val
    if incumbent is not None:
        alpha, beta, _ = incumbent
        val = alpha + float(np.dot(np.asarray(beta, dtype=np.float64), x))
        if val > best:
            best = val
    return objective_weight * best


def cut_master_rows(cuts, incumbent, total_weight, lower_bound):
    """sync_cuts!, src/sd_algorithm/cell.jl:163-202 with add_cut_to_master!, epigraph.jl:101-117:
    rows (discount alpha + (1 - discount) lb, discount beta), the incumbent cut last with discount 1."""
    rows = []
    for alpha, beta, wm in cuts:
        discount = wm / total_weight
        rows.append(np.concatenate([[discount * alpha + (1 - discount) * lower_bound],
                                    discount * np.asarray(beta, dtype=np.float64)]))
    if incumbent is not None:
        alpha, beta, _ = incumbent
        rows.append(np.concatenate([[1.0 * alpha + (1 - 1.0) * lower_bound], 1.0 * np.asarray(beta, dtype=np.float64)]))
    n1 = len(cuts[0][1]) if cuts else (len(incumbent[1]) if incumbent is not None else 0)
    return np.asarray(rows).reshape(len(rows), 1 + n1)


def cut_check_improvement(last, current, x_cand, x_inc, cost, q=0.2):
    """check_improvement, src/sd_algorithm/improvement.jl:19-49.  ``last`` / ``current`` are lists of
    (cuts, incumbent, total_weight, lower_bound, objective_weight), one per epigraph."""
    f = lambda x: float(np.dot(np.asarray(cost, dtype=np.float64), np.asarray(x, dtype=np.float64)))
    multi = lambda es, x: sum(cut_evaluate(c, i, x, tw, lb, w) for c, i, tw, lb, w in es)
    f_cand, f_inc = f(x_cand), f(x_inc)
    cand = multi(current, x_cand) + f_cand
    inc = multi(current, x_inc) + f_inc
    required = q * ((multi(last, x_cand) + f_cand) - (multi(last, x_inc) + f_inc))
    return cand, inc, required, cand < inc + required
