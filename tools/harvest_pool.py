#!/usr/bin/env python
"""Harvest REAL dual-vertex pools at the sizes BASELINE.json names.

The reference fills ``cell.dual_vertices`` with the optimal duals of the stage-2 LPs it solves at
the candidate and the incumbent of every SD iteration (``algorithm.jl:49-54``), deduplicated by the
rule of ``dual_set.jl:24-53``.  This script does the same thing offline: stage-2 LPs of sampled
scenarios (scipy's HiGHS dual simplex, as ``tools/make_golden.py``) at a sequence of first-stage
points around the expected-value solution, pushed IN ORDER through the oracle's dedup rule until K
distinct vertices exist.  Such pools hold what synthetic ones do not: clusters of vertices that differ
only beyond 2^-15 relative, exact ties at the harvest points and near-ties everywhere else -- the
regime the north star's 1e-12 argmax exemption exists for.

Two steps, so that the expensive one can run wherever scipy is (the GPU boxes have the same image):

  python tools/harvest_pool.py --lp storm ssn baa99-20
      (BUILD container only: reads /root/reference/spInput, data files not source)
      -> tests/golden/instances/<name>_lp.npz   stage-2 LP data with W in COO form (small, committed)

  python tools/harvest_pool.py storm 16384 [--procs P]
      -> tests/golden/pools/<name>_K<K>.npz     pool [K x m2] + the recipe (seed, points, solves);
         cached: an existing file with the same recipe is kept

Recipe (all counters are the splitmix64 generator of SURVEY.md 8(d)): point p is
x_p = clip(x_ev + 0.02 * (1 + p % 5) * span * (u(8 + p, j) - 1/2)), p = 0 is x_ev itself; LP number q
uses point q % n_points and the scenario drawn by inverse CDF from u(7, q * s + e); LPs without
complete recourse at a perturbed point are skipped; vertices are pushed in LP order.
"""
from __future__ import annotations

import argparse
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

INST = os.path.join(ROOT, "tests", "golden", "instances")
POOLS = os.path.join(ROOT, "tests", "golden", "pools")
RECIPE = 3          # bump when the recipe below changes (cached pools with another number are rebuilt)


def write_lp_fixture(name):
    """Stage-2 LP data of a shipped instance -> <name>_lp.npz (W as COO triplets)."""
    from tools.make_golden import load
    cor, tim, sto, st = load(name)
    ii, jj = np.nonzero(st.W)
    np.savez_compressed(os.path.join(INST, f"{name}_lp.npz"), n2=st.n2, m2=st.m2, W_i=ii.astype(np.int32),
                        W_j=jj.astype(np.int32), W_v=st.W[ii, jj], cost=st.cost, y_lower=st.y_lower,
                        y_upper=st.y_upper, directions=np.array(st.directions))
    print(f"{name}_lp.npz: W {st.W.shape} nnz {len(ii)}")


def load_lp(name):
    z = dict(np.load(os.path.join(INST, f"{name}.npz")))
    lp = dict(np.load(os.path.join(INST, f"{name}_lp.npz")))
    W = np.zeros((int(lp["m2"]), int(lp["n2"])))
    W[lp["W_i"], lp["W_j"]] = lp["W_v"]
    lp["W"] = W
    return z, lp


_G = {}


def _init(name):
    from scipy.optimize import linprog  # noqa: F401  (import cost paid once per worker)
    z, lp = load_lp(name)
    m2, n1 = int(z["m2"]), int(z["n1"])
    T = np.zeros((m2, n1))
    for j in range(n1):
        for q in range(z["T_colptr"][j], z["T_colptr"][j + 1]):
            T[z["T_rowval"][q], j] = z["T_nzval"][q]
    dirs = np.asarray(lp["directions"])
    L, G, E = (dirs == "L"), (dirs == "G"), (dirs == "E")
    _G.update(z=z, lp=lp, T=T, L=L, G=G, E=E,
              A_ub=np.vstack([lp["W"][L], -lp["W"][G]]), A_eq=lp["W"][E] if E.any() else None,
              bounds=[(None if np.isinf(l) else l, None if np.isinf(u) else u)
                      for l, u in zip(lp["y_lower"], lp["y_upper"])])


def _solve(args):
    """Dual vertex of one stage-2 LP (JuMP's sign convention for a MIN problem), or None."""
    from scipy.optimize import linprog
    x, values = args
    g = _G
    z = g["z"]
    r = z["rbar"].copy()
    T = g["T"]
    Tx = None
    for e, v in enumerate(values):
        if z["pos_col"][e] < 0:
            r[z["pos_row"][e]] = v
        else:                                   # no shipped instance has random Tbar entries
            if Tx is None:
                T = T.copy()
            T[z["pos_row"][e], z["pos_col"][e]] = v
    b = r - T @ x
    L, G, E = g["L"], g["G"], g["E"]
    b_ub = np.concatenate([b[L], -b[G]])
    kw = {}
    if g["A_eq"] is not None:
        kw.update(A_eq=g["A_eq"], b_eq=b[E])
    res = linprog(g["lp"]["cost"], A_ub=g["A_ub"] if len(b_ub) else None, b_ub=b_ub if len(b_ub) else None,
                  bounds=g["bounds"], method="highs-ds", **kw)
    if res.status != 0:
        return None
    dual = np.zeros(len(r))
    nL = int(L.sum())
    if len(b_ub):
        m = res.ineqlin.marginals
        dual[L] = m[:nL]
        dual[G] = -m[nL:]
    if E.any():
        dual[E] = res.eqlin.marginals
    return dual


def harvest_points(z, n_points):
    from oracle import oracle as O
    x0 = z["x_ev"]
    n1 = len(x0)
    up = np.where(np.isfinite(z["x_upper"]), z["x_upper"], np.inf)
    span = np.where(np.isfinite(z["x_upper"]), z["x_upper"], np.maximum(1.0, np.abs(x0))) - z["x_lower"]
    pts = [x0]
    for p in range(1, n_points):
        pts.append(np.clip(x0 + 0.02 * (1 + p % 5) * span * (O.u01(8 + p, np.arange(n1)) - 0.5), z["x_lower"], up))
    return pts


def sample_values(z, q0, n):
    from oracle import oracle as O
    s = len(z["pos_row"])
    u = O.u01(7, (np.arange(q0, q0 + n, dtype=np.uint64)[:, None] * np.uint64(s)
                  + np.arange(s, dtype=np.uint64)[None, :]))
    idx = (u[:, :, None] >= z["out_cdf"][None, :, :]).sum(axis=2)
    idx = np.minimum(idx, np.maximum(z["out_cnt"][None, :] - 1, 0))
    return np.take_along_axis(np.broadcast_to(z["out_vals"], (n,) + z["out_vals"].shape), idx[:, :, None], 2)[:, :, 0]


def pool_path(name, K):
    return os.path.join(POOLS, f"{name}_K{K}.npz")


def harvest(name, K, procs=0, n_points=64, verbose=True, max_solves=None):
    """Returns the [K x m2] pool (cached in tests/golden/pools/)."""
    from oracle import oracle as O
    path = pool_path(name, K)
    if os.path.exists(path):
        zc = np.load(path)
        if int(zc["recipe"]) == RECIPE and len(zc["pool"]) == K:
            return zc["pool"]
    z, _ = load_lp(name)
    m2 = int(z["m2"])
    pts = harvest_points(z, n_points)
    procs = procs or len(os.sched_getaffinity(0))
    t0 = time.time()
    # dedup in LP order through the oracle's push (dual_set.jl:84-93): a growing fixed-capacity pool
    import ctypes as C
    cap = K
    pool = np.zeros((cap, m2))
    hashes = np.zeros(cap, dtype=np.uint64)
    Kc, one = C.c_int64(0), C.c_int32(0)
    solves = fails = 0
    batch = max(256, 32 * procs)
    max_solves = max_solves or 40 * K
    with mp.Pool(procs, initializer=_init, initargs=(name,)) as workers:
        q0 = 0
        while Kc.value < K and q0 < max_solves:
            vals = sample_values(z, q0, batch)
            jobs = [(pts[(q0 + i) % n_points], vals[i]) for i in range(batch)]
            for dual in workers.imap(_solve, jobs, chunksize=8):     # imap keeps LP order
                solves += 1
                if dual is None:
                    fails += 1
                    continue
                if Kc.value < K:
                    O.lib().orc_pool_push(O._p(pool), O._p(hashes), C.byref(Kc), cap, m2, O._p(dual), C.byref(one))
            q0 += batch
            if verbose:
                print(f"  {name}: {Kc.value}/{K} distinct after {solves} LPs ({fails} infeasible), "
                      f"{time.time() - t0:.0f} s", flush=True)
    if Kc.value < K:
        raise RuntimeError(f"{name}: only {Kc.value} distinct vertices after {solves} LPs")
    os.makedirs(POOLS, exist_ok=True)
    np.savez_compressed(path, pool=pool, recipe=RECIPE, n_points=n_points, solves=solves, infeasible=fails,
                        seconds=time.time() - t0)
    return pool


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--lp", nargs="*", default=None, help="write <name>_lp.npz fixtures (build container only)")
    ap.add_argument("name", nargs="?")
    ap.add_argument("K", nargs="?", type=int)
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--points", type=int, default=64)
    a = ap.parse_args()
    if a.lp is not None:
        for nm in a.lp:
            write_lp_fixture(nm)
        sys.exit(0)
    t0 = time.time()
    P = harvest(a.name, a.K, a.procs, a.points)
    print(f"{a.name}: pool {P.shape} in {time.time() - t0:.0f} s -> {pool_path(a.name, a.K)} "
          f"({os.path.getsize(pool_path(a.name, a.K)) / 1e6:.1f} MB)")
