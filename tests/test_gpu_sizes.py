"""Parity AT THE SIZES BASELINE.json names, on REAL dual-vertex pools.

The pools are harvested by ``tools/harvest_pool.py`` -- stage-2 LP duals through the reference's dedup rule
(``dual_set.jl:24-53,84-93``) -- and committed under ``tests/golden/pools``.  Real pools hold what synthetic ones
do not: exact ties (storm's LPs are dual degenerate: about 28 of its 16 384 vertices tie for the best score of a
typical scenario) and near-ties at the 2^-15 resolution of the dedup rule, i.e. the regime the north star's
1e-12 argmax exemption exists for.  Checked through the C ABI against the oracle (reference
``subprob.jl:148-166``, ``epigraph.jl:134-143``):

* C2 baa99-20, K = 1 024 x N = 10 000: the oracle sweeps EVERYTHING, both points;
* C3 ssn, K = 5 000 x N = 100 000: 2 000 random scenarios under the argmax rule, both cuts against
  ``orc_build_sasa_cut(forced_idx)`` over ALL scenarios;
* C4 storm, K = 16 384 x N = 1 000 000 in E = 4 weighted epigraphs: 2 000 random scenarios per point, both cuts of
  an epigraph over all its 250 000 scenarios.

The number of exemptions used is reported (``-s`` shows it) and bounded.
"""
import numpy as np
import pytest

from tests.helpers import load_instance, load_pool, sampled_values_at

pytestmark = pytest.mark.gpu

CUT_RTOL = 1e-10


@pytest.fixture(scope="module")
def T():
    from sqlp_b200 import twosd
    return twosd


def coef_of(T, P):
    return T.sdSubprobCoefficients.from_tables(P.rbar, P.T_colptr, P.T_rowval, P.T_nzval, P.pos_row, P.pos_col)


def argmax_rule(oracle, P, vals, x, pool, got_val, got_idx, rel_gap=1e-12):
    """North-star rule on a set of scenarios, oracle on all host cores (index-order dots, the oracle's own
    arithmetic).  Returns the number of exemptions used."""
    ov, oi, _ = oracle.bench_argmax(P, vals, x, pool, threads=0, dot_kind=0)
    got_idx = np.asarray(got_idx)
    bad = np.nonzero(oi != got_idx)[0]
    for i in bad:
        assert 0 <= got_idx[i] < len(pool)
        sc, _ = oracle.score_pair(P, vals[i], x, pool[got_idx[i]])
        tol = rel_gap * max(abs(ov[i]), 1.0)
        assert ov[i] - sc <= tol, (f"scenario {i}: device picked {got_idx[i]} (oracle score {sc!r}), oracle picked "
                                   f"{oi[i]} ({ov[i]!r}): gap beyond {tol:g}")
    err = np.max(np.abs(np.asarray(got_val) - ov) / np.maximum(np.abs(ov), 1.0))
    assert err <= 1e-10, err
    return len(bad)


def cut_vs_forced(oracle, P, vals, w, x, pool, idx, cut, val):
    """alpha, beta, val against the oracle's accumulation on the device's (validated) selection, relative to the
    absolute-sum scale of every coefficient."""
    ref = oracle.build_sasa_cut(P, vals, w, x, pool, forced_idx=idx)
    assert ref["status"] == 0
    p = np.asarray(w) / ref["weight_mark"]
    sel = pool[idx]
    r_i = np.tile(P.rbar, (len(vals), 1))
    for e in range(P.s):
        r_i[:, P.pos_row[e]] = vals[:, e]
    sa = np.sum(p * np.abs(np.einsum("ij,ij->i", sel, r_i))) + 1e-300
    sb = (p[:, None] * np.abs(sel @ P.T_dense())).sum(axis=0) + 1e-300
    ea = abs(cut.alpha - ref["alpha"]) / max(sa, abs(ref["alpha"]))
    eb = np.max(np.abs(cut.beta - ref["beta"]) / np.maximum(sb, np.abs(ref["beta"])))
    assert ea <= CUT_RTOL and eb <= CUT_RTOL, (ea, eb)
    assert cut.weight_mark == ref["weight_mark"]
    assert abs(val - ref["val"]) <= CUT_RTOL * (np.sum(p * np.abs(ref["max_val"])) + 1e-300)
    return max(ea, eb)


def test_c2_baa99_full_oracle_sweep(T, oracle):
    """BASELINE.json configs[1]: every one of the 10 000 x 1 024 scores, both points."""
    P, z = load_instance("baa99-20")
    pool = load_pool("baa99-20", 1024)
    N = 10_000
    g = np.arange(N)
    vals = sampled_values_at(z, 1, g)
    dvs = T.sdDualVertexSet(m2=P.m2)
    ins, _ = dvs.push_many(pool)
    assert ins.all()
    epi = T.sdEpigraph(coef_of(T, P), 1.0, 0.0, dvs)
    epi.set_outcomes(z["out_vals"], z["out_cdf"], z["out_cnt"])
    epi.sample_scenarios(N, seed=1)
    w = np.ones(N)
    xs = (z["x_ev"], z["x_alt"])
    (cand, inc), val = epi.build_cuts2(*xs, with_val=True)
    exempt = 0
    for x, cut, v in zip(xs, (cand, inc), val):
        mv, mi = epi.argmax(x)
        exempt += argmax_rule(oracle, P, vals, x, pool, mv, mi)
        cut_vs_forced(oracle, P, vals, w, x, pool, mi, cut, v)
    print(f"C2 baa99-20 K=1024 x N=10000, full sweep: {exempt} exemptions of {2 * N}")
    assert exempt <= 2 * N // 100


def test_c3_ssn_full_size(T, oracle):
    """BASELINE.json configs[2]: K = 5 000 real vertices x N = 100 000 sampled scenarios."""
    P, z = load_instance("ssn")
    pool = load_pool("ssn", 5000)
    N = 100_000
    dvs = T.sdDualVertexSet(m2=P.m2)
    ins, _ = dvs.push_many(pool)
    assert ins.all()
    epi = T.sdEpigraph(coef_of(T, P), 1.0, 0.0, dvs)
    epi.set_outcomes(z["out_vals"], z["out_cdf"], z["out_cnt"])
    epi.sample_scenarios(N, seed=1)
    vals = sampled_values_at(z, 1, np.arange(N))
    w = np.ones(N)
    xs = (z["x_ev"], z["x_alt"])
    (cand, inc), val = epi.build_cuts2(*xs, with_val=True)
    pick = np.sort(np.random.default_rng(3).choice(N, size=2000, replace=False))
    exempt = 0
    for x, cut, v in zip(xs, (cand, inc), val):
        mv, mi = epi.argmax(x)
        exempt += argmax_rule(oracle, P, vals[pick], x, pool, mv[pick], mi[pick])
        cut_vs_forced(oracle, P, vals, w, x, pool, mi, cut, v)          # all 100 000 scenarios
        assert abs(cut.alpha + cut.beta @ x - v) <= 1e-10 * (abs(cut.alpha) + np.abs(cut.beta * x).sum())   # G5
    st = epi.screen_stats()
    print(f"C3 ssn K=5000 x N=100000: {exempt} exemptions of {2 * len(pick)} sampled; screening {st}")
    assert exempt <= 40


def test_c4_storm_full_size(T, oracle):
    """BASELINE.json configs[3] on one GPU: K = 16 384 real vertices x N = 1 000 000 scenarios in E = 4 weighted
    epigraphs (the sharded form of the same job is tests/test_gpu_dist.py and bench.py --gpus N)."""
    P, z = load_instance("storm")
    pool = load_pool("storm", 16384)
    E, Ne = 4, 250_000
    dvs = T.sdDualVertexSet(m2=P.m2)
    ins, _ = dvs.push_many(pool)
    assert ins.all() and len(dvs) == 16384
    coef = coef_of(T, P)
    epis = []
    for e in range(E):
        epi = T.sdEpigraph(coef, 1.0 / E, 0.0, dvs)
        epi.set_outcomes(z["out_vals"], z["out_cdf"], z["out_cnt"])
        epi.sample_scenarios(Ne, seed=101 + e, weight_seed=201 + e)
        epis.append(epi)
    xs = (z["x_ev"], z["x_alt"])
    out = T.build_cuts_at_candidate_and_incumbent(epis, *xs)
    rng = np.random.default_rng(4)
    exempt = checked = 0
    for e in (0, 3):                                   # two of the four epigraphs, 1 000 scenarios per point each
        g = np.arange(Ne)
        pick = np.sort(rng.choice(Ne, size=1000, replace=False))
        vals_pick = sampled_values_at(z, 101 + e, pick)
        for xi, x in enumerate(xs):
            mv, mi = epis[e].argmax(x)
            assert (mi >= 0).all() and (mi < 16384).all()
            exempt += argmax_rule(oracle, P, vals_pick, x, pool, mv[pick], mi[pick])
            checked += len(pick)
            if e == 0:                                 # both cuts of one epigraph over all its 250 000 scenarios
                vals = sampled_values_at(z, 101 + e, g)
                w = 0.5 + oracle.u01(201 + e, g)
                # G5 on the device's numbers, then the oracle's accumulation on the device's selection
                cut = out[e][xi]
                cut_vs_forced(oracle, P, vals, w, x, pool, mi, cut, float(np.dot(w / cut.weight_mark, mv)))
    print(f"C4 storm K=16384 x N=1e6 (E=4): {exempt} exemptions of {checked} sampled (exact ties of a dual "
          f"degenerate LP); screening {epis[0].screen_stats()}")
    assert exempt <= checked // 2


def test_score_minus_bias_shortcut_is_guarded(T, oracle):
    """with delta_T == 0 the winning dot is recovered as score - bias.  When |tau_k . x| is
    1e8 x |pi . (rbar + delta_r)| half an ulp of the score would show in alpha at the 1e-10 level: the
    reduction then recomputes the dot from D."""
    P, z = load_instance("storm")
    pool = z["pool"]
    N = 3000
    vals = sampled_values_at(z, 1, np.arange(N))
    w = 0.5 + oracle.u01(4, np.arange(N))
    dvs = T.sdDualVertexSet(m2=P.m2)
    dvs.push_many(pool)
    epi = T.sdEpigraph(coef_of(T, P), 1.0, 0.0, dvs)
    epi.add_scenarios(vals, w)
    for scale in (1e4, 1e8, 1e11):
        x = z["x_ev"] * scale + scale
        cut, v = epi.build_cut(x, with_val=True)
        mv, mi = epi.argmax(x)
        cut_vs_forced(oracle, P, vals, w, x, pool, mi, cut, v)
    # a pool whose rho and dot cancel: pi . rbar == -(pi . delta) for the mean scenario
    pool2 = pool.copy()
    mean_r = P.rbar.copy()
    rows = P.pos_row
    pool2[:, rows[0]] -= (pool2 @ mean_r) / mean_r[rows[0]]
    dvs2 = T.sdDualVertexSet(m2=P.m2)
    dvs2.push_many(pool2)
    epi2 = T.sdEpigraph(coef_of(T, P), 1.0, 0.0, dvs2)
    epi2.add_scenarios(vals, w)
    pool2 = np.stack(list(dvs2))
    for scale in (1.0, 1e8):
        x = z["x_alt"] * scale
        cut, v = epi2.build_cut(x, with_val=True)
        mv, mi = epi2.argmax(x)
        cut_vs_forced(oracle, P, vals, w, x, pool2, mi, cut, v)


@pytest.mark.parametrize("name,K,N", [("storm", 3000, 5000), ("ssn", 1200, 700), ("baa99-20", 1024, 130)])
def test_cut_reduction_by_vertex_weights(T, oracle, monkeypatch, name, K, N):
    """The two cut reductions -- one (rho, tau) row gathered per scenario (k_cut_partial) and per-vertex weight sums
    (k_cut_hist / k_cut_fold: the table is read once) -- regroup the same sum (epigraph.jl:134-143): both within
    1e-10 of the oracle on the same selection, each bitwise reproducible from run to run."""
    P, z = load_instance(name)
    pool = load_pool(name, {"storm": 16384, "ssn": 5000, "baa99-20": 1024}[name])[:K]
    vals = sampled_values_at(z, 5, np.arange(N))
    w = 0.5 + oracle.u01(9, np.arange(N))
    xs = (z["x_ev"], z["x_alt"])
    got = {}
    for mode in ("1", "2"):
        monkeypatch.setenv("SQLP_REDUCE", mode)
        ctx = T.Context(0)
        dvs = T.sdDualVertexSet(ctx=ctx, m2=P.m2)
        dvs.push_many(pool)
        epi = T.sdEpigraph(coef_of(T, P), 1.0, 0.0, dvs)
        epi.add_scenarios(vals, w)
        (c0, c1), val = epi.build_cuts2(*xs, with_val=True)
        (d0, d1), val2 = epi.build_cuts2(*xs, with_val=True)
        assert c0.alpha == d0.alpha and np.array_equal(c1.beta, d1.beta) and np.array_equal(val, val2)
        for x, cut, v in zip(xs, (c0, c1), val):
            mv, mi = epi.argmax(x)
            cut_vs_forced(oracle, P, vals, w, x, pool, mi, cut, v)
        one = epi.build_cut(xs[1], with_val=True)            # the single-point form
        assert abs(one[0].alpha - c1.alpha) <= 1e-12 * abs(c1.alpha)
        got[mode] = (c0, c1)
        epi.close(); dvs.close(); ctx.close()
    for a, b in zip(got["1"], got["2"]):
        assert abs(a.alpha - b.alpha) <= 1e-12 * abs(a.alpha)
        assert np.allclose(a.beta, b.beta, rtol=1e-11, atol=1e-9 * np.abs(a.beta).max())
