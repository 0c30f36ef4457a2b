"""GPU parity tests: the CUDA path (through the C ABI / the Python mirror of the reference's
interface) against the CPU oracle on the same inputs.  Bar: bit-exact for the dedup rule and
indices; argmax indices identical except where the score gap is below 1e-12 relative; cut
coefficients within 1e-10 relative (tolerances of BASELINE.json's north_star)."""
import numpy as np
import pytest

from tests.helpers import (check_argmax_parity, known_answers, load_instance,
                           sample_instance_values, synthetic_pool, synthetic_problem,
                           synthetic_values)

pytestmark = pytest.mark.gpu

CUT_RTOL = 1e-10


@pytest.fixture(scope="module")
def T():
    from sqlp_b200 import twosd
    return twosd


def coef_of(T, P):
    return T.sdSubprobCoefficients.from_tables(P.rbar, P.T_colptr, P.T_rowval, P.T_nzval,
                                               P.pos_row, P.pos_col)


def make_epi(T, P, pool, values, weights=None):
    dvs = T.sdDualVertexSet(m2=P.m2)
    if len(pool):
        dvs.push_many(pool)
    epi = T.sdEpigraph(coef_of(T, P), 1.0, 0.0, dvs)
    if len(values):
        epi.add_scenarios(values, weights)
    return dvs, epi


def check_cut(oracle, P, values, weights, x, pool, cut, val=None, epi=None):
    """Cut parity.  With ``epi`` given, the device's argmax at x is first checked against the
    oracle under the north-star rule, then the coefficients are compared on the device's
    (validated) selection -- real pools hold exact ties at their harvest point, and two
    valid selections give different (alpha, beta) with the same alpha + beta.x."""
    forced = None
    if epi is not None:
        mv, mi = epi.argmax(x)
        check_argmax_parity(P, values, x, pool, mv, mi)
        forced = mi
    ref = oracle.build_sasa_cut(P, values, weights, x, pool, forced_idx=forced)
    assert ref["status"] == 0
    if forced is not None:
        ref["max_idx"] = np.asarray(forced)
    p = np.asarray(weights) / ref["weight_mark"]
    Tm = P.T_dense()
    sel = pool[ref["max_idx"]]
    r_i = np.tile(P.rbar, (len(values), 1))
    for e in range(P.s):
        if P.pos_col[e] < 0:
            r_i[:, P.pos_row[e]] = values[:, e]
    sa = np.sum(p * np.abs(np.einsum("ij,ij->i", sel, r_i))) + 1e-300
    sb = (p[:, None] * np.abs(sel @ Tm)).sum(axis=0) + 1e-300
    assert abs(cut.alpha - ref["alpha"]) <= CUT_RTOL * max(sa, abs(ref["alpha"])), \
        (cut.alpha, ref["alpha"], sa)
    err = np.abs(cut.beta - ref["beta"])
    assert (err <= CUT_RTOL * np.maximum(sb, np.abs(ref["beta"])) + 1e-300).all(), err.max()
    assert cut.weight_mark == ref["weight_mark"]
    if val is not None:
        assert abs(val - ref["val"]) <= CUT_RTOL * (np.sum(p * np.abs(ref["max_val"])) + 1e-300)
    return ref


# ---- A5: dedup, bit exact ----------------------------------------------------------------

V1, V2, V3, V4, V5 = [1., 2, 3], [1.0000000001, 2, 3], [4., 5, 6], [4., 5, 6, 7], [3., 2, 1]


def test_dual_set_reference_cases(T):         # test/dual_set_test.jl:16-33, line for line
    dvs = T.sdDualVertexSet()
    sizes = []
    for v in (V1, V2, V3, V4, V5):             # V4 has another length: never equal, counted (dual_set.jl:26)
        assert T.push_(dvs, v) is dvs          # push! returns the set
        sizes.append(len(dvs))
    assert sizes == [1, 1, 2, 3, 4]
    dvs2 = T.sdDualVertexSet([V1, V2, V3, V4, V5])         # "Initialize with De-duplication"
    assert len(dvs2) == 4
    assert len(list(dvs2)) == 4                # "Iterate"
    assert [list(v) for v in dvs] == [V1, V3, V4, V5]      # insertion order, first copy kept
    assert [list(dvs[k]) for k in range(4)] == [V1, V3, V4, V5]
    # a second vector of the odd length is deduplicated among its own kind, and keeps the order
    assert dvs.push([4.0000000001, 5, 6, 7]) == (False, 2)
    assert dvs.push([9., 9, 9, 9]) == (True, 4) and len(dvs) == 5
    assert dvs.push(V5) == (False, 3)
    dvs.close(); dvs2.close()


def test_dedup_bit_exact_vs_oracle(T, oracle):
    rng = np.random.default_rng(11)
    m2, n = 40, 400
    base = rng.normal(size=(60, m2)) * rng.choice([1e-3, 1.0, 1e4], size=(60, 1))
    V = base[rng.integers(0, 60, size=n)].copy()
    # perturb below / around / above the 2^-16 relative resolution of the rule
    V *= 1.0 + rng.choice([0.0, 1e-12, 2.0 ** -17, 2.0 ** -16, 2.0 ** -15, 1e-3], size=(n, 1)) \
        * rng.choice([-1, 1], size=(n, m2))
    V[5, 3] = 0.0; V[6] = V[5]; V[6, 3] = -0.0            # +0 == -0
    V[7, 0] = np.nan; V[8] = V[7]                          # NaN never matches
    V[9] = 0.0; V[10] = 0.0                                # all-zero vectors
    V[11, :] = 5e-324; V[12, :] = 5e-324                   # subnormals pass through the round
    V[13, 0] = np.inf; V[14] = V[13]
    ref_pool, ref_ins, ref_idx = oracle.pool_push_many(m2, V)
    dvs = T.sdDualVertexSet(m2=m2)
    ins, idx = dvs.push_many(V)
    assert np.array_equal(ins, ref_ins.astype(bool))
    assert np.array_equal(idx, ref_idx)
    assert len(dvs) == len(ref_pool)
    got = np.stack(list(dvs))
    assert np.array_equal(got.view(np.uint64), ref_pool.view(np.uint64))
    for v in V[:40]:
        assert dvs.hash(v) == oracle.hash_dual_vector(v)
    # one-at-a-time pushes give the same answers as the batch
    dvs2 = T.sdDualVertexSet(m2=m2)
    for i in range(60):
        a, b = dvs2.push(V[i])
        assert (a, b) == (bool(ref_ins[i]), ref_idx[i])


def test_hash_gate_quirk(T):
    e = float.fromhex("0x1.0000e66666666p+0")
    dvs = T.sdDualVertexSet([[1., 1, 1], [e, e, e]])
    assert len(dvs) == 2        # every element rounds to 1.0 but the 1-norms round apart


def test_pool_growth_keeps_contents(T, oracle):
    m2 = 7
    V = synthetic_pool(m2, 2500)
    dvs = T.sdDualVertexSet(m2=m2)
    for lo in range(0, 2500, 300):             # crosses the initial 1024 capacity twice
        dvs.push_many(V[lo:lo + 300])
    assert len(dvs) == 2500
    for k in (0, 1023, 1024, 2499):
        assert np.array_equal(dvs[k], V[k])
    ins, idx = dvs.push_many(V[[5, 2400]])
    assert not ins.any() and list(idx) == [5, 2400]


# ---- lands known answers through the device path ----------------------------------------------

def test_lands_delta_and_eval_dual(T):        # sd_test.jl:36-41, 45-65
    ka = known_answers()
    P, z = load_instance("lands")
    rbar3 = P.rbar.copy(); rbar3[4] = 3.0
    coef3 = T.sdSubprobCoefficients.from_tables(rbar3, P.T_colptr, P.T_rowval, P.T_nzval,
                                                P.pos_row, P.pos_col)
    d = T.delta_coefficients(coef3, [5.0])
    assert d.delta_rhs[4] == 2.0 and np.count_nonzero(d.delta_rhs) == 1 and not d.delta_transfer
    coef = coef_of(T, P)
    e = ka["eval_dual"]
    for rhs, obj, dual in zip(e["rhs"], e["obj"], e["dual"]):
        assert T.eval_dual(coef, [rhs], e["x"], dual) == obj


def test_lands_named_scenarios_and_lookup_errors(T):
    P, _ = load_instance("lands")
    rows = {f"S2C{i + 1}": i for i in range(7)}
    cols = {f"X{i + 1}": i for i in range(4)}
    coef = T.sdSubprobCoefficients(P.rbar, P.T_colptr, P.T_rowval, P.T_nzval, 4, rows, cols,
                                   [("RHS", "S2C5")])
    assert coef.row_lookup["S2C5"] + 1 == 5 and coef.col_lookup["X2"] + 1 == 2   # sd_test.jl:17-18
    with pytest.raises(KeyError):
        coef.col_lookup["Y11"]                                                   # sd_test.jl:21
    d = T.delta_coefficients(coef, [(("RHS", "S2C5"), 5.0)])
    assert d.delta_rhs[4] == 5.0
    with pytest.raises(KeyError):
        T.delta_coefficients(coef, [(("RHS", "NOPE"), 5.0)])


def test_lands_argmax_equals_lp_objective(T):  # sd_test.jl:69-94
    ka = known_answers()["argmax"]
    P, _ = load_instance("lands")
    dvs = T.sdDualVertexSet(ka["pool"] + ka["pool"][:1])
    assert len(dvs) == 3
    coef = coef_of(T, P)
    epi = T.sdEpigraph(coef, 1.0, 0.0, dvs)
    for r in ka["scen_rhs"]:
        T.add_scenario_(epi, [r], 1.0)
    val, arg = T.argmax_procedure(coef, epi.scenario_delta, ka["x2"], dvs)
    assert list(val) == ka["lp_obj_at_x2"]
    for a, k in zip(arg, arg.index):
        assert np.array_equal(a, dvs[int(k)])


def test_lands_subgradient_and_sasa_closed_form(T):   # sd_test.jl:97-103, 207-235
    ka = known_answers()
    P, _ = load_instance("lands")
    sg = ka["subgradient"]
    dvs = T.sdDualVertexSet([sg["dual"]])
    epi = T.sdEpigraph(coef_of(T, P), 1.0, 0.0, dvs)
    T.add_scenario_(epi, [sg["rhs"]])
    assert list(T.build_sasa_cut(epi, sg["x"], dvs).beta) == sg["expect"]

    sa = ka["sasa"]
    d1, d2 = np.asarray(sa["my_dual"]), np.asarray(sa["my_dual_2_glpk_reconstructed"])
    dv = T.sdDualVertexSet([d1, d2])
    epi2 = T.sdEpigraph(coef_of(T, P), 0.5, 100.0, dv)
    T.add_scenario_(epi2, [3.0], 1.5)
    T.add_scenario_(epi2, [7.0], 0.5)
    x = np.asarray(sa["x"])
    assert [epi2.eval_dual(0, k, x) for k in (0, 1)] == sa["quoted_scores"]["scen3"]
    assert [epi2.eval_dual(1, k, x) for k in (0, 1)] == sa["quoted_scores"]["scen7"]
    Tm = P.T_dense()
    r1 = P.rbar.copy(); r1[4] = 3.0
    r2 = P.rbar.copy(); r2[4] = 7.0
    cut = T.build_sasa_cut(epi2, x, dv)
    assert cut.alpha == 1.5 / 2.0 * (d2 @ r1) + 0.5 / 2.0 * (d1 @ r2)
    assert np.array_equal(cut.beta, 1.5 / 2.0 * (-Tm.T @ d2) + 0.5 / 2.0 * (-Tm.T @ d1))
    assert cut.weight_mark == 2.0 and epi2.total_scenario_weight == 2.0


# ---- parity at the real instance shapes -------------------------------------------------------

@pytest.mark.parametrize("name,N", [("lands", 200), ("baa99-20", 600), ("ssn", 500), ("storm", 400)])
def test_real_instances_parity(T, oracle, name, N):
    P, z = load_instance(name)
    vals = sample_instance_values(z, N)
    w = 0.5 + oracle.u01(4, np.arange(N))
    extra = synthetic_pool(P.m2, 200, scale=float(np.abs(z["pool"]).max()))
    pool = np.vstack([z["pool"], extra])
    dvs, epi = make_epi(T, P, pool, vals, w)
    assert len(dvs) == len(pool)
    for x in (z["x_ev"], z["x_alt"]):
        mv, mi = epi.argmax(x)
        check_argmax_parity(P, vals, x, pool, mv, mi)
        cut, val = epi.build_cut(x, with_val=True)
        check_cut(oracle, P, vals, w, x, pool, cut, val, epi=epi)
    (cand, inc), val2 = epi.build_cuts2(z["x_ev"], z["x_alt"], with_val=True)
    check_cut(oracle, P, vals, w, z["x_ev"], pool, cand, val2[0], epi=epi)
    check_cut(oracle, P, vals, w, z["x_alt"], pool, inc, val2[1], epi=epi)
    # the fused two-point pass and the single-point pass agree bit for bit
    c1 = epi.build_cut(z["x_alt"])
    assert c1.alpha == inc.alpha and np.array_equal(c1.beta, inc.beta)


@pytest.mark.parametrize("N,K", [(1, 1), (127, 129), (128, 128), (129, 127), (1000, 300), (257, 1)])
def test_ragged_shapes(T, oracle, N, K):
    P = synthetic_problem(m2=50, n1=9, s=13)
    vals = synthetic_values(P, N)
    pool = synthetic_pool(P.m2, K)
    w = 0.5 + oracle.u01(4, np.arange(N))
    dvs, epi = make_epi(T, P, pool, vals, w)
    x = 10.0 * oracle.u01(3, np.arange(P.n1))
    mv, mi = epi.argmax(x)
    check_argmax_parity(P, vals, x, pool, mv, mi)
    check_cut(oracle, P, vals, w, x, pool, epi.build_cut(x))


@pytest.mark.parametrize("s", [1, 8, 20, 86, 117, 128])
def test_row_counts(T, oracle, s):
    P = synthetic_problem(m2=s + 30, n1=11, s=s, first_stoch_row=7)
    N, K = 300, 260
    vals = synthetic_values(P, N)
    pool = synthetic_pool(P.m2, K)
    dvs, epi = make_epi(T, P, pool, vals)
    x = 10.0 * oracle.u01(3, np.arange(P.n1))
    mv, mi = epi.argmax(x)
    check_argmax_parity(P, vals, x, pool, mv, mi)
    check_cut(oracle, P, vals, np.ones(N), x, pool, epi.build_cut(x))


def test_delta_T_path(T, oracle):
    """Random elements on Tbar (no shipped instance has them): d(x) is rebuilt per point."""
    P = synthetic_problem(m2=60, n1=14, s=22, n_T=8)
    N, K = 700, 333
    vals = synthetic_values(P, N)
    pool = synthetic_pool(P.m2, K)
    w = 0.5 + oracle.u01(4, np.arange(N))
    dvs, epi = make_epi(T, P, pool, vals, w)
    xs = [10.0 * oracle.u01(3, np.arange(P.n1)), 10.0 * oracle.u01(5, np.arange(P.n1))]
    for x in xs:
        mv, mi = epi.argmax(x)
        check_argmax_parity(P, vals, x, pool, mv, mi)
        check_cut(oracle, P, vals, w, x, pool, epi.build_cut(x))
    (cand, inc) = epi.build_cuts2(xs[0], xs[1])
    check_cut(oracle, P, vals, w, xs[0], pool, cand)
    check_cut(oracle, P, vals, w, xs[1], pool, inc)
    # delta readback and eval_dual in the reference's order are bit exact
    for i in (0, 5, N - 1):
        drhs, dT = oracle.delta_coefficients(P, vals[i])
        d = epi.delta(i)
        assert np.array_equal(d.delta_rhs, drhs)
        for e in range(P.s):
            if P.pos_col[e] >= 0:
                assert d.delta_transfer[(int(P.pos_row[e]), int(P.pos_col[e]))] == dT[e]
        for k in (0, K - 1):
            assert epi.eval_dual(i, k, xs[0]) == oracle.eval_dual(P, vals[i], xs[0], pool[k])


def test_first_index_on_ties_nan_and_empty_pool(T, oracle):
    P, _ = load_instance("lands")
    v = np.array([-4., -1, -12, 0, 44, 28, 5.5])
    # slots 0 and 2 tie exactly (slot 2 differs only far below the dedup resolution? no:
    # it must be a distinct vertex, so tie them through a row that multiplies zero)
    v2 = v.copy(); v2[3] = -7.0          # row S2C4: base = 0 - (-1)*x4, make x4 = 0
    nanv = np.full(7, np.nan)
    dvs = T.sdDualVertexSet([v2, nanv, v, v - 1.0])
    assert len(dvs) == 4
    epi = T.sdEpigraph(coef_of(T, P), 1.0, 0.0, dvs)
    epi.add_scenarios([[5.0], [3.0], [7.0]])
    x = np.array([3.0, 3.0, 3.0, 0.0])
    mv, mi = epi.argmax(x)
    ov, oi = oracle.argmax_procedure(P, [[5.0], [3.0], [7.0]], x, np.stack(list(dvs)))
    assert list(mi) == list(oi) == [0, 0, 0]       # strict '>' keeps the first maximum
    assert list(mv) == list(ov)
    # empty pool / all-NaN pool: UndefRefError in the reference
    for pool in ([], [nanv]):
        d = T.sdDualVertexSet(pool, m2=7)
        e = T.sdEpigraph(coef_of(T, P), 1.0, 0.0, d)
        e.add_scenarios([[5.0]])
        mv, mi = e.argmax(x)
        assert mi[0] == -1 and mv[0] == -np.inf
        with pytest.raises(T.NoArgmaxError):
            e.build_cut(x)
        with pytest.raises(T.NoArgmaxError):
            T.argmax_procedure(e.subproblem_coef, e.scenario_delta, x, d)
    with pytest.raises(T.SqlpError) as ei:
        epi.argmax(x, sense=T.MAX_SENSE)
    assert ei.value.code == -3
    # no scenarios: the reference's loop does not run -> sdCut(0, zeros, 0)
    d = T.sdDualVertexSet([v])
    e = T.sdEpigraph(coef_of(T, P), 1.0, 0.0, d)
    cut = e.build_cut(x)
    assert cut.alpha == 0.0 and not cut.beta.any() and cut.weight_mark == 0.0


def test_incremental_growth_matches_bulk_and_is_deterministic(T, oracle):
    """The SD loop's pattern: one scenario and two pushes per iteration (algorithm.jl:45-55)."""
    P, z = load_instance("baa99-20")
    N = 150
    vals = sample_instance_values(z, N)
    pool = z["pool"]
    dvs = T.sdDualVertexSet(m2=P.m2)
    epi = T.sdEpigraph(coef_of(T, P), 1.0, 0.0, dvs)
    x = z["x_alt"]
    for i in range(N):
        T.add_scenario_(epi, vals[i], 1.0)
        T.push_(dvs, pool[(2 * i) % len(pool)])
        T.push_(dvs, pool[(2 * i + 1) % len(pool)])
        if i in (0, 3, 77, N - 1):
            Kn = len(dvs)
            cut = T.build_sasa_cut(epi, x, dvs)
            check_cut(oracle, P, vals[:i + 1], np.ones(i + 1), x, np.stack(list(dvs))[:Kn], cut, epi=epi)
    assert len(dvs) == len(pool)
    dvs_b, epi_b = make_epi(T, P, np.stack(list(dvs)), vals)
    a, b = epi.build_cut(x), epi_b.build_cut(x)
    assert a.alpha == b.alpha and np.array_equal(a.beta, b.beta)
    for _ in range(3):                      # run-to-run bitwise identical
        c = epi.build_cut(x)
        assert c.alpha == a.alpha and np.array_equal(c.beta, a.beta)


def test_device_sampling_matches_host_sampling(T, oracle):
    P, z = load_instance("storm")
    N = 300
    vals = sample_instance_values(z, N, seed=21)
    dvs, epi = make_epi(T, P, z["pool"], [])
    epi.set_outcomes(z["out_vals"], z["out_cdf"], z["out_cnt"])
    epi.sample_scenarios(N, seed=21, weight_seed=4)
    ng, nl, tw = epi.counts()
    w = 0.5 + oracle.u01(4, np.arange(N))
    tot = 0.0
    for x in w:
        tot += float(x)
    assert (ng, nl, tw) == (N, N, tot)
    for i in (0, 1, 127, 128, N - 1):
        drhs, _ = oracle.delta_coefficients(P, vals[i])
        assert np.array_equal(epi.delta(i).delta_rhs, drhs)
    x = z["x_alt"]
    check_cut(oracle, P, vals, w, x, z["pool"], epi.build_cut(x), epi=epi)


def test_device_sampling_of_normal_and_uniform_elements(T, oracle):
    """INDEP NORMAL / UNIFORM elements (smps_sto.jl:118-127; transship is all NORMAL) mixed with
    DISCRETE ones: the device draws agree with the host twin to rounding (the normal quantile is
    normcdfinv on the device, scipy's ndtri on the host), the moments are right, and the cut built
    from the device's scenarios matches the oracle fed the twin's values."""
    P = synthetic_problem(m2=70, n1=12, s=24)
    s, N = P.s, 1000
    kind = np.arange(s) % 3
    base = P.rbar[P.pos_row]
    par_a = np.where(kind == 1, base, 0.8 * base)                 # mean | left
    par_b = np.where(kind == 1, (0.1 * base) ** 2, 1.2 * base)    # variance | right
    mo = 5
    out_vals = base[:, None] * np.array([0.8, 0.9, 1.0, 1.1, 1.2])[None, :]
    out_cdf = np.tile(np.array([0.1, 0.3, 0.6, 0.8, 1.0]), (s, 1))
    out_cnt = np.full(s, mo, dtype=np.int32)
    pool = synthetic_pool(P.m2, 200)
    dvs, epi = make_epi(T, P, pool, [])
    epi.set_outcomes(out_vals, out_cdf, out_cnt)
    epi.set_distributions(kind, par_a, par_b)
    epi.sample_scenarios(N, seed=33)
    twin = oracle.sample_twin(33, 0, N, kind, par_a, par_b, out_vals, out_cdf, out_cnt)
    got = np.array([epi.delta(i).delta_rhs[P.pos_row] + base for i in range(0, N, 7)])
    ref = twin[::7]
    assert np.array_equal(got[:, kind == 0], ref[:, kind == 0])                  # discrete: exact
    assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)) < 1e-12
    z = (twin[:, kind == 1] - par_a[kind == 1]) / np.sqrt(par_b[kind == 1])
    assert abs(z.mean()) < 0.05 and abs(z.std() - 1.0) < 0.05
    uu = (twin[:, kind == 2] - par_a[kind == 2]) / (par_b - par_a)[kind == 2]
    assert 0.0 < uu.min() and uu.max() < 1.0 and abs(uu.mean() - 0.5) < 0.02
    x = 10.0 * oracle.u01(3, np.arange(P.n1))
    check_cut(oracle, P, twin, np.ones(N), x, pool, epi.build_cut(x), epi=epi)
    # an all-continuous table needs no outcome tables
    dvs2, epi2 = make_epi(T, P, pool, [])
    epi2.set_distributions(np.full(s, 2), par_a, par_b)
    epi2.sample_scenarios(10, seed=5)
    assert epi2.counts()[0] == 10


def test_multi_epigraph_cell_call(T, oracle):
    """E = 4 weighted epigraphs sharing one pool, one library call."""
    P, z = load_instance("storm")
    pool = z["pool"]
    dvs = T.sdDualVertexSet(m2=P.m2)
    dvs.push_many(pool)
    E, N = 4, 520
    vals = sample_instance_values(z, N)
    w = 0.5 + oracle.u01(4, np.arange(N))
    epis = [T.sdEpigraph(coef_of(T, P), 0.25, 0.0, dvs) for _ in range(E)]
    for e in range(E):
        epis[e].add_scenarios(vals[e::E], w[e::E])
    out = T.build_cuts_at_candidate_and_incumbent(epis, z["x_ev"], z["x_alt"])
    for e in range(E):
        check_cut(oracle, P, vals[e::E], w[e::E], z["x_ev"], pool, out[e][0], epi=epis[e])
        check_cut(oracle, P, vals[e::E], w[e::E], z["x_alt"], pool, out[e][1], epi=epis[e])
        assert epis[e].incumbent_cut is out[e][1] and epis[e].cuts[-1] is out[e][0]


def test_storm_shape_properties_at_scale(T, oracle):
    """Full-width storm shape at a size the oracle cannot sweep: size-independent checks
    (G5: alpha + beta.x == sum p_i maxval_i; template-shift invariance G6) plus an oracle
    spot check on a random subset of scenarios."""
    P, z = load_instance("storm")
    N, K = 40000, 2048
    pool = np.vstack([z["pool"], synthetic_pool(P.m2, K - len(z["pool"]), scale=800.0)])
    dvs, epi = make_epi(T, P, pool, [])
    epi.set_outcomes(z["out_vals"], z["out_cdf"], z["out_cnt"])
    epi.sample_scenarios(N, seed=1, weight_seed=4)
    xs = (z["x_ev"], z["x_alt"])
    (cand, inc), val = epi.build_cuts2(*xs, with_val=True)
    for x, cut, v in zip(xs, (cand, inc), val):
        assert abs(cut.alpha + cut.beta @ x - v) <= 1e-10 * (abs(cut.alpha) + np.abs(cut.beta * x).sum())
    mv, mi = epi.argmax(xs[1])
    assert (mi >= 0).all() and (mi < K).all()
    vals = sample_instance_values(z, N, seed=1)
    pick = np.random.default_rng(0).choice(N, size=48, replace=False)
    check_argmax_parity(P, vals[pick], xs[1], pool, mv[pick], mi[pick])
    # G6: an epigraph whose template was frozen at another rbar gives the same cut
    rbar2 = P.rbar.copy(); rbar2[P.pos_row] += 3.25
    P2 = oracle.Problem(P.m2, P.n1, rbar2, P.T_colptr, P.T_rowval, P.T_nzval, P.pos_row, P.pos_col)
    epi2 = T.sdEpigraph(coef_of(T, P2), 1.0, 0.0, dvs)
    epi2.set_outcomes(z["out_vals"], z["out_cdf"], z["out_cnt"])
    epi2.sample_scenarios(N, seed=1, weight_seed=4)
    mv2, mi2 = epi2.argmax(xs[1])
    assert np.mean(mi2 == mi) > 0.9999
    c2 = epi2.build_cut(xs[1])
    assert abs(c2.alpha - inc.alpha) <= 1e-9 * abs(inc.alpha)


# ---- contraction plans: the even split of (unit, chunk) work and the streaming fallback ----

@pytest.mark.parametrize("plan", ["grid1", "grid3", "grid7", "grid50", "grid1000", "stream", "resident"])
@pytest.mark.parametrize("N,K,s", [(1000, 300, 13), (700, 1500, 117), (129, 127, 20), (3000, 129, 86)])
def test_contraction_plans(T, oracle, monkeypatch, plan, N, K, s):
    """Forced grid sizes cut the units at many different places (one CTA doing everything,
    several CTAs per unit, more CTAs than chunk-units) and the two fallback kernels replace the
    warp-specialised one; every plan must give the oracle's argmax and, bit for bit, the values
    and indices of the default plan."""
    base_ctx = T.default_context()           # created without overrides
    if plan in ("stream", "resident"):
        monkeypatch.setenv("SQLP_CONTRACT", plan)
    else:
        monkeypatch.setenv("SQLP_CONTRACT_GRID", plan[4:])
    forced_ctx = T.Context(0)                # reads the overrides at creation
    monkeypatch.delenv("SQLP_CONTRACT", raising=False)
    monkeypatch.delenv("SQLP_CONTRACT_GRID", raising=False)
    P = synthetic_problem(m2=s + 30, n1=11, s=s, first_stoch_row=5)
    vals = synthetic_values(P, N)
    pool = synthetic_pool(P.m2, K)
    x2 = [10.0 * oracle.u01(3, np.arange(P.n1)), 10.0 * oracle.u01(5, np.arange(P.n1))]
    out = []
    for ctx in (forced_ctx, base_ctx):       # the forced plan, then the default one
        dvs = T.sdDualVertexSet(ctx=ctx, m2=P.m2)
        dvs.push_many(pool)
        epi = T.sdEpigraph(coef_of(T, P), 1.0, 0.0, dvs)
        epi.add_scenarios(vals)
        res = [epi.argmax(x) for x in x2]
        (c0, c1), val = epi.build_cuts2(x2[0], x2[1], with_val=True)
        out.append((res, c0, c1))
        for x, (mv, mi) in zip(x2, res):
            check_argmax_parity(P, vals, x, pool, mv, mi)
    (ra, a0, a1), (rb, b0, b1) = out
    for (mva, mia), (mvb, mib) in zip(ra, rb):
        assert (mia == mib).all() and (mva == mvb).all()
    assert a0.alpha == b0.alpha and (a0.beta == b0.beta).all() and a1.alpha == b1.alpha


@pytest.mark.parametrize("s", [256, 400])
def test_wide_row_sets_use_the_streaming_kernel(T, oracle, s):
    """One unit of scenarios no longer fits in shared memory: the streaming kernel takes over."""
    P = synthetic_problem(m2=s + 30, n1=11, s=s, first_stoch_row=7)
    N, K = 200, 140
    vals = synthetic_values(P, N)
    pool = synthetic_pool(P.m2, K)
    dvs, epi = make_epi(T, P, pool, vals)
    x = 10.0 * oracle.u01(3, np.arange(P.n1))
    mv, mi = epi.argmax(x)
    check_argmax_parity(P, vals, x, pool, mv, mi)
    check_cut(oracle, P, vals, np.ones(N), x, pool, epi.build_cut(x))


def test_scenario_permutation_changes_cuts_only_at_rounding_level(T, oracle):
    """the cut is a weighted sum over scenarios, so a permutation of the scenarios
    (with their weights) permutes the argmax and moves (alpha, beta) by rounding only."""
    P, z = load_instance("ssn")
    N = 700
    vals = sample_instance_values(z, N, seed=3)
    w = 0.5 + oracle.u01(4, np.arange(N))
    perm = np.argsort(oracle.u01(11, np.arange(N)))
    x = z["x_alt"]
    _, e1 = make_epi(T, P, z["pool"], vals, w)
    _, e2 = make_epi(T, P, z["pool"], vals[perm], w[perm])
    (mv1, mi1), (mv2, mi2) = e1.argmax(x), e2.argmax(x)
    assert np.array_equal(mi1[perm], mi2) and np.array_equal(mv1[perm], mv2)      # bit for bit
    c1, c2 = e1.build_cut(x), e2.build_cut(x)
    assert abs(c1.alpha - c2.alpha) <= 1e-12 * max(1.0, abs(c1.alpha))
    assert np.max(np.abs(c1.beta - c2.beta)) <= 1e-12 * max(1.0, np.abs(c1.beta).max())
    assert abs(e1.total_scenario_weight - e2.total_scenario_weight) <= 1e-12 * e1.total_scenario_weight


def test_device_profiler_classes(T, oracle):
    """sqlp_ctx_profile_classes: every kernel class reports its scopes, device time and algorithmic work."""
    P, z = load_instance("storm")
    ctx = T.Context(0)
    dvs = T.sdDualVertexSet(ctx=ctx, m2=P.m2)
    epi = T.sdEpigraph(coef_of(T, P), 1.0, 0.0, dvs)
    ctx.profile(True)
    ctx.profile_classes(reset=True)
    dvs.push_many(z["pool"])
    N = 300
    epi.add_scenarios(sample_instance_values(z, N))
    epi.build_cuts2(z["x_ev"], z["x_alt"])
    prof = ctx.profile_classes(reset=True)
    ctx.profile(False)
    K, s = len(z["pool"]), P.s
    cols, _ = epi.view_columns()                 # flops are counted with the columns the launch swept: one per class
    assert cols <= K                             # of score-equivalent vertices (92 of these 94 harvested duals)
    assert prof["contract"][1] == 1 and prof["contract"][2] == 2.0 * s * cols * N
    assert prof["delta"][1] == 1 and prof["delta"][2] == 16.0 * s * N
    assert prof["reduce"][1] == 1 and prof["pool"][1] == 1 and prof["bias"][1] == 1
    assert all(prof[k][0] > 0 for k in ("contract", "delta", "reduce", "pool", "bias"))
    assert prof["screen"][1] == 0 and prof["fallback"][1] == 0      # a shape this small is swept in FP64
    assert ctx.profile_classes()["contract"][1] == 0         # reset


def test_degenerate_shapes(T, oracle):
    """No scenarios yet (the reference's loops run zero times: alpha = 0, beta = 0, weight_mark = 0),
    a 1 x 1 x 1 problem, and a template without any random element (s = 0)."""
    P = synthetic_problem(m2=30, n1=5, s=6)
    dvs, epi = make_epi(T, P, synthetic_pool(P.m2, 10), [])
    cut = epi.build_cut(np.ones(5))
    assert cut.alpha == 0.0 and (cut.beta == 0.0).all() and cut.weight_mark == 0.0
    mv, mi = epi.argmax(np.ones(5))
    assert mv.shape == (0,) and mi.shape == (0,)

    P1 = oracle.Problem(3, 1, np.array([1.0, 2.0, 0.0]), np.array([0, 1]), np.array([2]), np.array([-1.0]),
                        np.array([0], dtype=np.int32), np.array([-1], dtype=np.int32))
    pool1 = np.array([[2.0, -1.0, 0.5]])
    _, e1 = make_epi(T, P1, pool1, np.array([[4.0]]))
    check_cut(oracle, P1, np.array([[4.0]]), np.ones(1), np.array([3.0]), pool1, e1.build_cut(np.array([3.0])))

    P0 = oracle.Problem(4, 2, np.array([1.0, 0, 2, 0]), np.array([0, 1, 2]), np.array([1, 3]), np.array([-1.0, -2.0]),
                        np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32))
    pool0 = synthetic_pool(4, 3)
    _, e0 = make_epi(T, P0, pool0, np.zeros((5, 0)))
    x = np.array([1.0, 2.0])
    ref = oracle.build_sasa_cut(P0, np.zeros((5, 0)), np.ones(5), x, pool0)
    cut = e0.build_cut(x)
    assert abs(cut.alpha - ref["alpha"]) <= 1e-12 * abs(ref["alpha"]) and np.allclose(cut.beta, ref["beta"], rtol=1e-12)
    assert cut.weight_mark == 5.0


def test_c_abi_from_plain_c(tmp_path):
    """tests/abi_demo.c, compiled with gcc against include/sqlp_b200.h and linked to the library: the
    reference's weighted build_sasa_cut case (test/sd_test.jl:207-235) without Python in the loop."""
    import json
    import os
    import subprocess
    from sqlp_b200 import _lib
    here = os.path.dirname(os.path.abspath(__file__))
    exe = str(tmp_path / "abi_demo")
    libdir = os.path.dirname(_lib.SO_PATH)
    subprocess.run(["gcc", "-O1", "-o", exe, os.path.join(here, "abi_demo.c"), "-L", libdir, "-lsqlp_b200",
                    f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    r = json.loads(out.stdout)
    assert r["K"] == 2 and r["inserted"] == [1, 1, 0] and r["index"] == [0, 1, 0]
    # scores 168 vs 169 -> my_dual_2 for RHS = 3; 344 vs 331 -> my_dual for RHS = 7 (sd_test.jl:216-222)
    assert r["max_idx"] == [1, 0] and r["max_val"] == [169.0, 344.0]
    d1 = np.array([-4.0, -1.0, -12.0, -0.0, 44.0, 28.0, 5.5]); d2 = np.array([-0.5, 0.0, -8.5, 0.0, 40.5, 24.5, 4.5])
    r1 = np.array([0, 0, 0, 0, 3.0, 3.0, 2.0]); r2 = np.array([0, 0, 0, 0, 7.0, 3.0, 2.0])
    Tm = np.zeros((7, 4)); Tm[[0, 1, 2, 3], [0, 1, 2, 3]] = -1.0
    assert r["alpha"] == 1.5 / 2.0 * (d2 @ r1) + 0.5 / 2.0 * (d1 @ r2)            # sd_test.jl:229,233 (exact ==)
    assert r["beta"] == list(1.5 / 2.0 * (-Tm.T @ d2) + 0.5 / 2.0 * (-Tm.T @ d1))   # :230,234
    assert r["weight_mark"] == 2.0 and r["max_sense_status"] == -3
