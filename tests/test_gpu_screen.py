"""The screening pass (tcgen05, csrc/kernels_screen.cuh) against the FP64 sweep it stands in for.

Exactness is the bar, not a tolerance: with the pass forced on (``ctx.set_screen(2)``) ``argmax_procedure`` must
return the SAME indices and the SAME values, bit for bit, as with the pass off (``set_screen(0)``: every score in
FP64, reference ``subprob.jl:148-166``), and so must the cuts built from them -- on real pools, synthetic pools,
exact duplicates and near-ties, NaN / Inf vertices, empty pools, ragged sizes, few scenarios (the sweep split in
K-ranges) and many.  Parity of the FP64 sweep itself with the oracle is the business of the other test files.
"""
import numpy as np
import pytest

from tests.helpers import (load_instance, load_pool, sample_instance_values, sampled_values_at, synthetic_pool,
                           synthetic_problem, synthetic_values)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    from sqlp_b200 import twosd
    return twosd


@pytest.fixture()
def ctx(T):
    c = T.default_context()
    yield c
    c.set_screen(1)


def coef_of(T, P):
    return T.sdSubprobCoefficients.from_tables(P.rbar, P.T_colptr, P.T_rowval, P.T_nzval, P.pos_row, P.pos_col)


def both_ways(T, ctx, P, pool, vals, w, xs, expect_fallback=None):
    dvs = T.sdDualVertexSet(m2=P.m2)
    if len(pool):
        dvs.push_many(pool)
    epi = T.sdEpigraph(coef_of(T, P), 1.0, 0.0, dvs)
    epi.add_scenarios(vals, w)
    res = {}
    for mode in (0, 2):
        ctx.set_screen(mode)
        got = []
        for x in xs:
            got.append(epi.argmax(x))
        try:
            cuts, val = epi.build_cuts2(xs[0], xs[-1], with_val=True)
            got.append((np.array([cuts[0].alpha, cuts[1].alpha]), np.stack([cuts[0].beta, cuts[1].beta]), val))
        except T.NoArgmaxError:
            got.append("no argmax")
        res[mode] = got
    st = epi.screen_stats()
    assert st["passes"] >= 1, "the screening pass never ran"
    for a, b in zip(res[0], res[2]):
        if isinstance(a, str):
            assert a == b
            continue
        for u, v in zip(a, b):
            u, v = np.asarray(u), np.asarray(v)
            assert u.dtype == v.dtype and u.shape == v.shape
            assert np.array_equal(u.view(np.uint8), v.view(np.uint8)), (u, v)
    if expect_fallback is not None:
        assert (st["fallbacks"] > 0) == expect_fallback, st
    epi.close()
    dvs.close()
    return st


@pytest.mark.parametrize("name,K,N", [("baa99-20", 1024, 3000), ("ssn", 5000, 20000), ("storm", 2048, 5000)])
def test_real_pools_bit_identical(T, ctx, name, K, N):
    P, z = load_instance(name)
    pool = load_pool(name, {"baa99-20": 1024, "ssn": 5000, "storm": 16384}[name])[:K]
    vals = sampled_values_at(z, 1, np.arange(N))
    w = 0.5 + np.arange(N) % 7 / 7.0
    st = both_ways(T, ctx, P, pool, vals, w, (z["x_ev"], z["x_alt"]))
    print(name, st)


def test_storm_real_pool_full_size(T, ctx):
    """Storm's real duals tie by the dozen at full K.  (Before the operands were centred the candidate lists
    overflowed here and the pass handed over to the FP64 sweep; `test_crowd_overflows_and_falls_back` keeps that
    path under test.)"""
    P, z = load_instance("storm")
    pool = load_pool("storm", 16384)
    N = 4000
    vals = sampled_values_at(z, 7, np.arange(N))
    st = both_ways(T, ctx, P, pool, vals, np.ones(N), (z["x_ev"], z["x_alt"]))
    print("storm K=16384 real:", st)


@pytest.mark.parametrize("N,K,s", [(1000, 300, 13), (700, 1500, 117), (129, 127, 20), (3000, 129, 86), (5, 2000, 40),
                                   (40000, 4096, 117)])
def test_synthetic_shapes_bit_identical(T, ctx, N, K, s):
    P = synthetic_problem(m2=max(64, s + 11), n1=16, s=s, first_stoch_row=3)
    vals = synthetic_values(P, N, seed=3)
    pool = synthetic_pool(P.m2, K, seed=5, scale=700.0)
    x = 5.0 * np.cos(np.arange(P.n1))
    both_ways(T, ctx, P, pool, vals, None, (x, 2.0 * x + 1.0), expect_fallback=False)


def test_adversarial_pool(T, ctx):
    """Exact duplicates (on the stochastic rows), 2^-40 near-ties, a bias 1e6 above the dots, zero scenarios."""
    P = synthetic_problem(m2=80, n1=12, s=30, first_stoch_row=5)
    N, K = 2500, 1200
    vals = synthetic_values(P, N, seed=9)
    vals[7] = P.rbar[P.pos_row]                                    # a scenario with d = 0
    pool = synthetic_pool(P.m2, K, seed=6, scale=300.0)
    S = P.pos_row
    for k in range(100, 400, 3):                                   # same stochastic rows, same bias: exact ties
        pool[k] = pool[k - 100]
        pool[k, 0] += 1.0                                          # distinct for the dedup rule ...
    other = [j for j in range(P.m2) if j not in set(S.tolist())]
    base = P.rbar - P.T_dense() @ np.ones(P.n1)
    for k in range(100, 400, 3):
        j = next(j for j in other if j != 0 and abs(base[j]) > 1e-3)
        pool[k, j] -= base[0] / base[j]                            # ... with (almost) the same bias
    for k in range(500, 700):
        pool[k] = pool[k - 500] * (1.0 + 2.0 ** -40)               # near-ties (kept apart by column 1 below)
        pool[k, 1] += 1e-3 * k
    pool[800:900, other[1]] += 1e6 / max(abs(base[other[1]]), 1e-3)    # biases far above the dots
    x = np.ones(P.n1)
    both_ways(T, ctx, P, pool, vals, None, (x, 0.5 * x))


def test_nonfinite_vertices_and_empty_pool(T, ctx):
    P = synthetic_problem(m2=64, n1=10, s=24)
    N = 900
    vals = synthetic_values(P, N, seed=2)
    pool = synthetic_pool(P.m2, 600, seed=8)
    x = np.linspace(0.0, 3.0, P.n1)
    p1 = pool.copy(); p1[17, P.pos_row[3]] = np.nan               # NaN on a stochastic row: the pass steps aside
    both_ways(T, ctx, P, p1, vals, None, (x, x + 1.0), expect_fallback=True)
    p2 = pool.copy(); p2[5, 60] = np.inf                           # Inf off the stochastic rows: a +-Inf / NaN bias
    both_ways(T, ctx, P, p2, vals, None, (x, x + 1.0))
    p3 = pool.copy(); p3[9, 61] = np.nan
    both_ways(T, ctx, P, p3, vals, None, (x, x + 1.0))
    both_ways(T, ctx, P, pool[:0], vals, None, (x, x + 1.0))       # empty pool: nothing beats -Inf
    both_ways(T, ctx, P, np.full((3, P.m2), np.nan), vals, None, (x, x + 1.0))


def test_growing_pool_and_scenarios(T, ctx):
    """The bf16 operands follow pushes and add_scenario! like the FP64 view does (an SD run: one scenario and two
    vertices per iteration)."""
    P, z = load_instance("ssn")
    pool = load_pool("ssn", 5000)
    dvs = T.sdDualVertexSet(m2=P.m2)
    dvs.push_many(pool[:1500])
    epi = T.sdEpigraph(coef_of(T, P), 1.0, 0.0, dvs)
    vals = sampled_values_at(z, 2, np.arange(4000))
    epi.add_scenarios(vals[:3000], None)
    x = z["x_ev"]
    for it in range(6):
        dvs.push_many(pool[1500 + 40 * it: 1500 + 40 * (it + 1)])
        epi.add_scenarios(vals[3000 + 100 * it: 3000 + 100 * (it + 1)], None)
        ctx.set_screen(0)
        a = epi.argmax(x)
        ctx.set_screen(2)
        b = epi.argmax(x)
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[0].view(np.uint64), b[0].view(np.uint64))
    assert epi.screen_stats()["passes"] >= 6


def test_warm_start_from_previous_winners(T, ctx):
    """From the second pass on the scan of a scenario starts from the score of the vertices that won it at the
    previous pass (k_screen_seed) instead of from -Inf.  An SD-like run on storm's real pool -- the candidate moves
    every iteration (a little, or across the whole box), the incumbent now and then, vertices and scenarios arrive --
    must stay bit-identical to the FP64 sweep, and the lists must get shorter."""
    P, z = load_instance("storm")
    pool = load_pool("storm", 16384)
    dvs = T.sdDualVertexSet(m2=P.m2)
    dvs.push_many(pool[:5000])
    epi = T.sdEpigraph(coef_of(T, P), 1.0, 0.0, dvs)
    N0 = 40000            # enough units of 128 scenarios to fill the GPU without K-ranges: the row-staging decision runs
    vals = sampled_values_at(z, 9, np.arange(N0 + 1000))
    epi.add_scenarios(vals[:N0], 0.5 + np.arange(N0) % 3)
    x0, x1 = z["x_ev"], z["x_alt"]
    inc = x0
    emitted = []
    for it in range(9):
        dvs.push_many(pool[5000 + 30 * it: 5000 + 30 * (it + 1)])
        epi.add_scenarios(vals[N0 + 100 * it: N0 + 100 * (it + 1)], None)
        lam = [0.0, 0.02, 0.05, 1.0, 0.97, 0.5, 0.48, -0.3, 0.1][it]           # small steps and jumps
        cand = x0 + lam * (x1 - x0)
        if it % 4 == 3:
            inc = x0 + [0.0, 0.02, 0.05][it // 4] * (x1 - x0)                  # the incumbent becomes an earlier candidate
        res = {}
        for mode in (0, 2):
            ctx.set_screen(mode)
            cuts, val = epi.build_cuts2(cand, inc, with_val=True)
            if mode == 2:
                emitted.append(epi.screen_stats()["emitted"])      # of the two-point pass just run
            mv, mi = epi.argmax(cand)
            res[mode] = (np.array([cuts[0].alpha, cuts[1].alpha]), np.stack([cuts[0].beta, cuts[1].beta]), val, mv, mi)
        for u, v in zip(res[0], res[2]):
            u, v = np.asarray(u), np.asarray(v)
            assert np.array_equal(u.view(np.uint8), v.view(np.uint8)), it
    st = epi.screen_stats()
    assert st["fallbacks"] == 0 and st["passes"] >= 18, st
    print("emitted per pass:", emitted)
    assert emitted[2] < emitted[0]          # a small step of the candidate: far fewer stale entries than the cold scan


def test_recentring_of_both_operands(T, ctx):
    """The centres of the pool view and of the scenario set are taken again when four times as many columns /
    scenarios are there (until 1 024 / 256): every bf16 operand, `ctr . d_i` and the warm-start state have to follow.
    A pool growing 60 -> 4 100 vertices under an epigraph growing 40 -> 20 100 scenarios, the pass forced at every
    stage, two calls per stage (the second one warm-started where the sweep is not split), against the FP64 sweep."""
    P, z = load_instance("storm")
    pool = load_pool("storm", 16384)
    vals = sampled_values_at(z, 11, np.arange(20100))
    dvs = T.sdDualVertexSet(m2=P.m2)
    epi = T.sdEpigraph(coef_of(T, P), 1.0, 0.0, dvs)
    k_have = n_have = 0
    x0, x1 = z["x_ev"], z["x_alt"]
    for stage, (k_to, n_to) in enumerate(((60, 40), (250, 200), (1000, 1000), (4100, 20000), (4100, 20100))):
        if k_to > k_have:
            dvs.push_many(pool[k_have:k_to])
        epi.add_scenarios(vals[n_have:n_to], 0.5 + np.arange(n_to - n_have) % 5)
        k_have, n_have = k_to, n_to
        for rep, lam in enumerate((0.1 * stage, 0.1 * stage + 0.03)):
            cand = x0 + lam * (x1 - x0)
            res = {}
            for mode in (0, 2):
                ctx.set_screen(mode)
                cuts, val = epi.build_cuts2(cand, x1, with_val=True)
                mv, mi = epi.argmax(cand)
                res[mode] = (np.array([cuts[0].alpha, cuts[1].alpha]), np.stack([cuts[0].beta, cuts[1].beta]), val, mv, mi)
            for u, v in zip(res[0], res[2]):
                u, v = np.asarray(u), np.asarray(v)
                assert np.array_equal(u.view(np.uint8), v.view(np.uint8)), (stage, rep)
    st = epi.screen_stats()
    assert st["passes"] >= 20 and st["bad_operands"] == 0 and st["overflowed_lists"] == 0, st


CROWD_SCRIPT = """
import sys
import numpy as np
sys.path.insert(0, %(root)r)
from sqlp_b200 import twosd as T
from tests.helpers import synthetic_pool, synthetic_problem, synthetic_values
P = synthetic_problem(m2=80, n1=12, s=30, first_stoch_row=5)
N, K = 3000, 600
vals = synthetic_values(P, N, seed=4)
base_v = synthetic_pool(P.m2, 1, seed=3, scale=300.0)[0]
rng = np.random.default_rng(5)
pool = np.tile(base_v, (K, 1)) * (1.0 + 2e-5 * rng.standard_normal((K, P.m2)))
ctx = T.default_context()
dvs = T.sdDualVertexSet(m2=P.m2)
ins, _ = dvs.push_many(pool)
assert ins.sum() > 400, ins.sum()          # distinct under the dedup rule (16 significant binary digits)
coef = T.sdSubprobCoefficients.from_tables(P.rbar, P.T_colptr, P.T_rowval, P.T_nzval, P.pos_row, P.pos_col)
epi = T.sdEpigraph(coef, 1.0, 0.0, dvs)
epi.add_scenarios(vals, None)
x = np.ones(P.n1)
out = {}
for mode in (0, 2):
    ctx.set_screen(mode)
    cuts, val = epi.build_cuts2(x, 0.5 * x, with_val=True)
    mv, mi = epi.argmax(x)
    out[mode] = [np.array([cuts[0].alpha, cuts[1].alpha]), np.stack([cuts[0].beta, cuts[1].beta]), np.asarray(val), mv, mi]
for u, v in zip(out[0], out[2]):
    assert np.array_equal(np.asarray(u).view(np.uint8), np.asarray(v).view(np.uint8))
st = epi.screen_stats()
assert st["fallbacks"] >= 1 and st["overflowed_lists"] > N // 16 + 16, st
print("CROWD OK", st)
"""


def test_crowd_overflows_and_falls_back():
    """A crowd of vertices 2e-5 apart: with uncentred operands every one of them is within the bound of the best,
    the lists (64 entries) overflow for every scenario, and the FP64 sweep queued behind the pass -- gated on the
    pass's control block, no host round trip -- must run and give the answer."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", CROWD_SCRIPT % {"root": root}], env=dict(os.environ, SQLP_CENTRE="0"),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "CROWD OK" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])


SWITCH_SCRIPT = """
import sys
import numpy as np
sys.path.insert(0, %(root)r)
from sqlp_b200 import twosd as T
from tests.helpers import load_instance, load_pool, sampled_values_at
P, z = load_instance("storm")
pool = load_pool("storm", 16384)[:3000]
ctx = T.default_context()
dvs = T.sdDualVertexSet(m2=P.m2)
dvs.push_many(pool[:2500])
coef = T.sdSubprobCoefficients.from_tables(P.rbar, P.T_colptr, P.T_rowval, P.T_nzval, P.pos_row, P.pos_col)
epi = T.sdEpigraph(coef, 1.0, 0.0, dvs)
vals = sampled_values_at(z, 4, np.arange(2600))
epi.add_scenarios(vals[:2500], 0.5 + np.arange(2500) %% 4)
for it in range(4):
    dvs.push_many(pool[2500 + 100 * it: 2600 + 100 * it])
    epi.add_scenarios(vals[2500 + 25 * it: 2525 + 25 * it], None)
    x = z["x_ev"] + 0.3 * it * (z["x_alt"] - z["x_ev"])
    out = {}
    for mode in (0, 2):
        ctx.set_screen(mode)
        cuts, val = epi.build_cuts2(x, z["x_alt"], with_val=True)
        mv, mi = epi.argmax(x)
        out[mode] = [np.array([cuts[0].alpha, cuts[1].alpha]), np.stack([cuts[0].beta, cuts[1].beta]), np.asarray(val), mv, mi]
    for u, v in zip(out[0], out[2]):
        assert np.array_equal(np.asarray(u).view(np.uint8), np.asarray(v).view(np.uint8)), it
st = epi.screen_stats()
assert st["passes"] >= 8 and st["overflowed_lists"] == 0 and st["bad_operands"] == 0, st
print("SWITCH OK", st)
"""


@pytest.mark.parametrize("env", [{"SQLP_RESOLVE": "lanes"}, {"SQLP_RESOLVE": "dmma"}, {"SQLP_SEED": "0"},
                                 {"SQLP_CENTRE": "0"}, {"SQLP_CENTRE": "0", "SQLP_SEED": "0", "SQLP_RESOLVE": "lanes"},
                                 {"SQLP_TWINS": "0"}])
def test_every_variant_of_the_pass_is_bit_identical(env):
    """The switches that select an older form of a step of the pass (a lane per candidate row, DMMA chains, cold scan,
    uncentred operands, every vertex a column) are kept for measurements; each must give the FP64 sweep's bits too."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", SWITCH_SCRIPT % {"root": root}], env=dict(os.environ, **env),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SWITCH OK" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])


# ---- score-equivalent vertices ("twins", csrc/kernels_pool.cuh) ---------------------------------------------

def test_twins_leave_results_bit_identical(T, monkeypatch):
    """storm's real duals: 16 384 vertices, far fewer classes on the relevant rows.  The sweep over one column per
    class must give the argmax (pool slots!), the values and the cuts of the sweep over every vertex, bit for bit --
    with the pool pushed at once or a few vertices per iteration, FP64 sweep or screening pass."""
    P, z = load_instance("storm")
    pool = load_pool("storm", 16384)[:6000]
    N = 3000
    vals = sampled_values_at(z, 3, np.arange(N))
    w = 0.5 + (np.arange(N) % 5) / 5.0
    xs = (z["x_ev"], z["x_alt"])

    def run(twins, incremental, screen):
        monkeypatch.setenv("SQLP_TWINS", "1" if twins else "0")
        ctx = T.Context(0)
        ctx.set_screen(screen)
        dvs = T.sdDualVertexSet(ctx=ctx, m2=P.m2)
        epi = T.sdEpigraph(coef_of(T, P), 1.0, 0.0, dvs)
        epi.add_scenarios(vals, w)
        if incremental:
            dvs.push_many(pool[:900])
            for a in range(900, len(pool), 700):               # crosses capacity doublings of the pool
                dvs.push_many(pool[a:a + 700])
                epi.argmax(xs[0])
        else:
            dvs.push_many(pool)
        out = [epi.argmax(x) for x in xs]
        (c0, c1), val = epi.build_cuts2(*xs, with_val=True)
        cols = epi.view_columns()
        res = (out, np.array([c0.alpha, c1.alpha]), np.stack([c0.beta, c1.beta]), val, cols)
        epi.close(); dvs.close(); ctx.close()
        return res

    ref = run(False, False, 0)
    assert ref[4] == (len(pool), P.m2) or ref[4][0] == len(pool)
    for twins, inc, screen in ((True, False, 0), (True, True, 0), (True, False, 2), (True, True, 2)):
        got = run(twins, inc, screen)
        assert got[4][0] < len(pool) // 2 and got[4][1] < P.m2, got[4]       # classes, relevant rows
        for (rv, ri), (gv, gi) in zip(ref[0], got[0]):
            assert np.array_equal(ri, gi) and np.array_equal(rv.view(np.uint64), gv.view(np.uint64))
        for a, b in zip(ref[1:4], got[1:4]):
            assert np.array_equal(np.asarray(a).view(np.uint64), np.asarray(b).view(np.uint64))
    print("storm K=6000 real:", got[4][0], "classes on", got[4][1], "relevant rows of", P.m2)


def test_twins_keep_nonfinite_vertices_apart(T, monkeypatch):
    """0 * Inf is NaN: a vertex with a non-finite entry on an irrelevant row is neither a representative nor
    shadowed."""
    P = synthetic_problem(m2=64, n1=10, s=24)
    N, K = 700, 300
    vals = synthetic_values(P, N, seed=2)
    pool = synthetic_pool(P.m2, K, seed=8)
    rel = np.zeros(P.m2, bool)
    rel[P.pos_row] = True; rel[P.rbar != 0] = True; rel[P.T_rowval] = True
    free = np.nonzero(~rel)[0]
    assert len(free) >= 2
    for k in range(50, 150):                     # twins of vertices 0..99: differ on an irrelevant row only
        pool[k] = pool[k - 50]
        pool[k, free[0]] += 1.0 + k
    pool[10, free[1]] = np.inf                   # the representative of 60 is not finite: 60 must stay in the sweep
    pool[70, free[1]] = np.nan                   # a shadowed vertex that is not finite: kept apart, never wins
    x = np.linspace(0.5, 2.0, P.n1)

    def run(twins):
        monkeypatch.setenv("SQLP_TWINS", "1" if twins else "0")
        ctx = T.Context(0)
        dvs = T.sdDualVertexSet(ctx=ctx, m2=P.m2)
        dvs.push_many(pool)
        epi = T.sdEpigraph(coef_of(T, P), 1.0, 0.0, dvs)
        epi.add_scenarios(vals, None)
        mv, mi = epi.argmax(x)
        cut = epi.build_cut(x)
        cols = epi.view_columns()[0]
        epi.close(); dvs.close(); ctx.close()
        return mv, mi, cut, cols
    a, b = run(False), run(True)
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[0].view(np.uint64), b[0].view(np.uint64))
    assert a[2].alpha == b[2].alpha and np.array_equal(a[2].beta, b[2].beta)
    assert a[3] == K and b[3] == K - 100 + 2, (a[3], b[3])      # 100 twins, 2 of them kept apart
