"""The epigraph's cut list on the device against the reference's
known answers (test/sd_test.jl:150-194) and the oracle's restatement of evaluate_epigraph,
sync_cuts! / add_cut_to_master! and check_improvement."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import load_instance, synthetic_pool, synthetic_problem, synthetic_values

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    from sqlp_b200 import twosd
    return twosd


def lands_epi(T, weight, lb, n_scen=2):
    P, z = load_instance("lands")
    coef = T.sdSubprobCoefficients.from_tables(P.rbar, P.T_colptr, P.T_rowval, P.T_nzval, P.pos_row, P.pos_col)
    dvs = T.sdDualVertexSet(m2=P.m2)
    epi = T.sdEpigraph(coef, weight, lb, dvs)
    epi.add_scenarios(np.array([[5.0], [3.0]][:n_scen]))      # two scenarios of weight 1 (sd_test.jl:141-147)
    return epi


def test_reference_known_answers(T):                 # sd_test.jl:166-194
    e1, e2 = lands_epi(T, 0.5, 0.0), lands_epi(T, 0.5, 100.0)
    cut1 = T.sdCut(1.0, np.array([2.0, 3, 4, 5]), 1.0)
    cut2 = T.sdCut(6.0, np.array([7.0, 8, 9, 10]), 2.0)
    inc = T.sdCut(11.0, np.array([12.0, 13, 14, 15]), 1.0)
    e1.cuts_push(cut1); e1.cuts_push(cut2); e1.cuts_set_incumbent(inc)
    e2.cuts_push(cut1)
    assert e1.cuts_count() == (2, True) and e2.cuts_count() == (1, False)
    r1, r2 = e1.master_rows(), e2.master_rows()
    assert r1.shape == (3, 5) and (r1[2] == [11, 12, 13, 14, 15]).all()
    assert r2[0, 0] == 50.5 and (r2[0, 1:] == [1.0, 1.5, 2.0, 2.5]).all()     # 100 * 0.5 + 1.0 * 0.5
    x10 = np.full(4, 10.0)
    assert e1.evaluate(x10) == 551.0 * 0.5
    assert e2.evaluate(x10) == (141 / 2 + 100 / 2) * 0.5
    assert e2.evaluate(-np.ones(4)) == 100.0 * 0.5
    got = e1.cuts_get(1)
    assert got.alpha == 6.0 and (got.beta == cut2.beta).all() and got.weight_mark == 2.0
    assert e1.cuts_get(-1).alpha == 11.0
    e1.cuts_set_incumbent(None)
    assert e1.cuts_count() == (2, False) and e1.evaluate(x10) == 346.0 * 0.5
    e1.cuts_delete([0])
    assert e1.cuts_count() == (1, False) and e1.cuts_get(0).alpha == 6.0


def test_random_lists_against_the_oracle(T):
    P = synthetic_problem(m2=60, n1=37, s=10)
    coef = T.sdSubprobCoefficients.from_tables(P.rbar, P.T_colptr, P.T_rowval, P.T_nzval, P.pos_row, P.pos_col)
    dvs = T.sdDualVertexSet(m2=P.m2)
    lb, w = -250.0, 0.35
    epi = T.sdEpigraph(coef, w, lb, dvs)
    vals = synthetic_values(P, 40)
    epi.add_scenarios(vals, 0.5 + O.u01(4, np.arange(40)))
    tw = epi.total_scenario_weight
    rng = np.random.default_rng(5)
    cuts = [(float(rng.normal(0, 300)), rng.normal(0, 20, P.n1), float(rng.uniform(1, tw))) for _ in range(70)]
    inc = (float(rng.normal(0, 300)), rng.normal(0, 20, P.n1), tw)
    for a, b, m in cuts:
        epi.cuts_push(T.sdCut(a, b, m))
    epi.cuts_set_incumbent(T.sdCut(*inc))
    assert np.array_equal(epi.master_rows(), O.cut_master_rows(cuts, inc, tw, lb))       # no fused arithmetic
    for seed in range(5):
        x = rng.normal(0, 3, P.n1)
        ref = O.cut_evaluate(cuts, inc, x, tw, lb, w)
        assert abs(epi.evaluate(x) - ref) <= 1e-12 * max(1.0, abs(ref))
    drop = [0, 3, 4, 69]
    epi.cuts_delete(drop)
    kept = [c for j, c in enumerate(cuts) if j not in drop]
    assert epi.cuts_count() == (66, True)
    assert np.array_equal(epi.master_rows(), O.cut_master_rows(kept, inc, tw, lb))
    with pytest.raises(T.SqlpError):
        epi.cuts_delete([66])
    with pytest.raises(T.SqlpError):
        epi.cuts_commit()                      # nothing formed since the last commit


def test_commit_snapshot_and_improvement(T):
    """Two iterations' worth of cut formation on storm: the committed cuts are the bits build_cuts2
    returned, the snapshot is the list before the commit, and the incumbent test matches the oracle."""
    P, z = load_instance("storm")
    coef = T.sdSubprobCoefficients.from_tables(P.rbar, P.T_colptr, P.T_rowval, P.T_nzval, P.pos_row, P.pos_col)
    dvs = T.sdDualVertexSet(m2=P.m2)
    dvs.push_many(z["pool"])
    E = 2
    epis = [T.sdEpigraph(coef, 0.5, -1e7 * (e + 1), dvs) for e in range(E)]
    from tests.helpers import sample_instance_values
    xc, xi = z["x_ev"], z["x_alt"]
    cost = 1.0 + O.u01(9, np.arange(P.n1))
    state = [dict(cuts=[], inc=None) for _ in range(E)]
    for it in range(3):
        last = []
        for e, epi in enumerate(epis):
            epi.add_scenarios(sample_instance_values(z, 50, seed=10 * it + e))
            last.append((list(state[e]["cuts"]), state[e]["inc"], epi.total_scenario_weight, epi.lower_bound,
                         epi.objective_weight))
        T.build_cuts_at_candidate_and_incumbent(epis, xc, xi)
        cur = []
        for e, epi in enumerate(epis):
            epi.cuts_commit(True)
            cand, inc = epi.cuts[-1], epi.incumbent_cut
            state[e]["cuts"].append((cand.alpha, cand.beta, cand.weight_mark))
            state[e]["inc"] = (inc.alpha, inc.beta, inc.weight_mark)
            got = epi.cuts_get(epi.cuts_count()[0] - 1)
            assert got.alpha == cand.alpha and np.array_equal(got.beta, cand.beta) and got.weight_mark == cand.weight_mark
            gi = epi.cuts_get(-1)
            assert gi.alpha == inc.alpha and np.array_equal(gi.beta, inc.beta)
            cur.append((state[e]["cuts"], state[e]["inc"], epi.total_scenario_weight, epi.lower_bound,
                        epi.objective_weight))
            for x in (xc, xi):
                ref_last = O.cut_evaluate(*last[e][:2], x, *last[e][2:])
                assert abs(epi.evaluate(x, snapshot=True) - ref_last) <= 1e-12 * max(1.0, abs(ref_last))
        ref = O.cut_check_improvement(last, cur, xc, xi, cost)
        got = T.check_improvement_device(epis, xc, xi, cost)
        for a, b in zip(got[:3], ref[:3]):
            assert abs(a - b) <= 1e-12 * max(1.0, abs(b))
        assert got[3] == ref[3]


def test_sd_step_equals_the_separate_calls_bit_for_bit():
    """``sqlp_cell_sd_step`` (add_scenario! per epigraph, the iteration's pushes, both cuts of every epigraph;
    one synchronisation) against the same iteration made of ``add_scenario_``, ``push`` and
    ``build_cuts_at_candidate_and_incumbent``: identical dedup decisions, slots, cut bits and pool contents
    over several iterations with two weighted epigraphs, new and duplicate vertices, and dT elements."""
    from sqlp_b200 import twosd as T
    from tests.helpers import synthetic_problem, synthetic_values, synthetic_pool
    P = synthetic_problem(m2=60, n1=14, s=22, n_T=8)
    coef = lambda: T.sdSubprobCoefficients.from_tables(P.rbar, P.T_colptr, P.T_rowval, P.T_nzval, P.pos_row, P.pos_col)
    pool = synthetic_pool(P.m2, 150)
    vals = synthetic_values(P, 400)
    E, iters = 2, 6
    cells = []
    for _ in range(2):
        dvs = T.sdDualVertexSet(m2=P.m2)
        dvs.push_many(pool[:100])
        epis = [T.sdEpigraph(coef(), 0.5, 0.0, dvs) for _ in range(E)]
        for e, epi in enumerate(epis):
            epi.add_scenarios(vals[e:300:E], 0.5 + O.u01(4, np.arange(len(vals[e:300:E]))))
        cells.append((dvs, epis))
    (dvs_a, epis_a), (dvs_b, epis_b) = cells
    for it in range(iters):
        xc, xi = 10.0 * O.u01(30 + it, np.arange(P.n1)), 10.0 * O.u01(50 + it, np.arange(P.n1))
        scen = [vals[300 + it * E + e] for e in range(E)]
        w = np.array([1.0 + 0.25 * e + 0.5 * it for e in range(E)])
        verts = np.stack([pool[100 + 2 * it], pool[7 * it], pool[100 + 2 * it + 1], pool[100 + 2 * it]])  # new, dup, new, dup of new
        ins_a, idx_a, cuts_a = T.sd_step(epis_a, scen, w, verts, xc, xi)
        for e in range(E):
            T.add_scenario_(epis_b[e], scen[e], float(w[e]))
        pushed = [dvs_b.push(v) for v in verts]
        cuts_b = T.build_cuts_at_candidate_and_incumbent(epis_b, xc, xi)
        assert list(ins_a) == [p[0] for p in pushed] == [True, False, True, False]
        assert list(idx_a) == [p[1] for p in pushed]
        assert len(dvs_a) == len(dvs_b) == 100 + 2 * (it + 1)
        for (ca, ia), (cb, ib) in zip(cuts_a, cuts_b):
            for u, v in ((ca, cb), (ia, ib)):
                assert u.alpha == v.alpha and np.array_equal(u.beta, v.beta) and u.weight_mark == v.weight_mark
        for e in range(E):
            assert epis_a[e].counts() == epis_b[e].counts()
    assert np.array_equal(np.stack(list(dvs_a)), np.stack(list(dvs_b)))
    # and against the oracle on the final state of one epigraph
    n0 = len(vals[0:300:E])
    all_vals = np.vstack([vals[0:300:E]] + [vals[300 + it * E][None, :] for it in range(iters)])
    all_w = np.concatenate([0.5 + O.u01(4, np.arange(n0)), [1.0 + 0.5 * it for it in range(iters)]])
    ref = O.build_sasa_cut(P, all_vals, all_w, xc, np.stack(list(dvs_a)), forced_idx=epis_a[0].argmax(xc)[1])
    cut = cuts_a[0][0]
    assert abs(cut.alpha - ref["alpha"]) <= 1e-10 * (abs(ref["alpha"]) + 1.0)
    assert np.max(np.abs(cut.beta - ref["beta"])) <= 1e-10 * (np.abs(ref["beta"]).max() + 1.0)
    assert cut.weight_mark == ref["weight_mark"]
