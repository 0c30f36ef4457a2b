"""Scenario-sharded path on 2 GPUs (skipped on a single-GPU box): NCCL vertex broadcast and
partial all-gather inside the library, host plumbing through torch.distributed (gloo)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from sqlp_b200 import dist as D, twosd as T
        from tests.helpers import check_argmax_parity, load_instance, sample_instance_values
        ctx = D.init_context()
        assert (ctx.rank, ctx.world) == (rank, world)
        P, z = load_instance("storm")
        N = 1500
        vals = sample_instance_values(z, N)
        w = 0.5 + O.u01(4, np.arange(N))
        pool = z["pool"]
        coef = T.sdSubprobCoefficients.from_tables(P.rbar, P.T_colptr, P.T_rowval, P.T_nzval,
                                                   P.pos_row, P.pos_col)
        dvs = T.sdDualVertexSet(ctx=ctx, m2=P.m2)
        # only rank 0's vectors count: the others push garbage of the right shape
        src = pool if rank == 0 else np.full_like(pool, 7.0)
        ins, idx = dvs.push_many(np.vstack([src, src[:3]]))
        assert ins.sum() == len(pool) and len(dvs) == len(pool)
        assert np.array_equal(np.stack(list(dvs)), pool)
        epi = T.sdEpigraph(coef, 1.0, 0.0, dvs)
        epi.add_scenarios(vals, w)                     # SPMD: same arrays on every rank
        ng, nl, tw = epi.counts()
        assert ng == N and nl == D.local_count(N, rank, world)
        xs = (z["x_ev"], z["x_alt"])
        (cand, inc), val = epi.build_cuts2(*xs, with_val=True)
        ok = True
        for x, cut in zip(xs, (cand, inc)):
            mv, mi = epi.argmax(x)                      # this rank's scenarios, local order
            gmv = D.gather_scenario_results(mv, N)
            gmi = D.gather_scenario_results(mi, N)
            check_argmax_parity(P, vals, x, pool, gmv, gmi)
            ref = O.build_sasa_cut(P, vals, w, x, pool, forced_idx=gmi)
            ok &= abs(cut.alpha - ref["alpha"]) <= 1e-10 * abs(ref["alpha"])
            ok &= bool(np.allclose(cut.beta, ref["beta"], rtol=1e-10, atol=1e-6))
            ok &= cut.weight_mark == ref["weight_mark"]
        # sharded device sampling generates the same scenarios as one GPU would
        epi2 = T.sdEpigraph(coef, 1.0, 0.0, dvs)
        epi2.set_outcomes(z["out_vals"], z["out_cdf"], z["out_cnt"])
        epi2.sample_scenarios(N, seed=1, weight_seed=4)
        c2 = epi2.build_cut(xs[1])
        ok &= abs(c2.alpha - inc.alpha) <= 1e-12 * abs(inc.alpha)
        ok &= bool(np.allclose(c2.beta, inc.beta, rtol=1e-12, atol=1e-9))
        again = epi.build_cuts2(*xs)[1]                 # run-to-run bitwise identical
        ok &= again.alpha == inc.alpha and np.array_equal(again.beta, inc.beta)
        # the sharded cut equals the cut ONE GPU forms from the same scenarios, within 1e-10 of each
        # coefficient's absolute-sum scale (the two differ in the order of the weighted sum only)
        one = T.Context(ctx.device)
        dvs1 = T.sdDualVertexSet(ctx=one, m2=P.m2)
        dvs1.push_many(pool)
        epi1 = T.sdEpigraph(coef, 1.0, 0.0, dvs1)
        epi1.add_scenarios(vals, w)
        (cand1, inc1), _ = epi1.build_cuts2(*xs, with_val=True)
        for a, b, x in ((cand, cand1, xs[0]), (inc, inc1, xs[1])):
            sel = pool[D.gather_scenario_results(epi.argmax(x)[1], N)]
            p_i = w / w.sum()
            sa = np.sum(p_i * np.abs(sel @ P.rbar)) + 1e-300
            sb = (p_i[:, None] * np.abs(sel @ P.T_dense())).sum(axis=0) + 1e-300
            ok &= abs(a.alpha - b.alpha) <= 1e-10 * sa
            ok &= bool((np.abs(a.beta - b.beta) <= 1e-10 * sb).all())
        epi1.close(); dvs1.close()
        # a whole cell in one call: ONE all-gather for the E epigraphs (rows + their "no argmax" words)
        E = 3
        epis = [T.sdEpigraph(coef, 1.0 / E, 0.0, dvs) for _ in range(E)]
        for e in range(E):
            epis[e].add_scenarios(vals[e::E], w[e::E])
        outc = T.build_cuts_at_candidate_and_incumbent(epis, *xs)
        for e in range(E):
            for xi, x in enumerate(xs):
                gmi = D.gather_scenario_results(epis[e].argmax(x)[1], len(vals[e::E]))
                ref = O.build_sasa_cut(P, vals[e::E], w[e::E], x, pool, forced_idx=gmi)
                ok &= abs(outc[e][xi].alpha - ref["alpha"]) <= 1e-10 * abs(ref["alpha"])
                ok &= bool(np.allclose(outc[e][xi].beta, ref["beta"], rtol=1e-10, atol=1e-6))
        # a scenario without argmax on ONE rank only: every rank must raise together (the flag word travels with
        # the gathered row) and the communicator must stay usable afterwards
        Td = P.T_dense()
        base = [P.rbar - Td @ x for x in xs]
        el = next(e for e in range(P.s) if base[0][P.pos_row[e]] > 0 and base[1][P.pos_row[e]] > 0)
        j = int(P.pos_row[el])
        bad_pool = np.zeros((1, P.m2)); bad_pool[0, j] = np.inf      # bias = +Inf at both points
        dvs_b = T.sdDualVertexSet(ctx=ctx, m2=P.m2)
        dvs_b.push_many(bad_pool)
        epi_b = T.sdEpigraph(coef, 1.0, 0.0, dvs_b)
        vb = np.tile(P.rbar[P.pos_row], (300, 1))
        vb[:, el] += 1.0                                 # d_j > 0: score +Inf, a (degenerate) winner
        vb[130, el] -= 2.0                               # scenario 130 (second 128-block, rank 1): +Inf - Inf = NaN, no winner
        epi_b.add_scenarios(vb, None)
        raised = False
        try:
            epi_b.build_cuts2(*xs)
        except T.NoArgmaxError:
            raised = True
        ok &= raised
        again2 = epi.build_cuts2(*xs)[1]                # the next collective still lines up
        ok &= again2.alpha == inc.alpha and np.array_equal(again2.beta, inc.beta)
        q.put((rank, bool(ok), np.concatenate([[cand.alpha, inc.alpha], cand.beta, inc.beta]).tobytes()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_cuts(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert all(r[2] == res[0][2] for r in res)        # fixed rank-order sum: identical bits on every rank
    # ONE process driving the same GPUs (sqlp_ctx_create_multi) forms the same cuts, bit for bit
    assert _one_process_cuts(world) == res[0][2]


def _one_process_cuts(world):
    from oracle import oracle as O
    from sqlp_b200 import twosd as T
    from tests.helpers import check_argmax_parity, load_instance, sample_instance_values
    ctx = T.Context(devices=range(world))
    P, z = load_instance("storm")
    N = 1500
    vals = sample_instance_values(z, N)
    w = 0.5 + O.u01(4, np.arange(N))
    pool = z["pool"]
    coef = T.sdSubprobCoefficients.from_tables(P.rbar, P.T_colptr, P.T_rowval, P.T_nzval, P.pos_row, P.pos_col)
    dvs = T.sdDualVertexSet(ctx=ctx, m2=P.m2)
    ins, idx = dvs.push_many(np.vstack([pool, pool[:3]]))
    assert ins.sum() == len(pool) and len(dvs) == len(pool)
    epi = T.sdEpigraph(coef, 1.0, 0.0, dvs)
    epi.add_scenarios(vals, w)
    assert epi.counts()[:2] == (N, N)                  # one process: every scenario is "local"
    xs = (z["x_ev"], z["x_alt"])
    (cand, inc), val = epi.build_cuts2(*xs, with_val=True)
    for x, cut in zip(xs, (cand, inc)):
        mv, mi = epi.argmax(x)                          # GLOBAL scenario order
        check_argmax_parity(P, vals, x, pool, mv, mi)
        ref = O.build_sasa_cut(P, vals, w, x, pool, forced_idx=mi)
        assert abs(cut.alpha - ref["alpha"]) <= 1e-10 * abs(ref["alpha"])
        assert np.allclose(cut.beta, ref["beta"], rtol=1e-10, atol=1e-6)
    # the whole iteration in one call, as sd_iteration! would issue it
    E = 2
    epis = [T.sdEpigraph(coef, 0.5, 0.0, dvs) for _ in range(E)]
    for e in range(E):
        epis[e].add_scenarios(vals[e::E][:-1], w[e::E][:-1])
    new_v = pool[:2] * 1.5
    ins2, idx2, cuts = T.sd_step(epis, [vals[e::E][-1] for e in range(E)], [w[e::E][-1] for e in range(E)],
                                 np.vstack([new_v[0], pool[5], new_v[1], pool[7]]), *xs)
    assert list(ins2) == [True, False, True, False] and list(idx2) == [len(pool), 5, len(pool) + 1, 7]
    pool2 = np.vstack([pool, new_v])
    for e in range(E):
        for xi, x in enumerate(xs):
            mv, mi = epis[e].argmax(x)
            ref = O.build_sasa_cut(P, vals[e::E], w[e::E], x, pool2, forced_idx=mi)
            assert abs(cuts[e][xi].alpha - ref["alpha"]) <= 1e-10 * abs(ref["alpha"])
            assert np.allclose(cuts[e][xi].beta, ref["beta"], rtol=1e-10, atol=1e-6)
    out = np.concatenate([[cand.alpha, inc.alpha], cand.beta, inc.beta]).tobytes()
    for h in epis + [epi]:
        h.close()
    dvs.close()
    ctx.close()
    return out
