"""Row N4 on the device: an epigraph built by the library straight from SMPS files
(``sqlp_smps_load`` + ``sqlp_epi_create_smps``) samples the same scenarios and forms the same cuts, bit for
bit, as one built from host-extracted tables (``sqlp_epi_create`` + ``set_outcomes`` / ``set_distributions``),
and both match the oracle fed the host twin of the sampler.  Files are written on the fly (dT elements
included -- no shipped instance has them), so the test needs nothing outside the repo."""
import numpy as np
import pytest

from oracle import oracle as O
from sqlp_b200 import smps
from tests.helpers import write_smps, balanced_pool
from tests.test_gpu_parity import check_cut

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("continuous,n_T", [(False, 0), (False, 3), (True, 2)])
def test_epigraph_from_smps_files(tmp_path, continuous, n_T):
    from sqlp_b200 import twosd as T
    paths, ex = write_smps(str(tmp_path), seed=21 + n_T, continuous=continuous, n1=9, n2=12, m1=3, m2=30,
                           n_rhs_elems=11, n_T_elems=n_T)
    native = smps.NativeSmps(paths["cor"], paths["tim"], paths["sto"])
    st = native.stage2()
    pr, pc, kind, a, b, cnt, vals, probs = native.elements()
    cdf = np.cumsum(probs, axis=1)
    P = O.Problem(st.m2, st.n1, st.rbar, st.T_colptr, st.T_rowval, st.T_nzval, st.pos_row, st.pos_col)
    N, seed = 900, 77
    xs = (3.0 * O.u01(3, np.arange(st.n1)), 3.0 * O.u01(5, np.arange(st.n1)))
    twin = O.sample_twin(seed, 0, N, kind, a, b, vals, cdf, np.maximum(cnt, 1))
    pool = balanced_pool(P, xs[0], twin.mean(axis=0), 300)

    dvs_a = T.sdDualVertexSet(m2=st.m2)
    dvs_a.push_many(pool)
    epi_a = T.sdEpigraph.from_smps(native, 1.0, 0.0, dvs_a)
    assert epi_a.subproblem_coef.position_table == ex["positions"]

    dvs_b = T.sdDualVertexSet(m2=st.m2)
    dvs_b.push_many(pool)
    coef = T.sdSubprobCoefficients(st.rbar, st.T_colptr, st.T_rowval, st.T_nzval, st.n1,
                                   {r: i for i, r in enumerate(st.row_names)},
                                   {c: j for j, c in enumerate(st.x_names)}, ex["positions"])
    epi_b = T.sdEpigraph(coef, 1.0, 0.0, dvs_b)
    epi_b.set_outcomes(vals, cdf, np.maximum(cnt, 1))
    if continuous:
        epi_b.set_distributions(kind, a, b)

    cuts = []
    for epi in (epi_a, epi_b):
        epi.sample_scenarios(N, seed=seed)
        assert epi.counts()[0] == N
        cuts.append(epi.build_cuts2(*xs, with_val=True))
    ((a0, a1), va), ((b0, b1), vb) = cuts
    for ca, cb in ((a0, b0), (a1, b1)):                     # same tables -> same bits
        assert ca.alpha == cb.alpha and np.array_equal(ca.beta, cb.beta) and ca.weight_mark == cb.weight_mark
    assert np.array_equal(va, vb)
    for i in (0, 1, 127, 128, N - 1):
        da, db = epi_a.delta(i), epi_b.delta(i)
        assert np.array_equal(da.delta_rhs, db.delta_rhs) and da.delta_transfer == db.delta_transfer

    if not continuous:                                      # discrete draws: the twin's deltas exactly
        for i in (0, 5, N - 1):
            drhs, _ = O.delta_coefficients(P, twin[i])
            assert np.array_equal(epi_a.delta(i).delta_rhs, drhs)
    for x, cut in zip(xs, (a0, a1)):
        check_cut(O, P, twin, np.ones(N), x, pool, cut, epi=epi_a)

    # named scenarios resolve through the names the reader kept (add_scenario!, epigraph.jl:81-96)
    scen = [(pos, float(twin[0, e])) for e, pos in enumerate(ex["positions"])][::-1]
    T.add_scenario_(epi_a, scen, 2.0)
    assert epi_a.counts()[0] == N + 1
    drhs, _ = O.delta_coefficients(P, twin[0])
    assert np.array_equal(epi_a.delta(N).delta_rhs, drhs)


def test_sd_run_from_smps_text_equals_the_run_from_tables(tmp_path):
    """lands written as .cor/.tim/.sto text, read by the native reader, the epigraph built by
    ``sqlp_epi_create_smps`` and the whole SD loop run in fused mode (one ``sqlp_cell_sd_step`` per
    iteration): every cut equals, bit for bit, the cut of the same loop over an epigraph built from the
    fixture's tables with separate calls."""
    from sqlp_b200 import sd, twosd as T
    from tests.helpers import load_full_instance, make_cell, sample_instance_values, write_smps_from_full
    zf = load_full_instance("lands")
    prefix = write_smps_from_full(zf, str(tmp_path), "lands")
    native = smps.NativeSmps(prefix + ".cor", prefix + ".tim", prefix + ".sto")
    ft = smps.full_tables(native.cor(), native.stage2(), native.sto())
    cells, logs = [], ([], [])
    for from_files in (True, False):
        dvs = T.sdDualVertexSet(m2=int(zf["m2"]))
        if from_files:
            mk = lambda w, lb: T.sdEpigraph.from_smps(native, w, lb, dvs)
        else:
            coef = T.sdSubprobCoefficients.from_tables(zf["rbar"], zf["T_colptr"], zf["T_rowval"], zf["T_nzval"],
                                                       zf["pos_row"], zf["pos_col"])
            mk = lambda w, lb: T.sdEpigraph(coef, w, lb, dvs)
        cell, lp = make_cell(ft if from_files else zf, dvs, mk, np.full(4, 3.0), n_epi=2)
        cell.fused_step = from_files
        cells.append((cell, lp))
    vals = sample_instance_values(zf, 2 * 60, seed=5)
    for it in range(60):
        scen = [vals[2 * it], vals[2 * it + 1]]
        for (cell, lp), log in zip(cells, logs):
            sd.sd_iteration_(cell, scen, lambda i, x, v: lp.solve(x, v),
                             on_cuts=lambda i, cand, inc: log.append((cand.alpha, cand.beta.copy(), inc.alpha, inc.beta.copy())))
    assert len(logs[0]) == len(logs[1]) == 120
    for a, b in zip(*logs):
        assert a[0] == b[0] and a[2] == b[2] and np.array_equal(a[1], b[1]) and np.array_equal(a[3], b[3])
    assert np.array_equal(cells[0][0].x_incumbent, cells[1][0].x_incumbent)
    assert len(cells[0][0].dual_vertices) == len(cells[1][0].dual_vertices) > 3
