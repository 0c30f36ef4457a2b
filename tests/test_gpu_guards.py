"""Memory-safety check without a sanitizer.

``compute-sanitizer`` is closed on the GPU pool this library is developed on, so an out-of-bounds write has to be
caught by the library itself: with ``SQLP_GUARD=1`` every device buffer carries 512 bytes of a known pattern in
front of it and behind it, and ``sqlp_guard_check`` counts the zones that were written to.  Here the smallest and
the most ragged cases of every kernel of the path run under guards -- dedup pushes across capacity doublings,
``add_scenario!`` from values and sampled, with and without random T entries, the three FP64 sweeps, the
screening pass (tcgen05) with K-ranges, classes of score-equivalent vertices, both cut reductions, the cut list --
each in its own process (the switch is read once), and every guard zone must still hold its pattern.
"""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent('''
    import ctypes as C, sys
    import numpy as np
    sys.path.insert(0, %(root)r)
    from sqlp_b200 import twosd as T, _lib
    from tests.helpers import (load_instance, load_pool, sampled_values_at, synthetic_pool, synthetic_problem,
                               synthetic_values)

    def coef_of(P):
        return T.sdSubprobCoefficients.from_tables(P.rbar, P.T_colptr, P.T_rowval, P.T_nzval, P.pos_row, P.pos_col)

    seen = [0, 0]
    def check_guards():                 # before the buffers of a case are released
        n, bad = C.c_int64(), C.c_int64()
        _lib.check(_lib.lib().sqlp_guard_check(C.byref(n), C.byref(bad)))
        seen[0] += n.value
        seen[1] += bad.value

    ctx = T.default_context()
    for screen in (0, 2):
        ctx.set_screen(screen)
        for (N, K, s, n_T) in ((1, 1, 1, 0), (129, 127, 13, 0), (700, 1500, 117, 0), (300, 260, 24, 3), (5, 2100, 40, 0),
                               (2300, 300, 86, 0)):
            P = synthetic_problem(m2=max(64, s + 11), n1=16, s=s, n_T=n_T, first_stoch_row=3)
            vals = synthetic_values(P, N, seed=3)
            pool = synthetic_pool(P.m2, K, seed=5, scale=700.0)
            pool[K // 2:, P.m2 - 1] += 1.0            # (some vertices differ on an irrelevant row only)
            dvs = T.sdDualVertexSet(m2=P.m2)
            for a in range(0, K, 400):                 # pushes across capacity doublings, duplicates included
                dvs.push_many(np.vstack([pool[a:a + 400], pool[a:a + 3]]))
            epi = T.sdEpigraph(coef_of(P), 1.0, 0.0, dvs)
            epi.add_scenarios(vals[: N // 2 + 1], None)
            epi.add_scenarios(vals[N // 2 + 1:], 0.5 + np.arange(N - N // 2 - 1) %% 3)
            x = 5.0 * np.cos(np.arange(P.n1))
            epi.argmax(x)
            epi.build_cut(x)
            cuts = epi.build_cuts2(x, 2.0 * x + 1.0)
            epi.cuts_commit(True)
            epi.cuts_delete([0])
            epi.evaluate(x)
            epi.master_rows()
            epi.eval_dual(0, 0, x)
            epi.delta(0)
            T.sd_step([epi], [vals[0]], [1.0], np.vstack([pool[0] * 1.5, pool[K - 1]]), x, x + 1.0)
            check_guards()
            epi.close(); dvs.close()
    # real instance: sampled scenarios, classes, both reductions at a size where the weight-sum path runs
    P, z = load_instance("storm")
    pool = load_pool("storm", 3000)
    dvs = T.sdDualVertexSet(m2=P.m2)
    dvs.push_many(pool)
    epi = T.sdEpigraph(coef_of(P), 1.0, 0.0, dvs)
    epi.set_outcomes(z["out_vals"], z["out_cdf"], z["out_cnt"])
    epi.sample_scenarios(20001, seed=1, weight_seed=4)
    for screen in (0, 2):
        ctx.set_screen(screen)
        epi.build_cuts2(z["x_ev"], z["x_alt"])
    check_guards()
    print("GUARDS", seen[0], seen[1])
''')


@pytest.mark.parametrize("env", [{}, {"SQLP_CONTRACT": "stream"}, {"SQLP_CONTRACT": "resident", "SQLP_CONTRACT_GRID": "7"},
                                 {"SQLP_REDUCE": "2", "SQLP_RESOLVE": "dmma"}, {"SQLP_TWINS": "0", "SQLP_REDUCE": "1"}, {"SQLP_RESOLVE": "lanes", "SQLP_SEED": "0"}])
def test_no_guard_zone_is_written(env):
    e = dict(os.environ, SQLP_GUARD="1", **env)
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("GUARDS")][-1]
    n, bad = (int(v) for v in line.split()[1:])
    assert n >= 300, line           # the guards were really on
    assert bad == 0, line
