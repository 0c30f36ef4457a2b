"""The screening pass of DESIGN.md section 10, restated on the CPU (``oracle/screen.py``): bf16 x 2 operands,
three products, fp32 accumulation, a Cauchy-Schwarz error bound -- and the exact arithmetic deciding among the
survivors.  The claim under test: the survivors always contain the vertex the full FP64 sweep selects, so the
screened argmax IS the oracle's argmax (same index, same value), while only a handful of vertices per scenario
need the exact score.  CPU only; the kernel's own tests are in ``test_gpu_screen.py``."""
import numpy as np
import pytest

from oracle import screen as S
from tests.helpers import (load_instance, load_pool, sample_instance_values, synthetic_pool, synthetic_problem,
                           synthetic_values)


def test_bf16_split_error_budget():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(20000) * 10.0 ** rng.integers(-6, 7, 20000)
    hi, lo = S.split2(x)
    assert np.all(np.abs(x - hi) <= S.U_B * np.abs(x) * (1 + 2.0 ** -14))
    assert np.all(np.abs(x - hi - lo) <= 1.001 * S.U_B ** 2 * np.abs(x))
    for part in (hi, lo):                                  # both parts are bf16 numbers: 8 significant bits
        m, _ = np.frexp(part)
        assert np.all(m * 256 == np.round(m * 256))
    a, b = rng.standard_normal((50, 117)), rng.standard_normal((40, 117))
    err = np.abs(S.approx_dots(a, b) - b @ a.T)
    bound = S.eps(117) * np.outer(np.linalg.norm(b, axis=1), np.linalg.norm(a, axis=1))
    assert np.all(err <= bound) and err.max() > 1e-4 * bound.max()      # holds, and is not vacuous by 4 orders


@pytest.mark.parametrize("name,max_mean", [("baa99-20", 1.2), ("ssn", 1.2), ("storm", 2.0)])
def test_screened_argmax_is_the_oracle_argmax_on_real_instances(oracle, name, max_mean):
    P, z = load_instance(name)
    N = 400
    vals = sample_instance_values(z, N, seed=12)
    pool = z["pool"]
    for x in (z["x_ev"], z["x_alt"]):
        ov, oi = oracle.argmax_procedure(P, vals, x, pool)
        mv, mi, ncand = S.argmax_screened(P, vals, x, pool)
        assert np.array_equal(mi, oi) and np.array_equal(mv, ov)
        assert ncand.min() >= 1
        if x is z["x_alt"]:     # away from the harvest point (where real pools hold exact ties)
            assert ncand.mean() <= max_mean, ncand.mean()


def test_centred_operands_keep_the_argmax_and_shrink_the_lists_on_storms_real_pool(oracle):
    """Real LP duals of storm are a cloud around a common point: relative to its centre (and the scenarios relative
    to theirs) the error bound is ~10x smaller.  Same argmax either way; far fewer exact evaluations."""
    P, z = load_instance("storm")
    pool = load_pool("storm", 2048)
    vals = sample_instance_values(z, 200, seed=5)
    ov, oi = oracle.argmax_procedure(P, vals, z["x_alt"], pool)
    counts = {}
    for centre in (False, True):
        mv, mi, ncand = S.argmax_screened(P, vals, z["x_alt"], pool, centre=centre)
        assert np.array_equal(mi, oi) and np.array_equal(mv, ov)
        counts[centre] = ncand.mean()
    assert counts[True] < 0.5 * counts[False], counts


def test_warm_started_scan_never_loses_the_argmax(oracle):
    """k_screen_seed restated: the scan of a scenario starts from the score of ANY vertices (here: the winners at
    another point, random vertices, none) -- the survivors must still contain the oracle's argmax, and with the
    winners of a nearby point far fewer vertices are emitted than by the cold scan."""
    P, z = load_instance("storm")
    pool = load_pool("storm", 2048)
    vals = sample_instance_values(z, 160, seed=8)
    Sx = np.asarray(P.pos_row)
    D = vals - P.rbar[Sx][None, :]
    x0, x1 = z["x_ev"], z["x_ev"] + 0.03 * (z["x_alt"] - z["x_ev"])
    _, w0 = oracle.argmax_procedure(P, vals, x0, pool)           # winners at the previous point
    ov, oi = oracle.argmax_procedure(P, vals, x1, pool)
    bias = pool @ (P.rbar - P.T_dense() @ x1)
    rng = np.random.default_rng(3)
    emitted = {}
    for name, prev in (("cold", None), ("previous winners", np.stack([w0, w0[::-1]], axis=1)),
                       ("random vertices", rng.integers(-1, len(pool), size=(len(vals), 2)))):
        mask, n_emit = S.screen_scan(bias, pool[:, Sx], D, prev)
        assert mask[np.arange(len(vals)), oi].all(), name       # the argmax survives
        exact = bias[None, :] + D @ pool[:, Sx].T
        assert np.array_equal(np.where(mask, exact, -np.inf).argmax(axis=1), exact.argmax(axis=1)), name
        emitted[name] = n_emit.mean()
    assert emitted["previous winners"] < 0.5 * emitted["cold"], emitted


def test_scan_bounds_on_random_shapes_and_magnitudes():
    """The bound and the seed slack, restated: 400 random problems -- 1 to 40 rows, operands as clouds far from the
    origin or around it, scales from 1e-3 to 1e5, biases up to 1e9 above the dots, exact and 2^-40 / 2^-20 near-ties,
    random seeds -- and every vertex attaining the FP64 maximum of a scenario survives the scan, cold or seeded."""
    rng = np.random.default_rng(2026)
    for _ in range(200):
        s, K, N = int(rng.integers(1, 41)), int(rng.integers(2, 200)), int(rng.integers(1, 40))
        sp_, sd_ = 10.0 ** rng.uniform(-3, 5), 10.0 ** rng.uniform(-3, 4)
        PiS = rng.standard_normal(s) * sp_ * 10.0 ** rng.uniform(-2, 2) + rng.standard_normal((K, s)) * sp_
        D = rng.standard_normal(s) * sd_ * 10.0 ** rng.uniform(-2, 2) + rng.standard_normal((N, s)) * sd_
        bias = rng.standard_normal(K) * sp_ * sd_ * 10.0 ** rng.uniform(-2, 6) + 10.0 ** rng.uniform(0, 9) * rng.choice([0, 1, -1])
        for k in range(0, K - 1, 7):
            PiS[k + 1] = PiS[k] * (1 + rng.choice([0, 2.0 ** -40, 2.0 ** -20]))
            bias[k + 1] = bias[k] * (1 + rng.choice([0, 2.0 ** -45]))
        exact = bias[None, :] + D @ PiS.T
        ties = exact == exact.max(axis=1, keepdims=True)
        for prev in (None, rng.integers(-1, K, size=(N, 2))):
            mask, _ = S.screen_scan(bias, PiS, D, prev)
            assert (mask | ~ties).all()


@pytest.mark.parametrize("centre", [False, True])
def test_screened_argmax_with_ties_near_ties_and_dominant_bias(oracle, centre):
    """Adversarial pool: exact duplicates of the winner's stochastic part (first index must win), vertices one
    ulp-scale step away, a bias six orders above the dots, a NaN vertex, an all-zero scenario."""
    P = synthetic_problem(m2=90, n1=10, s=40)
    N, K = 300, 400
    vals = synthetic_values(P, N)
    vals[7] = P.rbar[P.pos_row]                              # delta == 0
    pool = synthetic_pool(P.m2, K, scale=30.0)
    pool[50] = pool[10]                                      # exact duplicate scores: index 10 must win over 50
    pool[51] = pool[10] * (1.0 + 2.0 ** -40)                 # a near tie
    det = [j for j in range(P.m2) if j not in set(P.pos_row.tolist())]
    pool[:, det[0]] *= 1.0e6                                 # bias >> dot
    pool[200, 3] = np.nan                                    # never wins (subprob.jl:156)
    x = 10.0 * oracle.u01(3, np.arange(P.n1))
    ov, oi = oracle.argmax_procedure(P, vals, x, pool)
    mv, mi, ncand = S.argmax_screened(P, vals, x, pool, centre=centre)
    assert np.array_equal(mi, oi) and np.array_equal(mv, ov)
    assert not (mi == 200).any() and not (mi == 50).any()
    assert ncand.max() < K
