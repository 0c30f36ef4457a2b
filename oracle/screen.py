"""CPU restatement of the SCREENING pass (kernel: ``csrc/kernels_screen.cuh``).  Test
infrastructure like the rest of ``oracle/``: nothing in the product imports it.

The argmax of subprob.jl:141-169 only needs the exact FP64 score of the vertices that can still win.  The
screening pass computes every score approximately with bf16 operands (each fp64 operand split into two bf16
parts, three products, fp32 accumulation -- what ``tcgen05.mma kind::f16`` does with fp32 accumulators in TMEM),
bounds its error rigorously, and keeps per scenario the vertices whose upper bound reaches the best lower
bound.  The exact arithmetic (the oracle's, ``argmax_procedure``) then decides among the candidates, so the
selected vertex is the one a full FP64 sweep selects: first index among the exact maxima.

Error budget for one dot of length s (u_b = 2**-8, bf16 round to nearest: 8 significant bits; u_s = 2**-24, fp32):
  x = x_h + x_l + r_x with |r_x| <= u_b**2 |x| (two successive bf16 roundings of fp64 values, the second of
  the exact remainder); kept products x_h y_h + x_h y_l + x_l y_h, each exact in fp32 (8 x 8 significant
  bits); dropped: x_l y_l, r_x y, x r_y  ->  at most (u_b**2 + 2 u_b**2 (1 + u_b)**2 ...) |x||y|, bounded below
  by 4 u_b**2 |x||y|;  the 3 s products are accumulated in fp32 in an unspecified order: at most
  gamma = (3 s + 2) * 2 u_s of sum |terms| <= (1 + u_b)**2 ... (the factor 2 covers truncating accumulators).
  With Cauchy-Schwarz, sum_j |pi_j| |d_j| <= ||pi|| ||d||, so
      |approx - exact| <= EPS(s) * ||pi_k|_S|| * ||d_i||,   EPS(s) = 1.01 * (4 u_b**2 + (3 s + 2) * 2 u_s).
  Both norms are one cheap pass each (K + N values), rounded UP.

Centred operands (``centre=True``): for any fixed vectors c and dbar,
      bias_k + pi_k . d_i = [bias_k + pi_k . dbar] + (pi_k - c) . (d_i - dbar) + c . (d_i - dbar),
  and the last term does not depend on k, so the argmax over k is that of the first two.  The pass multiplies the
  centred operands; every term of the bound is then relative to ||pi_k - c|| ||d_i - dbar||, which on real pools
  (clouds around a common point) is an order of magnitude below ||pi_k|| ||d_i||.  The FP64 roundings of the
  subtractions, of pi_k . dbar and of the exact scores themselves are bounded on the UNcentred magnitudes
  (``eabs``).
"""
from __future__ import annotations

import numpy as np

U_B = 2.0 ** -8
U_S = 2.0 ** -24


def bf16_round(x: np.ndarray) -> np.ndarray:
    """fp64 -> nearest bf16 (ties to even), returned as fp64.  Values are assumed in bf16's normal range."""
    f = np.asarray(x, dtype=np.float64).astype(np.float32)          # first rounding: fp64 -> fp32 (RN)
    b = f.view(np.uint32).astype(np.uint64)
    b = (b + 0x7FFF + ((b >> 16) & 1)) & 0xFFFF0000                 # RN-even on the low 16 bits
    out = b.astype(np.uint32).view(np.float32).astype(np.float64)
    return out.reshape(np.shape(x))


def split2(x: np.ndarray):
    """x ~ hi + lo, both bf16.  The double rounding fp64 -> fp32 -> bf16 costs at most one extra 2**-24
    relative, which EPS's 1.01 factor and the 4 u_b**2 (against the tight 3 u_b**2) absorb."""
    hi = bf16_round(x)
    lo = bf16_round(np.asarray(x, dtype=np.float64) - hi)
    return hi, lo


def eps(s: int) -> float:
    return 1.01 * (4.0 * U_B * U_B + (3 * s + 2) * 2.0 * U_S)


def approx_dots(PiS: np.ndarray, D: np.ndarray) -> np.ndarray:
    """[N, K] approximate PiS[k] . D[i] from the three bf16 products, accumulated in fp32."""
    ph, pl = (a.astype(np.float32) for a in split2(PiS))
    dh, dl = (a.astype(np.float32) for a in split2(D))
    acc = dh @ ph.T
    acc = acc + dh @ pl.T
    acc = acc + dl @ ph.T
    return acc.astype(np.float64)


def up(x):
    return np.nextafter(x, np.inf)


CENTRE_COLS, CENTRE_SCEN = 1024, 256          # kernels_screen.cuh SCR_CENTRE_COLS / SCR_CENTRE_SCEN


def screen(bias: np.ndarray, PiS: np.ndarray, D: np.ndarray, centre: bool = True):
    """Candidate mask [N, K]: vertex k stays for scenario i iff its upper bound reaches the best lower bound.
    bias[k] is exact (fp64).  NaN / -Inf biases (which never win, subprob.jl:156) are never candidates unless
    nothing else is."""
    s = PiS.shape[1]
    eabs = 0.0
    if centre and len(PiS) and len(D):
        with np.errstate(all="ignore"):
            c = np.nan_to_num(PiS[:CENTRE_COLS].mean(axis=0), nan=0.0, posinf=0.0, neginf=0.0)
            dbar = np.nan_to_num(D[:CENTRE_SCEN].mean(axis=0), nan=0.0, posinf=0.0, neginf=0.0)
            bias = bias + PiS @ dbar
            raw_p = np.sqrt((PiS * PiS).sum(axis=1))
            raw_d = np.sqrt((D * D).sum(axis=1))
            fin = np.isfinite(bias)
            eabs = (s + 8) * 2.0 ** -52 * ((np.abs(bias[fin]).max() if fin.any() else 0.0) +
                                           np.nanmax(raw_p[np.isfinite(raw_p)], initial=0.0) *
                                           np.nanmax(raw_d, initial=0.0) * 2.0)
        PiS, D = PiS - c, D - dbar
    approx = approx_dots(PiS, D)
    pn = up(np.sqrt(up((PiS * PiS).sum(axis=1))))
    dn = up(np.sqrt(up((D * D).sum(axis=1))))
    bound = up(eps(s) * np.outer(dn, pn)) + eabs
    sc = bias[None, :] + approx
    finite = np.isfinite(sc)
    lower = np.where(finite, sc - bound, -np.inf).max(axis=1)
    slack = np.abs(sc) * 2.0 ** -50 + np.abs(bias)[None, :] * 2.0 ** -50      # the fp64 roundings of sc +- bound
    return finite & (sc + bound + slack >= (lower - np.abs(lower) * 2.0 ** -50)[:, None])


def screen_scan(bias: np.ndarray, PiS: np.ndarray, D: np.ndarray, prev=None, centre: bool = True):
    """The scan as the kernel runs it (k_screen's epilogue + k_screen_seed), restated column by column: a scenario's
    lower bound L starts from the score of its previous winners (`prev[i]` = up to two vertex indices, -1 = none;
    ANY vertices give a valid bound), a vertex is EMITTED when its upper bound reaches the running L, and the entries
    that still reach the final L survive.  Returns (survivor mask [N, K], emitted count per scenario)."""
    s = PiS.shape[1]
    N, K = len(D), len(PiS)
    with np.errstate(all="ignore"):
        if centre and K and N:
            c = np.nan_to_num(PiS[:CENTRE_COLS].mean(axis=0), nan=0.0, posinf=0.0, neginf=0.0)
            dbar = np.nan_to_num(D[:CENTRE_SCEN].mean(axis=0), nan=0.0, posinf=0.0, neginf=0.0)
        else:
            c, dbar = np.zeros(s), np.zeros(s)
        b1 = bias + PiS @ dbar
        raw_p, raw_d = np.sqrt((PiS * PiS).sum(axis=1)), np.sqrt((D * D).sum(axis=1))
        fin = np.isfinite(b1)
        eabs = (s + 8) * 2.0 ** -52 * ((np.abs(b1[fin]).max() if fin.any() else 0.0) +
                                       np.nanmax(raw_p[np.isfinite(raw_p)], initial=0.0) * np.nanmax(raw_d, initial=0.0) * 2.0)
        P1, D1 = PiS - c, D - dbar
        approx = approx_dots(P1, D1)
        pn = up(np.sqrt(up((P1 * P1).sum(axis=1))))
        dn = up(np.sqrt(up((D1 * D1).sum(axis=1))))
        E = up(eps(s) * np.outer(dn, pn)) + eabs + 2.0 ** -50 * (np.abs(b1)[None, :] + np.abs(approx))
        t = b1[None, :] + approx
        L = np.full(N, -np.inf)
        if prev is not None:                                   # k_screen_seed: exact centred score of the previous winners
            prev = np.asarray(prev).reshape(N, -1)
            for j in range(prev.shape[1]):
                k = prev[:, j]
                ok = (k >= 0) & (k < K)
                kk = np.where(ok, k, 0)
                sc = b1[kk] + np.einsum("ij,ij->i", P1[kk], D1) - 2.0 * eabs
                sc = sc - 2.0 ** -40 * np.abs(sc)
                sc = np.where(ok & np.isfinite(sc), sc.astype(np.float32).astype(np.float64) - np.abs(sc) * 2.0 ** -23, -np.inf)
                L = np.maximum(L, sc)
        emitted = np.zeros((N, K), dtype=bool)
        for k in range(K):                                     # the scan, in column order
            tk, Ek = t[:, k], E[:, k]
            good = np.isfinite(tk)
            emitted[:, k] = good & (tk + Ek >= L)
            L = np.where(good, np.maximum(L, tk - Ek), L)
        ub = np.where(np.isfinite(t), t + E, -np.inf)
    return emitted & (ub >= L[:, None]), emitted.sum(axis=1)


def argmax_screened(P, values, x, pool, centre: bool = True):
    """max_val, max_idx as ``oracle.argmax_procedure`` -- exact arithmetic on the candidates only -- plus the
    candidate counts per scenario."""
    from . import oracle as O
    values = np.asarray(values, dtype=np.float64).reshape(-1, max(P.s, 1))
    pool = np.asarray(pool, dtype=np.float64).reshape(-1, P.m2)
    N, K = len(values), len(pool)
    assert all(c < 0 for c in P.pos_col), "screening is stated for the delta_T == 0 case"
    S = np.asarray(P.pos_row)
    base = P.rbar - P.T_dense() @ np.asarray(x, dtype=np.float64)
    bias = pool @ base
    D = values - P.rbar[S][None, :]
    mask = screen(bias, pool[:, S], D, centre)
    mv, mi = np.full(N, -np.inf), np.full(N, -1, dtype=np.int64)
    for i in range(N):
        cand = np.nonzero(mask[i])[0]
        if len(cand) == 0:
            continue
        v, j = O.argmax_procedure(P, values[i:i + 1], x, pool[cand])
        mv[i], mi[i] = v[0], (cand[j[0]] if j[0] >= 0 else -1)
    return mv, mi, mask.sum(axis=1)
