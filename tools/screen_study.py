"""Offline study (CPU, numpy) behind DESIGN.md section 10: how many (scenario, vertex) pairs would survive a
reduced-precision screening pass with a rigorous error bound, on the real instances' templates, outcome tables
and harvested pools.  Not product code; reads the committed fixtures and the oracle's counter generator."""
import numpy as np, sys
sys.path.insert(0,'/root/repo')
from tests.helpers import load_instance, sample_instance_values, synthetic_pool
from oracle import oracle as O
def trunc(x, bits):   # keep `bits` mantissa bits (round to nearest) -- bf16: 8, tf32: 11 (10 explicit)
    m, e = np.frexp(x); return np.ldexp(np.round(m * 2.0**bits) / 2.0**bits, e)
def split(x, bits, parts):
    out=[]; r=x.copy()
    for _ in range(parts):
        h=trunc(r,bits); out.append(h); r=r-h
    return out
for name in ["baa99-20","ssn","storm"]:
    P,z=load_instance(name)
    N=1500
    vals=sample_instance_values(z,N,seed=3)
    pool=z["pool"]
    K=len(pool)
    S=P.pos_row
    x=z["x_alt"]
    Tm=P.T_dense(); base=P.rbar-Tm@x
    bias=pool@base
    D=vals-P.rbar[S][None,:]          # [N,s]
    PS=pool[:,S]                       # [K,s]
    exact=bias[None,:]+D@PS.T          # [N,K]
    best=exact.max(1); arg=exact.argmax(1)
    srt=np.sort(exact,axis=1); gap=(srt[:,-1]-srt[:,-2])
    print(f"{name}: K={K} s={len(S)} median rel gap {np.median(gap/np.abs(best)):.2e}; |bias| {np.abs(bias).mean():.3g}, |dot| {np.abs(D@PS.T).mean():.3g}")
    absdot=np.abs(D)@np.abs(PS).T
    cs=np.linalg.norm(D,axis=1)[:,None]*np.linalg.norm(PS,axis=1)[None,:]
    for label,bits,parts,terms in [("tf32",11,1,1),("bf16x2 (3 products)",8,2,3),("bf16x3 (6 products)",8,3,6)]:
        dp=split(D,bits,parts); pp=split(PS,bits,parts)
        approx=np.zeros_like(exact)
        # products kept: all (a,b) with a+b < parts  (fp32 accumulate emulated in float64 + fp32 rounding of the sum)
        for a in range(parts):
            for b in range(parts-a):
                approx+= (dp[a]@pp[b].T)
        approx=approx.astype(np.float32).astype(np.float64)      # fp32 accumulator
        err=np.abs(approx-(D@PS.T))
        eps_emp=(err/np.maximum(absdot,1e-300)).max()
        # rigorous-ish bound: dropped terms 2^-(bits*parts)... use eps = 2^-(bits*parts-1)*... + fp32 accumulate s*2^-24
        eps=2.0**(-(bits*parts)+2)+len(S)*2.0**-24
        bound_abs=eps*absdot; bound_cs=eps*cs
        sc=bias[None,:]+approx
        for bl,bd in (("abs-dot bound",bound_abs),("Cauchy-Schwarz bound",bound_cs)):
            lower=(sc-bd).max(1)                # best guaranteed lower bound
            cand=(sc+bd)>=lower[:,None]
            ok=cand[np.arange(N),arg].all()
            print(f"   {label:22s} eps={eps:.2e} (empirical max {eps_emp:.2e}) {bl:22s}: candidates/scenario mean {cand.sum(1).mean():7.2f} max {cand.sum(1).max():5d}  of K={K}; true argmax kept: {ok}")
