#!/usr/bin/env python
"""bench.py -- cut-formation throughput of the B200 path (BASELINE.json's metric).

A *step* is the cut formation of one SD iteration (reference ``sd_iteration!``,
algorithm.jl:45-55 and :79-85, without the LP/QP solves) on the storm shape
(SURVEY.md 8(d) C4): for each of E = 4 weighted epigraphs append 1 scenario, push 2 dual
vertices (one new, one duplicate), then build the candidate cut and the regenerated
incumbent cut.  One *evaluation* is one score[k, i] at one x (subprob.jl:155-158), so a
step performs  2 * K * N  evaluations over all epigraphs.

  python bench.py [--gpus N] [--steps K] [--warmup W]            (N > 1: launched by torchrun)
  python bench.py --impl reference ...     the reference's CPU loop (oracle restatement)

Prints ONE JSON line (rank 0).  Workload is weak-scaled: every GPU holds --scen-per-gpu
scenarios; the pool is replicated.  Inputs (the scenario store, ~0.94 GB per GPU) are
larger than L2, so no explicit L2 flush is needed between steps.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "scenario_x_vertex_argmax_evals_per_sec"
UNIT = "evals/s"
FP64_PEAK_FILE = os.path.join(ROOT, "profiles", "fp64_peak.json")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scen-per-gpu", type=int, default=1_000_000)
    ap.add_argument("--vertices", type=int, default=16384)
    ap.add_argument("--epigraphs", type=int, default=4)
    ap.add_argument("--instance", default="storm")
    ap.add_argument("--pool", default="real", choices=["real", "synthetic"],
                    help="dual-vertex pool of the headline leg: harvested LP duals (tests/golden/pools, SURVEY C4) or the "
                         "round-1 pool (94 harvested + uniform synthetic vertices)")
    ap.add_argument("--screen", type=int, default=-1, help="override the library's screening mode (0 off, 1 auto, 2 always)")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the other-pool leg and the strong-scaled leg")
    ap.add_argument("--parity-n", type=int, default=512, help="scenarios of the bench's own state checked against the oracle")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU-baseline work")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dev-only", action="store_true",
                    help="kernel experiments: skip the end-to-end leg and the result checks")
    return ap.parse_args()


# ---------------------------------------------------------------- workload ---------------

def load_instance(name):
    if name.startswith("synth"):   # synth<s> or synth<s>T<n_T>
        body = name[5:] or "128"
        s_, _, t_ = body.partition("T")
        return synthetic_instance(int(s_), int(t_ or 0))
    z = dict(np.load(os.path.join(ROOT, "tests", "golden", "instances", f"{name}.npz")))
    return z


def synthetic_instance(s, n_T=0):
    """The storm-shaped synthetic template of SURVEY.md 8(d) C5: m2 = 4 s rows of which the first s
    are stochastic (RHS), n1 = s, Tbar = one -1 per first-stage column on rows s..2s-1,
    rbar_j = 100 + 400 u(6, j) on the stochastic rows, five equiprobable outcomes rbar_j * {0.8..1.2}.
    With n_T > 0 (the "delta-T variant") n_T stored entries of Tbar are random too, UNIFORM on
    Tbar * (0.9 .. 1.1): the path no shipped instance exercises (d depends on x, one contraction per point)."""
    m2, n1 = 4 * s, s
    rbar = np.zeros(m2)
    rbar[:s] = 100.0 + 400.0 * u01(6, np.arange(s))
    fac = np.array([0.8, 0.9, 1.0, 1.1, 1.2])
    tcols = (np.arange(n_T) * max(1, s // max(n_T, 1))) % n1
    ne = s + n_T
    out_vals = np.ones((ne, 5)); out_vals[:s] = rbar[:s, None] * fac[None, :]
    z = {"m2": np.int64(m2), "n1": np.int64(n1), "rbar": rbar,
         "T_colptr": np.arange(n1 + 1, dtype=np.int64), "T_rowval": np.arange(s, 2 * s, dtype=np.int64),
         "T_nzval": -np.ones(n1), "pos_row": np.concatenate([np.arange(s), s + tcols]).astype(np.int32),
         "pos_col": np.concatenate([-np.ones(s), tcols]).astype(np.int32), "out_vals": out_vals,
         "out_cdf": np.tile(np.cumsum(np.full(5, 0.2)), (ne, 1)), "out_cnt": np.full(ne, 5, dtype=np.int32),
         "pool": np.zeros((0, m2)), "x_ev": 10.0 * u01(3, np.arange(n1)), "x_alt": 10.0 * u01(5, np.arange(n1))}
    if n_T:
        z["kind"] = np.concatenate([np.zeros(s), np.full(n_T, 2)]).astype(np.int32)
        z["par_a"] = np.concatenate([np.zeros(s), np.full(n_T, -1.1)])
        z["par_b"] = np.concatenate([np.zeros(s), np.full(n_T, -0.9)])
    return z


def u01(seed, idx):
    idx = np.asarray(idx, dtype=np.uint64)
    with np.errstate(over="ignore"):
        zz = np.uint64(seed) ^ (idx * np.uint64(0x9E3779B97F4A7C15))
        zz = zz + np.uint64(0x9E3779B97F4A7C15)
        zz = (zz ^ (zz >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        zz = (zz ^ (zz >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        zz = zz ^ (zz >> np.uint64(31))
    return (zz >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def real_pool(name, n):
    """The first n vertices of the largest committed harvest of this instance (tools/harvest_pool.py: stage-2 LP
    duals through the reference's dedup rule, in LP order -- a prefix of a harvest IS the smaller harvest)."""
    d = os.path.join(ROOT, "tests", "golden", "pools")
    best = None
    for f in os.listdir(d) if os.path.isdir(d) else []:
        if f.startswith(name + "_K") and f.endswith(".npz"):
            k = int(f[len(name) + 2:-4])
            if k >= n and (best is None or k < best[0]):
                best = (k, f)
    if best is None:
        return None
    return np.ascontiguousarray(np.load(os.path.join(d, best[1]))["pool"][:n])


def make_pool(z, K, extra, kind="synthetic", name=None):
    """K + extra distinct vertices.  kind "real": harvested LP duals only.  kind "synthetic": the
    round-1 pool -- the 94 harvested vertices of the instance fixture, then pi_kj = scale * (2u - 1)."""
    m2 = int(z["m2"])
    if kind == "real" and name and not name.startswith("synth"):
        P = real_pool(name, K + extra)
        if P is not None:
            return P
        raise SystemExit(f"no harvested pool with {K + extra} vertices for {name}: run tools/harvest_pool.py")
    real = z["pool"]
    n_syn = K + extra - len(real)
    scale = float(np.abs(real).max()) if len(real) else 1000.0
    syn = scale * (2.0 * u01(2, np.arange(n_syn * m2, dtype=np.uint64)).reshape(n_syn, m2) - 1.0)
    return np.vstack([real, syn])


def sample_values(z, seed, g0, n):
    """Host twin of the device sampler (sqlp_epi_sample_scenarios)."""
    s = len(z["pos_row"])
    g = np.arange(g0, g0 + n, dtype=np.uint64)
    u = u01(seed, (g[:, None] * np.uint64(s) + np.arange(s, dtype=np.uint64)[None, :]))
    idx = (u[:, :, None] >= z["out_cdf"][None, :, :]).sum(axis=2)
    idx = np.minimum(idx, np.maximum(z["out_cnt"][None, :] - 1, 0))
    out = np.take_along_axis(np.broadcast_to(z["out_vals"], (n,) + z["out_vals"].shape),
                             idx[:, :, None], 2)[:, :, 0].copy()
    if "kind" in z:   # UNIFORM elements: left + (right - left) * ((bits + 1/2) / 2^53), like the device
        uo = u - (0.0) + 0.5 / 9007199254740992.0
        cont = z["kind"] == 2
        out[:, cont] = z["par_a"][cont] + (z["par_b"][cont] - z["par_a"][cont]) * uo[:, cont]
    return out


# ---------------------------------------------------------------- clocks -----------------

class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML during the timed region."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake"}

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.samples, self.reasons, self.max_mhz = device, [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        while not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.004)        # a timed region is a few tens of ms: several samples inside it

    def finish(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------- CPU reference arm -----

def cpu_reference_single_thread(z, pool, n_target_seconds, x2):
    """The same loop the way the reference itself runs it: ONE thread (the reference has no threading,
    so its "Julia thread count" is 1) and the oracle's index-order dots, on a short prefix."""
    from oracle import oracle as O
    P = O.Problem(int(z["m2"]), int(z["n1"]), z["rbar"], z["T_colptr"], z["T_rowval"], z["T_nzval"],
                  z["pos_row"], z["pos_col"])
    probe = sample_values(z, 1, 0, 2)
    t0 = time.perf_counter()
    O.bench_argmax(P, probe, x2[0], pool, threads=1, dot_kind=0)
    rate = len(probe) / (time.perf_counter() - t0)
    n = int(max(2, min(2000, rate * n_target_seconds / 2)))
    vals = sample_values(z, 1, 0, n)
    t0 = time.perf_counter()
    for x in x2:
        O.bench_argmax(P, vals, x, pool, threads=1, dot_kind=0)
    dt = time.perf_counter() - t0
    return {"value": 2.0 * n * len(pool) / dt, "unit": UNIT, "cores": 1,
            "sample": f"first {n} scenarios x all {len(pool)} vertices x 2 points in {dt:.2f} s; one thread, "
                      f"index-order dots (the reference's own structure: single-threaded Julia)"}


def cpu_reference(z, pool, n_target_seconds, x2, threads=0):
    """The reference's argmax loop (oracle restatement, reference loop structure, all host
    cores, 8-accumulator dots) on a bounded prefix of the workload's scenarios."""
    from oracle import oracle as O
    P = O.Problem(int(z["m2"]), int(z["n1"]), z["rbar"], z["T_colptr"], z["T_rowval"], z["T_nzval"],
                  z["pos_row"], z["pos_col"])
    # every host core this process may run on: torchrun exports OMP_NUM_THREADS=1 to its ranks, which
    # would silently turn the "all host threads" baseline into a single-threaded one
    nthr = (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else O.max_threads()) \
        if threads <= 0 else threads
    probe = sample_values(z, 1, 0, 2 * nthr)
    t0 = time.perf_counter()
    O.bench_argmax(P, probe, x2[0], pool, threads=nthr)
    dt = time.perf_counter() - t0
    rate = len(probe) / dt                       # scenarios / s for one x
    n = int(max(2 * nthr, min(20000, rate * n_target_seconds / 2)))
    vals = sample_values(z, 1, 0, n)
    t0 = time.perf_counter()
    for x in x2:                                 # the reference runs candidate and incumbent
        O.bench_argmax(P, vals, x, pool, threads=nthr)     # as two separate passes
    dt = time.perf_counter() - t0
    evals = 2.0 * n * len(pool)
    return {"value": evals / dt, "unit": UNIT, "cores": nthr, "kind": "port",
            "sample": f"first {n} scenarios x all {len(pool)} vertices x 2 points in {dt:.2f} s; "
                      f"C restatement of the reference loop (not Julia), OpenMP over scenarios, "
                      f"8-accumulator dots"}, dt, evals


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    z = load_instance(args.instance)
    pool = make_pool(z, args.vertices, 0, pool_kind(args), args.instance)
    n1 = int(z["n1"])
    x2 = [z["x_ev"], z["x_alt"]] if "x_ev" in z else [10 * u01(3, np.arange(n1)), 10 * u01(5, np.arange(n1))]
    per_step = max(1.0, min(30.0, 150.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_reference(z, pool, min(per_step, 2.0), x2)
    t_tot, e_tot, base = 0.0, 0.0, None
    for _ in range(args.steps):
        base, dt, ev = cpu_reference(z, pool, per_step, x2)
        t_tot += dt
        e_tot += ev
    val = e_tot / t_tot
    base["value"] = val
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": workload_config(args, z, args.gpus),
            "cpu_baseline": base,
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def pool_kind(args):
    return "synthetic" if args.instance.startswith("synth") else args.pool


def workload_config(args, z, world):
    return {"workload": f"{args.instance} shape: cut formation of one SD iteration "
                        f"(append 1 scenario + 2 pool pushes per epigraph, candidate + incumbent cut)",
            "pool": ("harvested stage-2 LP duals (HiGHS) through the reference's dedup rule" if pool_kind(args) == "real"
                     else "94 harvested + uniform synthetic vertices"),
            "instance": args.instance, "s": int(len(z["pos_row"])), "m2": int(z["m2"]), "n1": int(z["n1"]),
            "K_vertices": args.vertices, "N_scenarios_per_gpu": args.scen_per_gpu,
            "N_scenarios_total": args.scen_per_gpu * world, "epigraphs": args.epigraphs,
            "weights": "0.5+u", "points_per_step": 2,
            "points": "new candidate every iteration (<= 5 % step + 1 % jitter), incumbent replaced every 4th", "parallelism": f"scenario-shard x{world}",
            "l2": "inputs larger than L2 (scenario store 8*s_pad*N bytes per GPU)"}


# ---------------------------------------------------------------- our arm ----------------

class Cell:
    """A pool + E weighted epigraphs with their scenarios sampled on the device."""

    def __init__(self, T, ctx, z, coef, pool_all, K0, E, n_epi_global):
        self.T, self.ctx, self.z, self.E, self.K0, self.n_epi_global = T, ctx, z, E, K0, n_epi_global
        self.pool_all = pool_all
        m2 = int(z["m2"])
        self.dvs = T.sdDualVertexSet(ctx=ctx, m2=m2)
        ins, _ = self.dvs.push_many(pool_all[:K0])
        assert ins.all() and len(self.dvs) == K0, "workload vertices are not distinct under the dedup rule"
        self.epis = []
        for e in range(E):
            epi = T.sdEpigraph(coef, 1.0 / E, 0.0, self.dvs)
            epi.set_outcomes(z["out_vals"], z["out_cdf"], z["out_cnt"])
            if "kind" in z:
                epi.set_distributions(z["kind"], z["par_a"], z["par_b"])
            epi.sample_scenarios(n_epi_global, seed=101 + e, weight_seed=201 + e)
            self.epis.append(epi)
        self.t_next = 0          # SD iterations performed so far on this cell

    def step_inputs(self, t):
        """Iteration t: E new scenarios, E new vertices + E duplicates (pushed new, dup, new, dup ...)."""
        z, E, K0, m2 = self.z, self.E, self.K0, int(self.z["m2"])
        g = self.n_epi_global + t
        scen = np.stack([sample_values(z, 101 + e, g, 1)[0] for e in range(E)])      # [E, s]
        newv = self.pool_all[K0 + t * E: K0 + (t + 1) * E]                            # [E, m2]
        dupv = self.pool_all[(7 * t + 3 * np.arange(E)) % K0]
        verts = np.empty((2 * E, m2))
        verts[0::2] = newv
        verts[1::2] = dupv
        return scen, verts

    def evals_of(self, t):
        """Evaluations of iteration t: 2 points x K x N with K, N after this iteration's pushes / scenarios."""
        return 2.0 * (self.K0 + self.E * (t + 1)) * ((self.n_epi_global + t + 1) * self.E)

    def close(self):
        for e in self.epis:
            e.close()
        self.dvs.close()


def run_ours(args):
    import torch
    import torch.distributed as dist
    from sqlp_b200 import twosd as T
    from sqlp_b200 import _lib
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        ids = [T.Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx = T.Context(local, rank, world, ids[0])
    else:
        ctx = T.Context(local)
    stream = torch.cuda.Stream(device=dev)     # a real (non-NULL) stream shared with the library
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    if args.screen >= 0:
        ctx.set_screen(args.screen)
    L = _lib.lib()

    z = load_instance(args.instance)
    m2, n1, s = int(z["m2"]), int(z["n1"]), len(z["pos_row"])
    E, K0 = args.epigraphs, args.vertices
    n_steps = args.steps + args.warmup
    total_steps = 2 * n_steps + 2            # device-resident leg + e2e leg (+ slack)
    kind = pool_kind(args)
    pool_all = make_pool(z, K0, E * total_steps, kind, args.instance)
    x_ev = np.ascontiguousarray(z["x_ev"])
    x_alt = np.ascontiguousarray(z["x_alt"])

    def points_of(t):
        """Candidate and incumbent of SD iteration t.  As in an SD run the candidate is a new point every iteration
        (here: a step of up to 5 % of the way from the EV solution towards the alternative point plus a 1 % jitter of
        every coordinate) and the incumbent is replaced by an earlier candidate every fourth iteration -- nothing
        in the timed region sees the same pair of points twice."""
        def cand(k):
            return x_ev + 0.05 * float(u01(11, k)) * (x_alt - x_ev) + 0.01 * np.abs(x_ev) * (u01(13 + k, np.arange(n1)) - 0.5)
        k_inc = (t // 4) * 4 - 1
        return np.ascontiguousarray(cand(t)), np.ascontiguousarray(x_alt if k_inc < 0 else cand(k_inc))
    coef = T.sdSubprobCoefficients.from_tables(z["rbar"], z["T_colptr"], z["T_rowval"], z["T_nzval"],
                                               z["pos_row"], z["pos_col"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        tns = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(tns, op=dist.ReduceOp.MAX)
        return float(tns.item())


    def device_leg(cell, warmup, steps, want_profiles):
        """`warmup` untimed + `steps` timed SD iterations with every input already resident in HBM and no host
        synchronisation inside the timed region.  One library call per class of work: E x add_scenario!, ONE push
        of the iteration's 2E vertices, E x (candidate + incumbent cut)."""
        Ecell = cell.E
        out_dev = torch.zeros(Ecell, 2, n1 + 2, dtype=torch.float64, device=dev)
        t0 = cell.t_next
        staged = []
        for t in range(warmup + steps):
            scen, verts = cell.step_inputs(t0 + t)
            staged.append((torch.from_numpy(scen).to(dev), torch.from_numpy(verts).to(dev),
                           torch.from_numpy(np.concatenate(points_of(t0 + t))).to(dev)))
        torch.cuda.synchronize()

        cell_handles = (C.c_void_p * Ecell)(*[e._h for e in cell.epis])

        def dev_step(t):
            scen_d, verts_d, x2_dev = staged[t]
            for e, epi in enumerate(cell.epis):
                _lib.check(L.sqlp_epi_add_scenarios_dev(epi._h, 1, C.c_void_p(scen_d[e].data_ptr()), None))
            _lib.check(L.sqlp_pool_push_dev(cell.dvs._h, 2 * Ecell, C.c_void_p(verts_d.data_ptr())))
            _lib.check(L.sqlp_cell_build_cuts2_dev(Ecell, cell_handles, C.c_void_p(x2_dev.data_ptr()),
                                                   C.c_void_p(out_dev.data_ptr())))

        # warm-up steps: after the first one every kernel class is bracketed by events (roofline_other); the timed
        # steps bracket only the argmax kernels, so the instrumentation costs two event records per launch
        prof_warm = None
        for t in range(warmup):
            dev_step(t)
            if t == 0 and want_profiles:
                ctx.synchronize()
                ctx.profile(True)
                ctx.profile_classes(reset=True)
        if want_profiles:
            prof_warm = ctx.profile_classes(reset=True)
        barrier()
        ctx.profile(2)
        ctx.profile_classes(reset=True)
        launches0 = ctx.launch_count()
        sampler = ClockSampler(local)
        sampler.start()
        barrier()           # again: starting the clock sampler costs a rank a few ms of host time, and a rank that
        #                     enters the timed region early only waits for the others at the first exchange
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cuprof = os.environ.get("SQLP_BENCH_CUPROF") == "1"     # ncu --profile-from-start off: the timed steps only
        if cuprof:
            torch.cuda.cudart().cudaProfilerStart()
        ev0.record(stream)
        for t in range(warmup, warmup + steps):
            dev_step(t)
        ev1.record(stream)
        barrier()
        if cuprof:
            torch.cuda.cudart().cudaProfilerStop()
        clocks = sampler.finish()
        ms_mine = ev0.elapsed_time(ev1)
        ms = max_over_ranks(ms_mine)
        per_rank = None
        if world > 1:     # which GPU was the slow one, and at what clock (8 GPUs share a box's power budget)
            mine = torch.tensor([ms_mine, clocks["sm_mhz"] or 0.0, float(len(clocks["reasons"]))], dtype=torch.float64, device=dev)
            allr = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allr, mine)
            per_rank = {"ms": [round(float(a[0]), 3) for a in allr], "sm_mhz": [float(a[1]) for a in allr],
                        "throttle_reasons_seen": [int(a[2]) for a in allr]}
        launches = ctx.launch_count() - launches0
        prof = ctx.profile_classes(reset=True)
        ctx.profile(False)
        evals = sum(cell.evals_of(t0 + t) for t in range(warmup, warmup + steps))
        cell.t_next = t0 + warmup + steps
        K_after = len(cell.dvs)
        assert K_after == cell.K0 + Ecell * cell.t_next, (K_after, cell.K0, Ecell, cell.t_next)
        cols, nrel = cell.epis[0].view_columns()
        return {"ms": ms, "ms_per_step": ms / max(1, steps), "value": evals / (ms * 1e-3), "launches": launches,
                "prof": prof, "prof_warm": prof_warm, "clocks": clocks, "out": out_dev.cpu().numpy(),
                "screen": cell.epis[0].screen_stats(), "per_rank": per_rank,
                "sweep": {"pool_vertices": K_after, "columns_swept": cols, "relevant_rows": nrel, "rows": m2,
                          "note": "evaluations are counted as the reference performs them (2 points x K x N per "
                                  "iteration); vertices equal on every relevant row have bit-identical scores and "
                                  "only the first of a class can be the first maximum (subprob.jl:156), so the "
                                  "sweep visits one column per class -- same indices, values and cuts"}}

    # ---------------------------------------------------------------- headline cell ----------
    t_setup = time.perf_counter()
    ctx.profile(True)
    ctx.profile_classes(reset=True)
    n_epi_global = (args.scen_per_gpu * world) // E
    cell = Cell(T, ctx, z, coef, pool_all, K0, E, n_epi_global)
    ctx.synchronize()
    t_setup = time.perf_counter() - t_setup
    prof_setup = ctx.profile_classes(reset=True)

    # bulk add_scenario! from realised values (the 16 s bytes-per-scenario form of the delta build):
    # a scratch epigraph, values already on the device, timed by the library's event scopes
    n_bulk = min(262144, max(128, n_epi_global))
    vals_bulk = torch.from_numpy(sample_values(z, 7, 0, 4096)).to(dev).repeat((n_bulk + 4095) // 4096, 1)[:n_bulk].contiguous()
    scratch = T.sdEpigraph(coef, 1.0, 0.0, cell.dvs)
    for _ in range(4):
        _lib.check(L.sqlp_epi_add_scenarios_dev(scratch._h, n_bulk, C.c_void_p(vals_bulk.data_ptr()), None))
        ctx.synchronize()
        prof_bulk = ctx.profile_classes(reset=True)      # keeps the last (warm) one
    scratch.close()
    del scratch, vals_bulk
    ctx.profile(False)

    # ---- leg 1: device-resident inputs ---------------------------------------------------------
    leg = device_leg(cell, args.warmup, args.steps, True)
    ms_dev, value, launches, clocks = leg["ms"], leg["value"], leg["launches"], leg["clocks"]
    prof_steps, prof_warm, cut_check = leg["prof"], leg["prof_warm"], leg["out"]

    if args.dev_only:
        if rank == 0:
            print(json.dumps({"dev_only": True, "value": value, "ms_per_step": leg["ms_per_step"], "pool": kind,
                              "prof": prof_steps, "screen": leg["screen"]}))
        return

    # ---- leg 2: end to end through the blocking host API (host buffers in, cuts out) -------
    base_t = cell.t_next
    host_in = []
    for t in range(n_steps):
        scen, verts = cell.step_inputs(base_t + t)
        host_in.append((torch.from_numpy(scen).pin_memory(), torch.from_numpy(verts).pin_memory(),
                        [torch.from_numpy(x).pin_memory() for x in points_of(base_t + t)]))
    x_c, x_i = points_of(base_t + n_steps - 1)          # the last iteration's points: what the checks below use
    alpha = np.zeros((E, 2)); beta = np.zeros((E, 2, n1)); wm = np.zeros(E); val = np.zeros((E, 2))
    handles = (C.c_void_p * E)(*[e._h for e in cell.epis])
    ins = np.zeros(2 * E, dtype=np.int32); idx = np.zeros(2 * E, dtype=np.int64)

    def e2e_step(t):
        # the call a host makes once per SD iteration: scenarios, the iteration's dual vertices and both
        # points in (host buffers), dedup decisions and both cuts of every epigraph out
        scen_p, verts_p, (xc_p, xi_p) = host_in[t]
        _lib.check(L.sqlp_cell_sd_step(E, handles, C.c_void_p(scen_p.data_ptr()), None, 2 * E,
                                       C.c_void_p(verts_p.data_ptr()), ins.ctypes.data_as(C.c_void_p),
                                       idx.ctypes.data_as(C.c_void_p), C.c_void_p(xc_p.data_ptr()),
                                       C.c_void_p(xi_p.data_ptr()), alpha.ctypes.data_as(C.c_void_p),
                                       beta.ctypes.data_as(C.c_void_p), wm.ctypes.data_as(C.c_void_p),
                                       val.ctypes.data_as(C.c_void_p)))

    for t in range(args.warmup):
        e2e_step(t)
    barrier()
    t0 = time.perf_counter()
    for t in range(args.warmup, n_steps):
        e2e_step(t)
    barrier()
    ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3)
    evals_e2e = sum(cell.evals_of(base_t + t) for t in range(args.warmup, n_steps))
    cell.t_next = base_t + n_steps
    e2e_value = evals_e2e / (ms_e2e * 1e-3)
    h2d = (E * s + 2 * E * m2 + 2 * n1) * 8
    d2h = E * (2 * (n1 + 2) + 1) * 8 + 2 * E * 16
    assert list(ins) == [1, 0] * E, f"dedup decisions of the last step: {ins}"      # new, duplicate, new, ...

    # G5 on the last step's cuts: alpha + beta.x == sum_i p_i max_val_i
    for e in range(E):
        for xi_, x in enumerate((x_c, x_i)):
            lhs = alpha[e, xi_] + beta[e, xi_] @ x
            assert abs(lhs - val[e, xi_]) <= 1e-9 * (abs(alpha[e, xi_]) + np.abs(beta[e, xi_] * x).sum()), \
                "cut invariant violated"
    assert np.isfinite(cut_check).all()

    # ---- parity of the bench's OWN state against the oracle -----------------------------------
    parity = (parity_sample(args, T, dist if world > 1 else None, cell, world, rank, alpha, beta, x_c, x_i, pool_all)
              if args.parity_n > 0 else {"n": 0, "skipped": "--parity-n 0"})

    # ---- the other pool, and the strong-scaled form of the job, in the same run -----------------
    other = strong = None
    if not args.no_extra_legs and not args.instance.startswith("synth"):
        cell.close()
        del cell
        torch.cuda.empty_cache()
        okind = "synthetic" if kind == "real" else "real"
        opool = make_pool(z, K0, E * (n_steps + 1), okind, args.instance)
        ocell = Cell(T, ctx, z, coef, opool, K0, E, n_epi_global)
        oleg = device_leg(ocell, args.warmup, args.steps, False)
        other = leg_summary(oleg, okind, ocell)
        ocell.close()
        del ocell
        if world > 1:
            # C4 as BASELINE.json names it: N = scen_per_gpu scenarios IN TOTAL, sharded over the GPUs
            scell = Cell(T, ctx, z, coef, pool_all, K0, E, args.scen_per_gpu // E)
            sleg = device_leg(scell, args.warmup, args.steps, False)
            strong = leg_summary(sleg, kind, scell, traffic_shape=False)
            strong["N_total"] = args.scen_per_gpu
            # the weak-scaled step of this run does, per GPU, exactly the work of the 1-GPU form of this job
            strong["efficiency_vs_n1"] = (leg["ms_per_step"] / world) / sleg["ms_per_step"]
            strong["efficiency_note"] = ("1-GPU time taken from this run's weak-scaled step (the same per-GPU work: "
                                         f"{args.scen_per_gpu} scenarios x all vertices)")
            scell.close()
            del scell

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    rl = rooflines(leg, args)
    # the HBM-bound kernels: algorithmic bytes over event-timed device time
    hbm_peak, hbm_src = 6552.3, "fallback (MEASURED_PEAKS.json absent)"
    mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(mp):
        with open(mp) as fh:
            hbm_peak, hbm_src = float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy bandwidth, burst)"

    def hbm_entry(kernel, prof, what):
        ms, n, byts = prof
        if n == 0 or ms <= 0:
            return None
        gbs = byts / (ms * 1e-3) * 1e-9
        return {"kernel": kernel, "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": gbs / hbm_peak, "scopes": n, "avg_ms": ms / n, "algorithmic_bytes_per_scope": byts / n,
                "what": what}

    others = [
        hbm_entry("k_delta_build<sampled>", prof_setup["delta"],
                  f"setup: {n_epi_global} scenarios per epigraph drawn on the device, 8 s bytes out per scenario"),
        hbm_entry("k_delta_build<values>", prof_bulk["delta"],
                  f"{n_bulk} scenarios from realised values resident in HBM, 16 s bytes per scenario"),
        hbm_entry("k_cut_partial + k_sum_groups", prof_warm["reduce"],
                  "warm-up steps after the first: N (idx + weight + winning dot) + (rho, tau) table + cut per point"),
        hbm_entry("k_pool_push", prof_warm["pool"],
                  "warm-up steps after the first: hash scan 8 K + vector in/out per push (2E pushes per scope; "
                  "latency bound)"),
        hbm_entry("k_base + k_bias", prof_warm["bias"],
                  "warm-up steps after the first: one pass over the K x m2 pool for both points, once per cell "
                  "(epigraphs with the same template share it)"),
    ]
    others = [o for o in others if o]
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": leg["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic scenarios (counter RNG over the real storm outcome tables) on the real storm template; "
                    + ("REAL dual-vertex pool: harvested stage-2 LP duals" if kind == "real"
                       else "94 harvested + synthetic dual vertices"),
            "config": workload_config(args, z, world),
            "cut_formation_ms_per_sd_iter": leg["ms_per_step"],
            "roofline": rl["dominant"], "roofline_kernels": rl["all"], "roofline_other": others,
            "peak_hbm_source": hbm_src, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / max(1, args.steps)},
            "gpu_launches": launches, "setup_s": t_setup, "screening": leg["screen"], "sweep": leg["sweep"],
            "per_rank": leg["per_rank"],
            "parity_sample": parity}
    if other:
        line["other_pool"] = other
    if strong:
        line["strong"] = strong
    if not args.no_cpu_baseline and world == 1:
        base, _, _ = cpu_reference(z, pool_all[:K0], args.cpu_seconds, [x_c, x_i])
        base["single_thread"] = cpu_reference_single_thread(z, pool_all[:K0], min(4.0, args.cpu_seconds), [x_c, x_i])
        line["cpu_baseline"] = base
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def leg_summary(leg, kind, cell, traffic_shape=True):
    r = rooflines(leg, None, kind if traffic_shape else "-")
    return {"pool": kind, "ms_per_step": leg["ms_per_step"], "value": leg["value"], "unit": UNIT,
            "gpu_launches": leg["launches"], "screening": leg["screen"], "sweep": leg["sweep"], "per_rank": leg["per_rank"],
            "roofline": r["dominant"],
            "roofline_kernels": r["all"],
            "share_of_step": r["dominant"]["share_of_step"] if r["dominant"] else None, "clocks": leg["clocks"]}


def rooflines(leg, args, kind=None):
    """One entry per argmax kernel class that ran in the timed steps; `dominant` = the one with the most device time.
    achieved = work of the launches (flops counted with the pool size each launch saw) / their CUDA-event time."""
    prof, ms_dev = leg["prof"], leg["ms"]
    fp64_peak, fp64_src = None, "unmeasured"
    if os.path.exists(FP64_PEAK_FILE):
        with open(FP64_PEAK_FILE) as fh:
            fp64_peak = json.load(fh)["dfma"]["sustained_tflops"]
        fp64_src = "measured DFMA-chain microkernel, sustained (profiles/fp64_peak.json; MEASURED_PEAKS.json has no FP64 figure)"
    bf16_peak, bf16_src, bf16_burst = 1392.7, "fallback", None
    mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(mp):
        with open(mp) as fh:
            mpj = json.load(fh)
        bf16_peak, bf16_src = float(mpj["bf16_tflops_sustained"]), "MEASURED_PEAKS.json bf16_tflops_sustained (cuBLAS, inside a long run)"
        bf16_burst = float(mpj.get("bf16_tflops", 0.0)) or None
    traffic = {}
    tfile = os.path.join(ROOT, "profiles", "contract_traffic.json")
    if os.path.exists(tfile):   # dram bytes of one launch at the bench shape, from the committed ncu captures
        with open(tfile) as fh:
            traffic = json.load(fh)
    out = []

    def entry(cls, kernel, pipe, peak, peak_src, note, tkey):
        ms, n, work = prof[cls]
        if n == 0 or ms <= 0 or work <= 0:
            return
        tf = work / (ms * 1e-3) * 1e-12
        # (captured at the default bench shape only: storm, K = 16 384, 250k scenarios per epigraph)
        at_shape = args is None or (args.instance == "storm" and args.vertices == 16384 and args.scen_per_gpu == 1_000_000 and args.epigraphs == 4)
        tj = traffic.get(f"{tkey}_{kind or pool_kind(args)}" if tkey == "screen" else tkey) if at_shape else None
        tj = tj if isinstance(tj, dict) else None
        out.append({"bound": "tensor", "pipe": pipe, "kernel": kernel, "achieved": tf, "peak": peak, "unit": "TFLOP/s",
                    "frac": tf / peak if peak else None,
                    "traffic": (tj["dram_bytes_read"] + tj["dram_bytes_write"]) if tj else None,
                    "traffic_source": tj["source"] if tj else None, "peak_source": peak_src, "launches": n,
                    "avg_launch_ms": ms / n, "share_of_step": ms / ms_dev if ms_dev else None, "note": note})
        if cls == "screen" and bf16_burst:
            # the kernel runs for a few milliseconds inside a step with other kernels: between the two measured figures
            out[-1]["peak_burst"] = bf16_burst
            out[-1]["frac_of_burst"] = tf / bf16_burst

    entry("contract", "k_contract_ws<NX=2> + k_argmax_fixup", "fp64 tensor (DMMA, mma.sync.m8n8k4.f64; tcgen05 has no fp64)",
          fp64_peak, fp64_src,
          "flops = 2*s*K*N per launch with the K the launch saw, counted once although the launch serves both points "
          "(candidate and incumbent share the contraction); s = algorithmic rows (117 at storm, 120 executed)", "fp64")
    entry("screen", "k_screen<NX=2> (tcgen05.mma kind::f16, bf16 x 2 operands, fp32 accumulators in TMEM)",
          "5th-generation tensor cores, bf16", bf16_peak, bf16_src,
          "EXECUTED flops = 2 * 3 products * sp * K_pad * N_pad per launch (sp = rows padded to 16, K to 256, N to 128); "
          "algorithmically the launch stands for 2*s*K*N fp64 flops of the sweep it replaces", "screen")
    entry("fallback", "k_contract_ws behind a screening pass (ran because the pass fell back)", "fp64 tensor (DMMA)",
          fp64_peak, fp64_src, "time only: launches that found the pass successful exit at once", "fp64")
    res = prof["resolve"]
    extra = None
    if res[1] > 0:
        extra = {"kernel": "k_screen_resolve", "launches": res[1], "avg_launch_ms": res[0] / res[1],
                 "share_of_step": res[0] / ms_dev if ms_dev else None,
                 "what": "exact FP64 (DMMA chain) scores of the candidates, first-index maximum"}
    dom = max(out, key=lambda r: r["share_of_step"] or 0.0) if out else None
    return {"dominant": dom, "all": out + ([extra] if extra else [])}


def parity_sample(args, T, dist, cell, world, rank, alpha, beta, x_c, x_i, pool_all):
    """A sample of the bench's own final state against the oracle (reference subprob.jl:148-166,
    epigraph.jl:134-143): argmax of `parity-n` local scenarios at full K under the north-star rule, and the cuts of
    epigraph 0 against the oracle's accumulation over ALL its scenarios on the device's selection."""
    from oracle import oracle as O
    z, E = cell.z, cell.E
    K = len(cell.dvs)
    pool = pool_all[:K]                      # every push of a new vertex was inserted (asserted), in this order
    P = O.Problem(int(z["m2"]), int(z["n1"]), z["rbar"], z["T_colptr"], z["T_rowval"], z["T_nzval"], z["pos_row"], z["pos_col"])
    rng = np.random.default_rng(12345 + rank)
    res = {"n": 0, "mismatch": 0, "exempt": 0, "max_rel_val_err": 0.0, "max_rel_cut_err": None, "K": K}
    per = max(1, args.parity_n // (2 * E))
    mi_epi0 = {}
    for e, epi in enumerate(cell.epis):
        n_glob, n_loc, _ = epi.counts()
        for xi_, x in enumerate((x_c, x_i)):
            mv, mi = epi.argmax(x)
            if e == 0:
                mi_epi0[xi_] = (mv, mi)
            loc = np.sort(rng.choice(n_loc, size=min(per, n_loc), replace=False))
            glob = (loc // 128 * world + rank) * 128 + loc % 128           # rank r holds blocks r, r + world, ...
            vals = np.concatenate([sample_values(z, 101 + e, int(g), 1) for g in glob]) if len(glob) else np.zeros((0, P.s))
            ov, oi, _ = O.bench_argmax(P, vals, x, pool, threads=0, dot_kind=0)
            for j in np.nonzero(oi != mi[loc])[0]:
                sc, _ = O.score_pair(P, vals[j], x, pool[mi[loc][j]])
                if ov[j] - sc <= 1e-12 * max(abs(ov[j]), 1.0):
                    res["exempt"] += 1
                else:
                    res["mismatch"] += 1
            res["n"] += len(loc)
            if len(loc):
                res["max_rel_val_err"] = max(res["max_rel_val_err"],
                                             float(np.max(np.abs(mv[loc] - ov) / np.maximum(np.abs(ov), 1.0))))
    # cuts of epigraph 0 on the device's selection, all scenarios (every rank contributes its indices)
    epi = cell.epis[0]
    n_glob, n_loc, W = epi.counts()
    idx_glob = {}
    for xi_ in (0, 1):
        mi = mi_epi0[xi_][1]
        if world > 1:
            import torch
            cap = (n_glob // 128 // world + 2) * 128
            mine = torch.full((cap,), -1, dtype=torch.int64, device="cuda")
            mine[:n_loc] = torch.from_numpy(mi).cuda()
            allr = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allr, mine)
            full = np.full(n_glob, -1, dtype=np.int64)
            for r in range(world):
                loc = np.arange(cap)
                g = (loc // 128 * world + r) * 128 + loc % 128
                ok = g < n_glob
                a = allr[r].cpu().numpy()
                full[g[ok]] = a[ok]
            idx_glob[xi_] = full
        else:
            idx_glob[xi_] = mi
    if rank == 0 and n_glob > 1_500_000:
        res["cut_check"] = f"skipped: {n_glob} scenarios in epigraph 0 (the host copy of their values would not fit)"
    elif rank == 0:
        g = np.arange(n_glob, dtype=np.uint64)
        vals = np.empty((n_glob, P.s))
        for a in range(0, n_glob, 32768):
            vals[a:a + 32768] = sample_values(z, 101, a, min(32768, n_glob - a))
        w = 0.5 + u01(201, g)
        w[cell.n_epi_global:] = 1.0          # the scenarios the SD iterations appended carry weight 1 (algorithm.jl:46)
        worst = 0.0
        for xi_, x in enumerate((x_c, x_i)):
            ref = O.build_sasa_cut(P, vals, w, x, pool, forced_idx=idx_glob[xi_])
            p = w / ref["weight_mark"]
            sel = pool[idx_glob[xi_]]
            r_i = np.tile(P.rbar, (n_glob, 1))
            r_i[:, P.pos_row] = vals
            sa = np.sum(p * np.abs(np.einsum("ij,ij->i", sel, r_i))) + 1e-300
            sb = (p[:, None] * np.abs(sel @ P.T_dense())).sum(axis=0) + 1e-300
            worst = max(worst, abs(alpha[0, xi_] - ref["alpha"]) / max(sa, abs(ref["alpha"])),
                        float(np.max(np.abs(beta[0, xi_] - ref["beta"]) / np.maximum(sb, np.abs(ref["beta"])))))
        res["max_rel_cut_err"] = worst
        res["cut_check"] = f"epigraph 0, both points, all {n_glob} scenarios, oracle accumulation on the device's selection"
    res["rule"] = "index identical, or oracle score of the device's pick within 1e-12 relative of the oracle's best (exempt)"
    if world > 1:
        import torch
        t = torch.tensor([res["n"], res["mismatch"], res["exempt"]], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        res["n"], res["mismatch"], res["exempt"] = (int(v) for v in t.tolist())
    assert res["mismatch"] == 0, res
    assert res["max_rel_cut_err"] is None or res["max_rel_cut_err"] <= 1e-10, res
    return res


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
