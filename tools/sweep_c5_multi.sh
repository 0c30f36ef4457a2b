#!/bin/bash
# BASELINE.json config C5 on G GPUs of one box: synthetic storm-shaped template (s = 128), K x N grid, N = scenarios
# IN TOTAL (sharded over the GPUs), plus the delta-T variant (8 random Tbar entries per scenario) at one point.
# usage: tools/sweep_c5_multi.sh G [tag]   -> gpurun_out/sweep_c5_g<G>.jsonl and a markdown table on stdout
cd "$(dirname "$0")/.."
G=${1:-1}; TAG=${2:-sweep_c5_g$G}
mkdir -p gpurun_out; : > gpurun_out/$TAG.jsonl
run() {
  if [ "$G" -gt 1 ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $G "$@" 2>>gpurun_out/$TAG.err | grep '^{' | tail -1 >> gpurun_out/$TAG.jsonl
  else
    python bench.py "$@" 2>>gpurun_out/$TAG.err | grep '^{' | tail -1 >> gpurun_out/$TAG.jsonl
  fi
}
for K in ${KS:-1024 16384 65536}; do
  for N in ${NS:-10000 1000000 10000000}; do
    steps=5; [ $((K * N / G / 1000000)) -gt 20000 ] && steps=2
    run --no-cpu-baseline --no-extra-legs --instance synth128 --vertices $K --scen-per-gpu $((N / G)) --epigraphs 1 --steps $steps --warmup 3
  done
done
run --no-cpu-baseline --no-extra-legs --instance synth128T8 --vertices 16384 --scen-per-gpu $((1000000 / G)) --epigraphs 1 --steps 3 --warmup 3
python - "$TAG" <<'PY'
import json, sys
print("| instance | K | N total | GPUs | ms / SD iteration | evals/s | dominant kernel | achieved | of its peak | exact evaluations per scenario-point | parity sample (n / mismatch / exempt) | SM MHz | throttle |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for l in open(f"gpurun_out/{sys.argv[1]}.jsonl"):
    try: d = json.loads(l)
    except Exception: continue
    c, r, sc, ps, ck = d["config"], d["roofline"] or {}, d.get("screening", {}), d.get("parity_sample", {}), d["clocks"]
    kern = (r.get("kernel") or "-").split("(")[0].split("+")[0].strip()
    ev = sc.get("evaluated", 0) / max(1, c["N_scenarios_per_gpu"] * c["epigraphs"] * 2) if sc.get("passes") else None
    print(f'| {c["instance"]} | {c["K_vertices"]} | {c["N_scenarios_total"]} | {d["n_gpus"]} | {d["ms_per_step"]:.3f} | {d["value"]:.3e} | {kern} | '
          f'{(r.get("achieved") or 0):.1f} {r.get("unit", "")} | {(r.get("frac") or 0):.3f} | {"-" if ev is None else f"{ev:.2f}"} | '
          f'{ps.get("n")} / {ps.get("mismatch")} / {ps.get("exempt")} | {ck.get("sm_mhz")} | {",".join(ck.get("reasons") or []) or "none"} |')
PY
