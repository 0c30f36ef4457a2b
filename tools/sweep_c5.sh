#!/bin/bash
# BASELINE.json config C5 on one GPU: synthetic storm-shaped template (s = 128), K x N grid.
# -> gpurun_out/sweep_c5.jsonl and a markdown table on stdout
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; : > gpurun_out/sweep_c5.jsonl
for K in ${KS:-1024 4096 16384 65536}; do
  for N in ${NS:-10000 100000 1000000 10000000}; do
    steps=5; [ $((K * N / 1000000)) -gt 20000 ] && steps=2; [ $((K * N / 1000000)) -gt 200000 ] && steps=1
    python bench.py --no-cpu-baseline --instance synth128 --vertices $K --scen-per-gpu $N --epigraphs 1 \
        --steps $steps --warmup 3 2>>gpurun_out/sweep_c5.err | tail -1 >> gpurun_out/sweep_c5.jsonl
  done
done
python - <<'PY'
import json
print("| K | N | ms / SD iteration | evals/s | contraction TFLOP/s | of FP64 peak | e2e evals/s |")
print("|---|---|---|---|---|---|---|")
for l in open("gpurun_out/sweep_c5.jsonl"):
    try: d=json.loads(l)
    except Exception: continue
    c=d["config"]; r=d["roofline"]
    print(f'| {c["K_vertices"]} | {c["N_scenarios_per_gpu"]} | {d["ms_per_step"]:.3f} | {d["value"]:.3e} | {r["achieved"]:.2f} | {r["frac"]:.3f} | {d["e2e"]["value"]:.3e} |')
PY
