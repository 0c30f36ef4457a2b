#!/bin/bash
# One bench line per named shape of BASELINE.json (configs C2, C3, the per-GPU shard of C4, C4 on one GPU
# and three C5 points) -> gpurun_out/shapes.jsonl
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; : > gpurun_out/shapes.jsonl
run() { echo "# $*" >&2; python bench.py --no-cpu-baseline "$@" 2>>gpurun_out/shapes.err | tail -1 >> gpurun_out/shapes.jsonl; }
run --instance baa99-20 --vertices 1024 --scen-per-gpu 10000 --epigraphs 1 --steps 20 --warmup 5
run --instance ssn --vertices 5000 --scen-per-gpu 100000 --epigraphs 1 --steps 10 --warmup 3
run --instance storm --vertices 16384 --scen-per-gpu 125000 --epigraphs 4 --steps 10 --warmup 3
run --instance storm --vertices 16384 --scen-per-gpu 1000000 --epigraphs 4 --steps 5 --warmup 3
run --instance synth128 --vertices 1024 --scen-per-gpu 10000 --epigraphs 1 --steps 20 --warmup 5
run --instance synth128 --vertices 8192 --scen-per-gpu 1000000 --epigraphs 1 --steps 5 --warmup 3
run --instance synth128 --vertices 65536 --scen-per-gpu 1000000 --epigraphs 1 --steps 3 --warmup 3
python - <<'PY'
import json
print("| shape | s | K | N/GPU | E | ms/iter | evals/s | contraction TFLOP/s | of FP64 peak | e2e evals/s |")
print("|---|---|---|---|---|---|---|---|---|---|")
for l in open("gpurun_out/shapes.jsonl"):
    try: d=json.loads(l)
    except Exception: continue
    c=d["config"]; r=d["roofline"]
    print(f'| {c["instance"]} | {c["s"]} | {c["K_vertices"]} | {c["N_scenarios_per_gpu"]} | {c["epigraphs"]} | {d["ms_per_step"]:.3f} | {d["value"]:.3e} | {r["achieved"]:.2f} | {r["frac"]:.3f} | {d["e2e"]["value"]:.3e} |')
PY
