#!/bin/bash
# GPU suite, then the small (latency-bound) shapes and the default bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; : > gpurun_out/small_shapes.jsonl
rc=0
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/small_tests.log 2>&1; r=$?
echo "tests: rc=$r $(tail -1 gpurun_out/small_tests.log)"; [ $r -ne 0 ] && rc=$r
run() { python bench.py --no-cpu-baseline "$@" 2>>gpurun_out/small_shapes.err | tail -1 >> gpurun_out/small_shapes.jsonl; }
SQLP_PDL=0 run --instance baa99-20 --vertices 1024 --scen-per-gpu 10000 --epigraphs 1 --steps 100 --warmup 10
SQLP_PDL=0 run --instance ssn --vertices 3000 --scen-per-gpu 3000 --epigraphs 1 --steps 100 --warmup 10
run --instance baa99-20 --vertices 1024 --scen-per-gpu 10000 --epigraphs 1 --steps 100 --warmup 10
run --instance ssn --vertices 3000 --scen-per-gpu 3000 --epigraphs 1 --steps 100 --warmup 10
run --instance synth128 --vertices 1024 --scen-per-gpu 10000 --epigraphs 1 --steps 50 --warmup 10
run --instance storm --vertices 16384 --scen-per-gpu 125000 --epigraphs 4 --steps 10 --warmup 3
if [ "$1" = "full" ]; then python bench.py > gpurun_out/small_bench.json 2> gpurun_out/small_bench.err; echo "bench rc=$?"; fi
python - <<'PY'
import json
for l in open("gpurun_out/small_shapes.jsonl"):
    try: d = json.loads(l)
    except Exception: print("bad line", l[:200]); continue
    c = d["config"]; r = d["roofline"]
    print(f'{c["instance"]:9s} K={c["K_vertices"]:6d} N={c["N_scenarios_per_gpu"]:8d} ms/iter={d["ms_per_step"]:.4f} '
          f'e2e_ms={d["e2e"]["ms_per_step"]:.4f} launches={d["gpu_launches"]} contraction={r["avg_launch_ms"]:.4f} ms frac={r["frac"]:.3f}')
    print("      " + "  ".join(f'{o["kernel"].split(" ")[0]}={o["avg_ms"]*1e3:.1f}us' for o in d["roofline_other"][2:]))
PY
exit $rc
