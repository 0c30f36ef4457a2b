#!/bin/bash
# GPU suite with programmatic dependent launch on (default) and off, then the small shapes both ways and
# the default bench line.  Writes gpurun_out/ab_*.log and gpurun_out/ab_shapes.jsonl.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; : > gpurun_out/ab_shapes.jsonl
rc=0
t0=$SECONDS
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/ab_tests_pdl1.log 2>&1; r=$?
echo "tests pdl=1: rc=$r $(tail -1 gpurun_out/ab_tests_pdl1.log) [$((SECONDS-t0)) s]"; [ $r -ne 0 ] && rc=$r
t0=$SECONDS
SQLP_PDL=0 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cuts.py tests/test_gpu_smps.py -m gpu -x -q > gpurun_out/ab_tests_pdl0.log 2>&1; r=$?
echo "tests pdl=0: rc=$r $(tail -1 gpurun_out/ab_tests_pdl0.log) [$((SECONDS-t0)) s]"; [ $r -ne 0 ] && rc=$r
run() { pdl=$1; shift; SQLP_PDL=$pdl python bench.py --no-cpu-baseline "$@" 2>>gpurun_out/ab_shapes.err | tail -1 | sed "s/^{/{\"pdl\": $pdl, /" >> gpurun_out/ab_shapes.jsonl; }
for pdl in 0 1; do
  run $pdl --instance baa99-20 --vertices 1024 --scen-per-gpu 10000 --epigraphs 1 --steps 50 --warmup 10
  run $pdl --instance synth128 --vertices 1024 --scen-per-gpu 10000 --epigraphs 1 --steps 50 --warmup 10
  run $pdl --instance ssn --vertices 5000 --scen-per-gpu 100000 --epigraphs 1 --steps 10 --warmup 3
  run $pdl --instance storm --vertices 16384 --scen-per-gpu 125000 --epigraphs 4 --steps 10 --warmup 3
done
python bench.py > gpurun_out/ab_bench.json 2> gpurun_out/ab_bench.err; r=$?
echo "bench: rc=$r"; [ $r -ne 0 ] && rc=$r
python - <<'PY'
import json
for l in open("gpurun_out/ab_shapes.jsonl"):
    try: d = json.loads(l)
    except Exception: print("bad line", l[:200]); continue
    c = d["config"]; r = d["roofline"]
    print(f'pdl={d["pdl"]} {c["instance"]:9s} K={c["K_vertices"]:6d} N={c["N_scenarios_per_gpu"]:8d} ms/iter={d["ms_per_step"]:.4f} '
          f'e2e_ms={d["e2e"]["ms_per_step"]:.4f} launches={d["gpu_launches"]} contraction={r["avg_launch_ms"]:.4f} ms frac={r["frac"]:.3f}')
PY
tail -c 600 gpurun_out/ab_bench.json
exit $rc
