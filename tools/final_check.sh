#!/bin/bash
# What the driver runs at round end, in one GPU call: the GPU suite, smoke(), the default bench line and the
# reference arm (short).  Writes gpurun_out/final_*.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rc=0
python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; r=$?; echo "tests rc=$r $(tail -1 gpurun_out/final_tests.log)"; [ $r -ne 0 ] && rc=$r
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; r=$?; echo "smoke rc=$r $(tail -1 gpurun_out/final_smoke.log)"; [ $r -ne 0 ] && rc=$r
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; r=$?; echo "bench rc=$r"; [ $r -ne 0 ] && rc=$r
python -c "
import json
d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], 'cpu', d['cpu_baseline']['value'], d['clocks'])
"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err; r=$?; echo "reference arm rc=$r $(tail -c 400 gpurun_out/final_reference.json)"; [ $r -ne 0 ] && rc=$r
exit $rc
