#!/bin/bash
# One GPU: the whole GPU suite, device legs (both pools; the per-GPU share of the 8-GPU strong-scaled job; the
# ordered FP64 histogram for comparison), the default bench.
mkdir -p gpurun_out
T=${1:-x}
(timeout 1100 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/r02${T}_tests.log 2>&1
: > gpurun_out/r02${T}_legs.jsonl
run() { echo "# $*" >> gpurun_out/r02${T}_legs.err; echo "# $*" >> gpurun_out/r02${T}_legs.jsonl; env "$1" timeout 300 python bench.py --dev-only --no-cpu-baseline --no-extra-legs "${@:2}" 2>>gpurun_out/r02${T}_legs.err | grep '^{' | tail -1 >> gpurun_out/r02${T}_legs.jsonl; }
run A=1 --pool real
run SQLP_HIST=float --pool real
run A=1 --pool synthetic
run A=1 --pool real --scen-per-gpu 125000
run A=1 --pool real --scen-per-gpu 250000
timeout 600 python bench.py > gpurun_out/r02${T}_bench_real.json 2> gpurun_out/r02${T}_bench_real.err
tail -n 5 gpurun_out/r02${T}_tests.log; tail -c 300 gpurun_out/r02${T}_bench_real.err; python - <<PY
import json
for l in open('gpurun_out/r02${T}_legs.jsonl'):
    if l.startswith('#'): print(l.strip()); continue
    try:
        j=json.loads(l); print('   ms', round(j.get('ms_per_step',0),3), 'screen', j.get('screen'), 'prof', json.dumps(j.get('prof'))[:600])
    except Exception as e: print('   ?', l[:200])
j=json.load(open('gpurun_out/r02${T}_bench_real.json'))
print(j['ms_per_step'], j['e2e'], j['parity_sample'])
PY
