#!/bin/bash
# Fourth GPU session of round 2 (one GPU): warm start of the screening scan (k_screen_seed), points that move
# every step in the bench.
mkdir -p gpurun_out
T=${1:-t}
(timeout 900 python -m pytest tests/test_gpu_screen.py tests/test_gpu_guards.py tests/test_gpu_sizes.py -m gpu -q -x -s 2>&1 | grep -v "^$" | tail -12) > gpurun_out/r02${T}_tests.log 2>&1
: > gpurun_out/r02${T}_legs.jsonl
run() { echo "# $*" >> gpurun_out/r02${T}_legs.err; echo "# $*" >> gpurun_out/r02${T}_legs.jsonl; env "$1" timeout 300 python bench.py --dev-only --no-cpu-baseline --no-extra-legs "${@:2}" 2>>gpurun_out/r02${T}_legs.err | grep '^{' | tail -1 >> gpurun_out/r02${T}_legs.jsonl; }
run A=1 --pool real
run SQLP_RESOLVE=lanes --pool real
run A=1 --pool real --screen 0
run A=1 --pool synthetic
run SQLP_RESOLVE=lanes --pool synthetic
timeout 600 python bench.py > gpurun_out/r02${T}_bench_real.json 2> gpurun_out/r02${T}_bench_real.err
tail -n 6 gpurun_out/r02${T}_tests.log; tail -c 300 gpurun_out/r02${T}_bench_real.err; python - <<PY
import json
for l in open('gpurun_out/r02${T}_legs.jsonl'):
    if l.startswith('#'): print(l.strip()); continue
    try:
        j=json.loads(l); print('   ms', round(j.get('ms_per_step',0),3), 'screen', j.get('screen'), 'prof', json.dumps(j.get('prof'))[:600])
    except Exception as e: print('   ?', l[:200])
j=json.load(open('gpurun_out/r02${T}_bench_real.json'))
print(j['ms_per_step'], j['e2e'], j['parity_sample'])
PY
