#!/bin/bash
# One GPU: screening tests, device legs on both pools, then ncu of the chain (launch list + full capture).
mkdir -p gpurun_out
T=${1:-v}
(timeout 900 python -m pytest tests/test_gpu_screen.py tests/test_gpu_guards.py -m gpu -q -x 2>&1 | tail -6) > gpurun_out/r02${T}_tests.log 2>&1
: > gpurun_out/r02${T}_legs.jsonl
run() { echo "# $*" >> gpurun_out/r02${T}_legs.err; echo "# $*" >> gpurun_out/r02${T}_legs.jsonl; env "$1" timeout 300 python bench.py --dev-only --no-cpu-baseline --no-extra-legs "${@:2}" 2>>gpurun_out/r02${T}_legs.err | grep '^{' | tail -1 >> gpurun_out/r02${T}_legs.jsonl; }
run A=1 --pool real
run A=1 --pool synthetic
export SQLP_BENCH_CUPROF=1
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02${T}_launches_real.csv python bench.py --dev-only --pool real --steps 2 --warmup 3 > gpurun_out/r02${T}_ncu_real.log 2>&1
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_screen" -c 10 -o gpurun_out/r02${T}_full_real python bench.py --dev-only --pool real --steps 1 --warmup 3 > gpurun_out/r02${T}_ncu_full.log 2>&1
tail -n 4 gpurun_out/r02${T}_tests.log; python - <<PY
import json
for l in open('gpurun_out/r02${T}_legs.jsonl'):
    if l.startswith('#'): print(l.strip()); continue
    try:
        j=json.loads(l); print('   ms', round(j.get('ms_per_step',0),3), 'screen', j.get('screen'), 'prof', json.dumps(j.get('prof'))[:600])
    except Exception as e: print('   ?', l[:200])
PY
