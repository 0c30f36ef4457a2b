#!/bin/bash
# One GPU, quick: the screening / guard / size tests and the device leg of the bench on both pools.
mkdir -p gpurun_out
T=${1:-w}
(timeout 900 python -m pytest tests/test_gpu_screen.py tests/test_gpu_guards.py tests/test_gpu_sizes.py -m gpu -q -x 2>&1 | tail -6) > gpurun_out/r02${T}_tests.log 2>&1
: > gpurun_out/r02${T}_legs.jsonl
run() { echo "# $*" >> gpurun_out/r02${T}_legs.err; echo "# $*" >> gpurun_out/r02${T}_legs.jsonl; env "$1" timeout 300 python bench.py --dev-only --no-cpu-baseline --no-extra-legs "${@:2}" 2>>gpurun_out/r02${T}_legs.err | grep '^{' | tail -1 >> gpurun_out/r02${T}_legs.jsonl; }
run A=1 --pool real
run A=1 --pool synthetic
for extra in "$@"; do [ "$extra" = "$T" ] || run $extra; done
tail -n 4 gpurun_out/r02${T}_tests.log; python - <<PY
import json
for l in open('gpurun_out/r02${T}_legs.jsonl'):
    if l.startswith('#'): print(l.strip()); continue
    try:
        j=json.loads(l); print('   ms', round(j.get('ms_per_step',0),3), 'screen', j.get('screen'), 'prof', json.dumps(j.get('prof'))[:600])
    except Exception as e: print('   ?', l[:200])
PY
