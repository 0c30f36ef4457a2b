#!/bin/bash
# Third GPU session of round 2 (one GPU): centred screening operands + row-major view for the exact decision.
# Tests that pin bit-identity first, then the device leg of the bench with the pass forced / automatic / off on both
# pools, with and without centring, and the programmatic-dependent-launch A/B of the decision kernel.
mkdir -p gpurun_out
T=${1:-p}
(timeout 900 python -m pytest tests/test_gpu_screen.py tests/test_gpu_guards.py -m gpu -q -x 2>&1 | tail -8) > gpurun_out/r02${T}_tests.log 2>&1
: > gpurun_out/r02${T}_legs.jsonl
run() { echo "# $*" >> gpurun_out/r02${T}_legs.err; echo "# $*" >> gpurun_out/r02${T}_legs.jsonl; env "$1" timeout 300 python bench.py --dev-only --no-cpu-baseline --no-extra-legs "${@:2}" 2>>gpurun_out/r02${T}_legs.err | grep '^{' | tail -1 >> gpurun_out/r02${T}_legs.jsonl; }
run A=1 --pool real
run A=1 --pool real --screen 2
run SQLP_CENTRE=0 --pool real --screen 2
run A=1 --pool real --screen 0
run A=1 --pool synthetic
run SQLP_CENTRE=0 --pool synthetic
run SQLP_PDL=0 --pool synthetic
run SQLP_PDL=0 --pool real --screen 2
timeout 600 python bench.py > gpurun_out/r02${T}_bench_real.json 2> gpurun_out/r02${T}_bench_real.err
tail -n 4 gpurun_out/r02${T}_tests.log; tail -c 300 gpurun_out/r02${T}_bench_real.err; python - <<'PY'
import json,sys,glob
T=sys.argv[1] if len(sys.argv)>1 else 'p'
for l in open(glob.glob('gpurun_out/r02*_legs.jsonl')[-1]):
    if l.startswith('#'): print(l.strip()); continue
    try:
        j=json.loads(l); print('   ms', round(j.get('ms_per_step',0),3), 'screen', j.get('screen'), 'prof', json.dumps(j.get('prof'))[:600])
    except Exception as e: print('   ?', l[:200])
PY
