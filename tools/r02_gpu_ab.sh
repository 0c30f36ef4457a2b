#!/bin/bash
# A/B on one box: device leg of the bench, real pool, alternating a switch (three rounds each).
mkdir -p gpurun_out
T=${1:-ab}; SW=${2:-SQLP_FADD2=0}
: > gpurun_out/r02${T}.txt
for round in 1 2 3; do
  for E in A=1 $SW; do
    env $E timeout 300 python bench.py --dev-only --no-cpu-baseline --no-extra-legs --pool real 2>/dev/null | grep '^{' | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); p=j['prof']; print('$E', round(j['ms_per_step'],3), 'screen', round(p['screen'][0]/p['screen'][1],4), 'resolve', round(p['resolve'][0]/p['resolve'][1],4))" >> gpurun_out/r02${T}.txt
  done
done
cat gpurun_out/r02${T}.txt
