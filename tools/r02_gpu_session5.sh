#!/bin/bash
# ncu evidence for the centred + warm-started screening chain (one GPU).  Each capture only after the same command
# exited 0 without ncu (session 4).
mkdir -p gpurun_out
T=${1:-s}
export SQLP_BENCH_CUPROF=1
for P in real synthetic; do
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02${T}_launches_${P}.csv python bench.py --dev-only --pool $P --steps 2 --warmup 3 > gpurun_out/r02${T}_ncu_${P}.log 2>&1
done
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_screen|k_cut_hist|k_cut_fold|k_bias" -c 14 -o gpurun_out/r02${T}_full_real python bench.py --dev-only --pool real --steps 1 --warmup 3 > gpurun_out/r02${T}_ncu_full.log 2>&1
ls -la gpurun_out | grep r02${T}
