#!/bin/bash
# One GPU session of round 2: probe, the GPU test suite, both bench pools, launch lists and ncu captures.
mkdir -p gpurun_out
T=${1:-g}
(timeout 200 tools/tc_probe 16384 250000 117 0 0) > gpurun_out/r02${T}_probe.log 2>&1
(timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | tail -40) > gpurun_out/r02${T}_tests.log 2>&1
timeout 600 python bench.py > gpurun_out/r02${T}_bench_real.json 2> gpurun_out/r02${T}_bench_real.err
timeout 300 python bench.py --pool synthetic --no-extra-legs --no-cpu-baseline > gpurun_out/r02${T}_bench_synth.json 2> gpurun_out/r02${T}_bench_synth.err
export SQLP_BENCH_CUPROF=1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02${T}_launches_real.csv python bench.py --dev-only --steps 2 --warmup 2 > gpurun_out/r02${T}_ncu_real.log 2>&1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02${T}_launches_synth.csv python bench.py --dev-only --pool synthetic --steps 2 --warmup 2 > gpurun_out/r02${T}_ncu_synth.log 2>&1
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_screen|k_contract_ws|k_cut_hist|k_cut_fold|k_bias" -c 12 -o gpurun_out/r02${T}_full python bench.py --dev-only --steps 1 --warmup 2 > gpurun_out/r02${T}_ncu_full.log 2>&1
tail -n 4 gpurun_out/r02${T}_probe.log; tail -n 4 gpurun_out/r02${T}_tests.log; tail -c 300 gpurun_out/r02${T}_bench_real.err; tail -c 300 gpurun_out/r02${T}_bench_synth.err; ls -la gpurun_out | grep r02${T}
