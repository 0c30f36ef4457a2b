#!/bin/bash
# Second GPU session of round 2 (one GPU): the test suite, both bench pools, the named small / medium shapes on their
# real pools, a memory-checker attempt on the smallest cases, and full ncu captures of the screening chain.
mkdir -p gpurun_out
T=${1:-m}
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/r02${T}_tests.log 2>&1
timeout 600 python bench.py > gpurun_out/r02${T}_bench_real.json 2> gpurun_out/r02${T}_bench_real.err
timeout 300 python bench.py --pool synthetic --no-extra-legs --no-cpu-baseline > gpurun_out/r02${T}_bench_synth.json 2> gpurun_out/r02${T}_bench_synth.err
: > gpurun_out/r02${T}_shapes.jsonl
run() { echo "# $*" >> gpurun_out/r02${T}_shapes.err; timeout 300 python bench.py --no-cpu-baseline --no-extra-legs "$@" 2>>gpurun_out/r02${T}_shapes.err | grep '^{' | tail -1 >> gpurun_out/r02${T}_shapes.jsonl; }
run --instance baa99-20 --vertices 1024 --scen-per-gpu 10000 --epigraphs 1 --steps 20 --warmup 5
run --instance baa99-20 --vertices 1024 --scen-per-gpu 10000 --epigraphs 1 --steps 20 --warmup 5 --screen 0
run --instance ssn --vertices 5000 --scen-per-gpu 100000 --epigraphs 1 --steps 10 --warmup 3
run --instance ssn --vertices 5000 --scen-per-gpu 100000 --epigraphs 1 --steps 10 --warmup 3 --screen 0
run --instance storm --vertices 16384 --scen-per-gpu 125000 --epigraphs 4 --steps 10 --warmup 3
run --instance storm --vertices 16384 --scen-per-gpu 1000000 --epigraphs 4 --steps 5 --warmup 3 --screen 2
(timeout 200 compute-sanitizer --tool memcheck --print-limit 10 tools/tc_probe 1000 3000 20 0 0 2>&1 | tail -25) > gpurun_out/r02${T}_sanitizer_probe.log 2>&1
(SQLP_BENCH_NOCPU=1 timeout 400 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_screen.py -q -k "synthetic_shapes and 129" 2>&1 | tail -25) > gpurun_out/r02${T}_sanitizer_py.log 2>&1
export SQLP_BENCH_CUPROF=1
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_screen|k_cut_hist|k_cut_fold|k_bias" -c 8 -o gpurun_out/r02${T}_full_synth python bench.py --dev-only --pool synthetic --steps 1 --warmup 2 > gpurun_out/r02${T}_ncu_full.log 2>&1
tail -n 3 gpurun_out/r02${T}_tests.log; tail -c 300 gpurun_out/r02${T}_bench_real.err; tail -n 5 gpurun_out/r02${T}_sanitizer_probe.log; wc -l gpurun_out/r02${T}_shapes.jsonl; ls -la gpurun_out | grep r02${T}
