#!/bin/bash
# Final one-GPU session of round 2: the GPU suite, smoke(), the default bench (real pool) and the synthetic-pool bench,
# then the ncu evidence of the same commands (launch lists of both pools, one full capture of the step's kernels).
mkdir -p gpurun_out
T=${1:-fin}
(timeout 1100 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/r02${T}_tests.log 2>&1
(timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3) > gpurun_out/r02${T}_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/r02${T}_bench_real.json 2> gpurun_out/r02${T}_bench_real.err
timeout 400 python bench.py --pool synthetic --no-extra-legs > gpurun_out/r02${T}_bench_synth.json 2> gpurun_out/r02${T}_bench_synth.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02${T}_bench_reference.json 2> gpurun_out/r02${T}_bench_reference.err
export SQLP_BENCH_CUPROF=1
for P in real synthetic; do
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02${T}_launches_${P}.csv python bench.py --dev-only --pool $P --steps 2 --warmup 3 > gpurun_out/r02${T}_ncu_${P}.log 2>&1
done
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_screen|k_cut_hist|k_cut_fold|k_bias|k_sum_groups|k_pool_push" -c 16 -o gpurun_out/r02${T}_full_real python bench.py --dev-only --pool real --steps 1 --warmup 3 > gpurun_out/r02${T}_ncu_full.log 2>&1
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_screen<|k_screen_decide" -c 2 -o gpurun_out/r02${T}_full_synth python bench.py --dev-only --pool synthetic --steps 1 --warmup 3 > gpurun_out/r02${T}_ncu_full_synth.log 2>&1
tail -n 3 gpurun_out/r02${T}_tests.log; cat gpurun_out/r02${T}_smoke.log; tail -c 200 gpurun_out/r02${T}_bench_real.err; head -c 250 gpurun_out/r02${T}_bench_real.json; echo; head -c 250 gpurun_out/r02${T}_bench_synth.json; echo; head -c 400 gpurun_out/r02${T}_bench_reference.json; ls -la gpurun_out | grep r02${T}
