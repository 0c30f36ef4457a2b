#!/bin/bash
# Round evidence in one GPU call: the default bench line, the shapes report, the ncu launch list of the same
# bench command (after it exited 0 without ncu) and one `ncu --set full` capture of the step's kernels.
# usage: bash tools/profile_round.sh <tag>      -> gpurun_out/<tag>_*
cd "$(dirname "$0")/.."
tag=${1:-rXX}
mkdir -p gpurun_out
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || { echo "bench failed"; tail -5 gpurun_out/${tag}_bench.err; exit 1; }
echo "bench ok: $(python -c "import json;d=json.load(open('gpurun_out/${tag}_bench.json'));print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])")"
bash tools/shapes_report.sh > gpurun_out/${tag}_shapes.md 2> gpurun_out/${tag}_shapes.err; cp gpurun_out/shapes.jsonl gpurun_out/${tag}_shapes.jsonl
cat gpurun_out/${tag}_shapes.md
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_b2.log 2>&1 || { echo "bench (2 steps) failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 32768 -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_list.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -s 32768 -c 14 -o gpurun_out/${tag}_step \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --epigraphs 1 --scen-per-gpu 250000 > gpurun_out/${tag}_ncu_full.log 2>&1; echo "full capture rc=$?"
ls -la gpurun_out/ | grep ${tag}
