#!/bin/bash
# run bench for each (library, stagger, prefetch) combination
cd "$(dirname "$0")/.."
for so in build/variants/*.so; do
 for st in ${STAGGERS:-0 500 1000 2000 4000}; do
  for pf in ${PREFETCHES:-0}; do
    out=$(SQLP_B200_LIB=$PWD/$so SQLP_CONTRACT_STAGGER_NS=$st SQLP_CONTRACT_PREFETCH=$pf timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline ${SWEEP_ARGS:-} 2>&1 | tail -1)
    echo "$so stagger=$st prefetch=$pf: $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print("achieved %.2f TF frac %.3f avg_launch_ms %.2f ms_per_step %.2f" % (r["achieved"], r["frac"] or 0, r["avg_launch_ms"], d["ms_per_step"]))' 2>&1 | tail -1)"
  done
 done
done
