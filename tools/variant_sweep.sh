#!/bin/bash
# Build and time contraction variants on the GPU box (MI STAGES PREFETCH CTAS per line).
# usage: tools/variant_sweep.sh "8 7 3 2" "4 5 2 3" ...
set -u
cd "$(dirname "$0")/.."
cp sqlp_b200/libsqlp_b200.so /tmp/libsqlp_keep.so
for v in "$@"; do
  set -- $v
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
    -DSQLP_VARIANT_MI=$1 -DSQLP_VARIANT_STAGES=$2 -DSQLP_VARIANT_PREFETCH=$3 -DSQLP_VARIANT_CTAS=$4 -DSQLP_VARIANT_KG=${5:-2} ${6:-} \
    -Xcompiler -fPIC -shared -o sqlp_b200/libsqlp_b200.so sqlp_b200/csrc/sqlp_api.cu -ldl || continue
  out=$(timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline ${SWEEP_ARGS:-} 2>&1 | tail -1)
  echo "variant MI=$1 S=$2 P=$3 CTAS=$4 KG=${5:-2} ${6:-}: $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print("achieved %.2f TF frac %.3f avg_launch_ms %.2f ms_per_step %.2f" % (r["achieved"], r["frac"] or 0, r["avg_launch_ms"], d["ms_per_step"]))' 2>&1 | tail -1)"
done
cp /tmp/libsqlp_keep.so sqlp_b200/libsqlp_b200.so
