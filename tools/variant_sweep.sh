#!/bin/bash
# Contraction variant sweep.  Step 1 (CPU box, no GPU needed):  tools/variant_sweep.sh build "WR MI KG CTAS" ...
# compiles one library per variant into build/variants/.  Step 2 (GPU box):  tools/variant_sweep.sh run
# times every library found there with bench.py (SQLP_B200_LIB selects the library).
set -u
cd "$(dirname "$0")/.."
mkdir -p build/variants
if [ "$1" = build ]; then
  shift
  for v in "$@"; do
    set -- $v
    out=build/variants/res_wr$1_mi$2_kg$3_c$4${5:+_$5}.so
    ( nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
        -DSQLP_RES_WR=$1 -DSQLP_RES_MI=$2 -DSQLP_RES_KG=$3 -DSQLP_RES_CTAS=$4 ${6:-} -Xptxas -v \
        -Xcompiler -fPIC -shared -o $out sqlp_b200/csrc/sqlp_api.cu -ldl 2>&1 \
        | grep -A2 "k_contract_residentINS_11ResidentCfgILi2" | grep -E "spill|Used" | tr '\n' ' '; echo " <- $out" ) &
  done
  wait
else
  for so in build/variants/*.so; do
    out=$(SQLP_B200_LIB=$PWD/$so timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline ${SWEEP_ARGS:-} 2>&1 | tail -1)
    echo "$so: $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print("achieved %.2f TF frac %.3f avg_launch_ms %.2f ms_per_step %.2f" % (r["achieved"], r["frac"] or 0, r["avg_launch_ms"], d["ms_per_step"]))' 2>&1 | tail -1)"
  done
fi
