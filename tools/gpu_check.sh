#!/bin/bash
# GPU parity suite under several contraction plans (default, forced small grids that cut units
# at many places, the streaming fallback), then one bench line.  Writes gpurun_out/check_*.log.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rc=0
run() { # name, env...
  name=$1; shift
  env "$@" python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/check_$name.log 2>&1
  r=$?; echo "$name: rc=$r $(tail -1 gpurun_out/check_$name.log)"; [ $r -ne 0 ] && rc=$r
}
run default SQLP_X=1
run grid7 SQLP_CONTRACT_GRID=7
run grid3 SQLP_CONTRACT_GRID=3
run grid50 SQLP_CONTRACT_GRID=50
run grid1000 SQLP_CONTRACT_GRID=1000
run stream SQLP_CONTRACT=stream
run resident SQLP_CONTRACT=resident
run resident_grid7 SQLP_CONTRACT=resident SQLP_CONTRACT_GRID=7
exit $rc
