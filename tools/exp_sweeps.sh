#!/bin/bash
# Developer experiment runner (GPU box): times the coarse stage for the product library and for every variant library under
# pope_b200/variants/ (tools/build_variant.sh), sigma 1 and sigma 3.06.  usage: tools/exp_sweeps.sh [variant ...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/exp_sweeps.log
: > $out
run() {  # label, env..., then args after --
  local label=$1; shift
  echo "== $label" >> $out
  env "$@" 2>&1 | tail -1 >> $out
}
run "product sigma1" python tools/time_sweeps.py
run "product sigma3" python tools/time_sweeps.py sigma 3.06
for d in ${EXP_DEBUG:-}; do
  run "product debug=$d sigma1" POPE_TC_DEBUG=$d python tools/time_sweeps.py
  run "product debug=$d sigma3" POPE_TC_DEBUG=$d python tools/time_sweeps.py sigma 3.06
done
for v in "$@"; do
  lib=$PWD/pope_b200/variants/libpope_b200_$v.so
  run "$v sigma1" POPE_B200_LIB=$lib python tools/time_sweeps.py
  run "$v sigma3" POPE_B200_LIB=$lib python tools/time_sweeps.py sigma 3.06
done
cat $out
