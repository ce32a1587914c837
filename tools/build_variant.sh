#!/bin/bash
# Developer builds of the library with experiment switches: tools/build_variant.sh <name> "<-D flags>"
#   -> pope_b200/variants/libpope_b200_<name>.so, selected at run time with POPE_B200_LIB=<path> (pope_b200/_lib.py).
# Only coarse_tc.cu, coarse_finalize.cu and fine.cu are recompiled with the flags; the other objects are shared with the product build.
set -e
cd "$(dirname "$0")/../pope_b200/csrc"
name=$1; flags=$2
mkdir -p ../variants build_$name
for f in coarse_api coarse_simt fine_tf retrieval pipeline points_io pose; do
  [ -f build/$f.o ] || make build/$f.o >/dev/null
  cp -u build/$f.o build_$name/$f.o
done
rm -f build_$name/coarse_tc.o build_$name/coarse_finalize.o build_$name/fine.o
make -j4 BUILD=build_$name OUT=../variants/libpope_b200_$name.so EXTRA="$flags" >/dev/null
grep -A1 "sweep_tc_kernelILi3ELb0" build_$name/coarse_tc.ptxas.log | grep -E "spill|Used" | head -3
echo "built variants/libpope_b200_$name.so"
