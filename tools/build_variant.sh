#!/bin/bash
# Dev helper: build an experimental libdvo_b200 variant:  tools/build_variant.sh <name> [extra nvcc flags...]
# -> build/exp/libdvo_<name>.so (+ ptxas -v log).  Uses -DDVO_FAST_BUILD (headline kernel variants only).
set -e
name=$1; shift
fast=-DDVO_FAST_BUILD; [ -n "$FULL" ] && fast=   # FULL=1: every kernel variant
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/build/exp/$name
mkdir -p $out
src=$root/dense-visual-odometry_b200/csrc
pids=()
for u in $src/*.cu; do
  b=$(basename $u .cu)
  nvcc -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 $fast "$@" \
       -Xptxas -v -c -o $out/$b.o $u > $out/$b.log 2>&1 &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -o $root/build/exp/libdvo_$name.so $out/*.o
grep -A1 "align_kernelILi0ELi0ELi0ELi128ELi2ELi0" $out/variants_128_g0.log | grep -E "registers|spill" | head -3
