#!/bin/bash
# Builds tuning variants of libsuews_b200.so into scratch/variants/ (launch geometry, inlining switches); time them on
# the GPU box with tools/run_variants.sh (SUEWS_B200_LIB selects the library a process loads).
#   usage: tools/variants.sh "name:-DFLAG1 -DFLAG2" ...
set -e
cd "$(dirname "$0")/../suews_b200/csrc"
mkdir -p ../../scratch/variants
for v in "$@"; do
  name="${v%%:*}"; defs="${v#*:}"
  ( /usr/local/cuda/bin/nvcc -DSU_PHASE_SYNC $defs -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
      -Xcompiler -fPIC -shared -o ../../scratch/variants/lib_$name.so suews_b200.cu suews_kernel_small.cu suews_refapi.cpp -lcudart -Xptxas -v 2>&1 \
      | grep -A2 "suews_run_kernelILb0" | grep -o "Used [0-9]* registers\|[0-9]* bytes stack frame" | tr '\n' ' ' | sed "s/^/$name: /"; echo ) &
done
wait
