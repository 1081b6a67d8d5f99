#!/bin/bash
# Builds tuning variants of libsuews_b200.so into build/variants/ (launch geometry, inlining switches); time them on
# the GPU box with tools/run_variants.sh (SUEWS_B200_LIB selects the library a process loads).
#   usage: tools/variants.sh "name:-DFLAG1 -DFLAG2" ...
set -e
cd "$(dirname "$0")/../suews_b200/csrc"
mkdir -p ../../build/variants
for v in "$@"; do
  name="${v%%:*}"; defs="${v#*:}"
  ( make -s OUT=../../build/variants/lib_$name.so OBJDIR=../../build/variants/obj_$name \
       DEFS="-DSU_PHASE_SYNC -DSU_EXPLOG_INLINE -DSU_DEV_INLINE $defs" 2>&1 \
      | grep -A2 "suews_run_kernelILi2ELi0ELi0" | grep -o "Used [0-9]* registers\|[0-9]* bytes stack frame" | tr '\n' ' ' | sed "s/^/$name: /"; echo ) &
done
wait
