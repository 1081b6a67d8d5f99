#!/bin/bash
# Builds tuning variants of libsuews_b200.so into build/variants/ (launch geometry, inlining switches); time them on
# the GPU box with tools/run_variants.sh (SUEWS_B200_LIB selects the library a process loads).
#   usage: tools/variants.sh "name|MAKEVAR=.. MAKEVAR=..|-DFLAG1 -DFLAG2" ...     (built one after the other, 8 jobs each)
set -e
cd "$(dirname "$0")/../suews_b200/csrc"
mkdir -p ../../build/variants
for v in "$@"; do
  IFS='|' read -r name mvars defs <<< "$v"
  make -s -j8 ../../build/variants/lib_$name.so OUT=../../build/variants/lib_$name.so OBJDIR=../../build/variants/obj_$name $mvars \
       DEFS="-DSU_PHASE_SYNC -DSU_EXPLOG_INLINE -DSU_DEV_INLINE $defs" 2>&1 \
      | grep -A2 "suews_run_kernelILi2ELi0ELi0" | grep -o "Used [0-9]* registers\|[0-9]* bytes stack frame\|[0-9]* bytes spill stores" | tr '\n' ' ' | sed "s/^/$name: /"; echo; rm -rf ../../build/variants/obj_$name
done
