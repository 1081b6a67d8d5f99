#!/bin/bash
# time several builds of the library on the profiling workload (run on the GPU box)
# usage: tools/try_variants.sh lib.so[:ncell] ...
cp suews_b200/libsuews_b200.so /tmp/orig.so
for spec in "$@"; do
  f=${spec%%:*}; n=${spec##*:}; [ "$n" = "$spec" ] && n=151552
  cp "$f" /tmp/variant.so && cp /tmp/variant.so suews_b200/libsuews_b200.so
  echo "== $f ncell=$n"
  python tools/prof_run.py $n 48 2>&1 | grep "launch 2"
done
cp /tmp/orig.so suews_b200/libsuews_b200.so
