#!/bin/bash
# time several builds of the library on the profiling workload (run on the GPU box)
cp suews_b200/libsuews_b200.so /tmp/orig.so
for f in "$@"; do
  cp "$f" suews_b200/libsuews_b200.so
  echo "== $f"
  python tools/prof_run.py 151552 48 2>&1 | grep "launch 2"
done
cp /tmp/orig.so suews_b200/libsuews_b200.so
