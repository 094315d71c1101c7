#!/bin/bash
# Traversal layout 0 (pairs) / 1 (4-wide nodes): Mrays/s of a warmed render per workload.
for w in teapot dragon vwvan blobs871k blobs20m cornell; do
  spp=64; [ $w = blobs20m ] && spp=32; [ $w = vwvan ] && spp=32
  for m in 0 1; do
    echo -n "$w wide=$m: "
    TB_WIDE=$m python tools/profile_run.py $w $spp | tail -1
  done
done
