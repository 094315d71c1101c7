#!/bin/bash
# Material-class hit queues on / off (TB_CLASS_SORT overrides the automatic policy), Mrays/s of a warmed render.
for w in teapot dragon vwvan blobs871k blobs20m cornell; do
  spp=64; [ $w = blobs20m ] && spp=32; [ $w = vwvan ] && spp=32
  for m in 0 1; do
    echo -n "$w class_sort=$m: "
    TB_CLASS_SORT=$m python tools/profile_run.py $w $spp | tail -1
  done
done
