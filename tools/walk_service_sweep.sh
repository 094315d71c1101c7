#!/bin/bash
# k_walk: how many finished rays end a traversal phase (TB_WALK_SERVICE), Mrays/s of a warmed render.
for w in blobs20m vwvan blobs871k; do
  for v in 8 2 4 12 16 24 8; do
    echo -n "$w service_at=$v: "
    TB_WALK_SERVICE=$v python tools/profile_run.py $w 32 | tail -1
  done
done
