#!/bin/bash
# Traversal experiments built by tools/build_variant.py (TB_LIB selects the library), Mrays/s of a warmed render.
# usage: tools/variant_sweep.sh variant [variant ...]      ("base" = the product library)
for w in teapot dragon vwvan blobs20m; do
  spp=64; [ $w = blobs20m ] && spp=32; [ $w = vwvan ] && spp=32
  for v in "$@"; do
    lib=tracerboy_b200/lib/libtb_var_$v.so
    [ $v = base ] && lib=tracerboy_b200/lib/libtracerboy_b200.so
    echo -n "$w $v: "
    TB_LIB=$PWD/$lib python tools/profile_run.py $w $spp | tail -1
  done
done
