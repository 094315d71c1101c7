for v in base tlmb9 tlmb10 tlmb12 base; do
  lib=tracerboy_b200/lib/libtb_var_$v.so; [ $v = base ] && lib=tracerboy_b200/lib/libtracerboy_b200.so
  for w in dragon blobs20m; do echo -n "$v: "; TB_LIB=$PWD/$lib python tools/build_only.py $w | tail -1; done
done
