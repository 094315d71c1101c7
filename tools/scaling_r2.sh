#!/bin/bash
# Multi-GPU evidence on one 8-GPU box (gpurun --gpus 8): the library's NCCL communicator at 2 / 4 / 8 ranks.
#   dragon (BASELINE configs[2]) STRONG scaling, 256 spp per step in total: sample sharding at N = 1, 2, 4, 8 and row bands at N = 8
#   Teapot weak scaling at N = 8 (the driver's own SCALE run uses this default)
#   the bit-identity / fixed-order tests at 2, 4 and 8 ranks
tag=${1:-r2s}
out=gpurun_out
run() { # N, extra args..., output name
  n=$1; shift; name=$1; shift
  if [ $n = 1 ]; then python bench.py "$@" > $out/${tag}_$name.json 2> $out/${tag}_$name.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" > $out/${tag}_$name.json 2> $out/${tag}_$name.err; fi
  tail -c 300 $out/${tag}_$name.err
}
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4 > $out/${tag}_multi_tests.log; cat $out/${tag}_multi_tests.log
for n in 1 2 4 8; do run $n dragon_strong_samples_$n --workload dragon --scaling strong --steps 3 --no-cpu-baseline; done
run 8 dragon_rows_8 --workload dragon --shard rows --steps 3 --no-cpu-baseline
run 8 teapot_weak_8 --steps 5 --no-cpu-baseline
run 8 blobs20m_weak_8 --workload blobs20m --spp 32 --steps 2 --no-cpu-baseline
python - <<PY
import json, glob
base = None
for f in sorted(glob.glob("$out/${tag}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "failed", e); continue
    print(f.split("/")[-1], "N=%d" % d["n_gpus"], d["scaling"], "value %.0f Mrays/s" % d["value"], "e2e %.0f" % d["e2e"]["value"], "ms/step %.2f" % d["ms_per_step"],
          "reduce ms %s" % d.get("reduce_ms_per_step"), "samples/s %.3g" % d["samples_per_s"])
PY
