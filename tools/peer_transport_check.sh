#!/bin/bash
# The two transports of the reduction side by side on an N-GPU box (gpurun --gpus N): the bit-identity tests at N ranks,
# then dragon strong scaling at N ranks with each transport (reduce ms per step is the figure to compare).
n=${1:-2}; tag=${2:-p$n}; out=gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "${3:-${n}-}" 2>&1 | tail -15 > $out/${tag}_multi_tests.log; cat $out/${tag}_multi_tests.log
for t in peer nccl; do
  TB_COMM_TRANSPORT=$t timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --workload dragon --scaling strong --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_dragon_strong_${n}_$t.json 2> $out/${tag}_dragon_strong_${n}_$t.err
  tail -c 300 $out/${tag}_dragon_strong_${n}_$t.err
  python - <<PY
import json
d = json.loads(open("$out/${tag}_dragon_strong_${n}_$t.json").read().strip().splitlines()[-1])
print("$t", "N=%d" % d["n_gpus"], "value %.0f" % d["value"], "ms/step %.2f" % d["ms_per_step"], "reduce ms", d.get("reduce_ms_per_step"), d["comm"])
PY
done
