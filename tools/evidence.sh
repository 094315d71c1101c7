#!/bin/bash
# End-of-round evidence on one B200 (run through gpurun): tests, bench lines for every workload, the reference arm,
# ncu launch lists and one ncu --set full capture of k_extend. Everything lands in gpurun_out/<tag>_*.
tag=${1:-ev}
out=gpurun_out
python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1; tail -n 2 $out/${tag}_tests.log
python bench.py > $out/${tag}_bench_teapot.json 2> $out/${tag}_bench_teapot.err
python bench.py --impl reference > $out/${tag}_bench_teapot_reference.json 2> $out/${tag}_bench_teapot_reference.err
python bench.py --workload cornell > $out/${tag}_bench_cornell.json 2> $out/${tag}_bench_cornell.err
python bench.py --workload dragon --steps 2 > $out/${tag}_bench_dragon.json 2> $out/${tag}_bench_dragon.err
python bench.py --workload vwvan --steps 3 > $out/${tag}_bench_vwvan.json 2> $out/${tag}_bench_vwvan.err
python bench.py --workload blobs20m --spp 32 --steps 2 --no-cpu-baseline > $out/${tag}_bench_blobs20m.json 2> $out/${tag}_bench_blobs20m.err
for w in teapot dragon vwvan; do
  TB_FIF=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/${tag}_launches_${w}.csv python tools/profile_run.py $w 4 > $out/${tag}_prof_${w}.log 2>&1
done
TB_FIF=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_extend -s 24 -c 2 -o $out/${tag}_extend_teapot python tools/profile_run.py teapot 2 > $out/${tag}_ncu_extend.log 2>&1
for w in teapot dragon vwvan blobs20m; do python tools/build_only.py $w; done > $out/${tag}_build.log 2>&1
for f in teapot cornell dragon vwvan blobs20m; do python - <<PY
import json
try:
    d = json.loads(open("$out/${tag}_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3), "build ms", round(d["bvh_build_ms"], 2))
except Exception as e:
    print("$f", "failed", e)
PY
done
tail -c 400 $out/${tag}_bench_teapot_reference.json
