#!/bin/bash
# Final pass of the round on one B200: the whole GPU test suite, then the bench lines of every workload + the reference arm.
tag=${1:-r2z}; out=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | grep -v "pbrt-parser\|created object" | tail -12 > $out/${tag}_gpu_tests.log; cat $out/${tag}_gpu_tests.log
python bench.py > $out/${tag}_bench_teapot.json 2> $out/${tag}_bench_teapot.err
python bench.py --impl reference > $out/${tag}_bench_teapot_reference.json 2> $out/${tag}_bench_teapot_reference.err
python bench.py --workload cornell --no-cpu-baseline > $out/${tag}_bench_cornell.json 2> $out/${tag}_bench_cornell.err
python bench.py --workload dragon --steps 2 --no-cpu-baseline > $out/${tag}_bench_dragon.json 2> $out/${tag}_bench_dragon.err
python bench.py --workload vwvan --steps 3 --no-cpu-baseline > $out/${tag}_bench_vwvan.json 2> $out/${tag}_bench_vwvan.err
python bench.py --workload blobs20m --spp 32 --steps 2 --no-cpu-baseline > $out/${tag}_bench_blobs20m.json 2> $out/${tag}_bench_blobs20m.err
python bench.py --workload blobs871k --steps 2 --no-cpu-baseline > $out/${tag}_bench_blobs871k.json 2> $out/${tag}_bench_blobs871k.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
# launch list of the bench command itself (per-launch durations: serialised, cold caches -> shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/${tag}_launches_bench_teapot.csv python bench.py --steps 1 --warmup 1 --spp 8 --no-cpu-baseline > $out/${tag}_launches_bench_teapot.log 2>&1
for f in teapot cornell dragon vwvan blobs20m blobs871k; do python - <<PY
import json
try:
    d = json.loads(open("$out/${tag}_bench_$f.json").read().strip().splitlines()[-1])
    inc = d.get("incoherent") or {}
    print("$f", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "incoherent", round(inc.get("mrays_per_s", 0), 1), "frac", round(d["roofline"]["frac"], 3), "build ms", round(d["bvh_build_ms"], 2))
except Exception as e:
    print("$f", "failed", e)
PY
done
tail -c 400 $out/${tag}_bench_teapot_reference.json
