#!/bin/bash
# Round-2 evidence on one B200 (run through gpurun): bench lines for every workload + the reference arm, the ncu launch
# list of the bench command itself, and `ncu --set full` captures of the traversal kernels on three workloads
# (L2-resident Teapot and dragon, HBM-resident 20.8 M triangles). Everything lands in gpurun_out/<tag>_*.
tag=${1:-r2}
out=gpurun_out
EXTRA="--metrics lts__t_bytes.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed"
python bench.py > $out/${tag}_bench_teapot.json 2> $out/${tag}_bench_teapot.err
python bench.py --impl reference > $out/${tag}_bench_teapot_reference.json 2> $out/${tag}_bench_teapot_reference.err
python bench.py --workload cornell > $out/${tag}_bench_cornell.json 2> $out/${tag}_bench_cornell.err
python bench.py --workload dragon --steps 2 > $out/${tag}_bench_dragon.json 2> $out/${tag}_bench_dragon.err
python bench.py --workload vwvan --steps 3 > $out/${tag}_bench_vwvan.json 2> $out/${tag}_bench_vwvan.err
python bench.py --workload blobs20m --spp 32 --steps 2 --no-cpu-baseline > $out/${tag}_bench_blobs20m.json 2> $out/${tag}_bench_blobs20m.err
python bench.py --workload blobs871k --steps 2 --no-cpu-baseline > $out/${tag}_bench_blobs871k.json 2> $out/${tag}_bench_blobs871k.err
# launch list of the bench command itself (per-launch durations: serialised, cold caches -> shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/${tag}_launches_bench_teapot.csv python bench.py --steps 1 --warmup 1 --spp 8 --no-cpu-baseline > $out/${tag}_launches_bench_teapot.log 2>&1
for w in teapot dragon blobs20m; do
  TB_FIF=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/${tag}_launches_${w}.csv python tools/profile_run.py $w 4 > $out/${tag}_prof_${w}.log 2>&1
done
# ncu --set full of the traversal kernels at BOUNCE 0 of the second rendered frame (one frame at a time: every bounce is
# k_extend<0> + three resume rounds [+ k_extend<1> shadow + k_extend<2> walk on scenes with lights / glass])
skip_teapot=24; skip_dragon=32; skip_blobs20m=36
for w in teapot dragon blobs20m; do
  eval s=\$skip_$w
  TB_FIF=1 timeout 900 ncu --set full $EXTRA --import-source on --clock-control none -k regex:k_extend -s $s -c 6 -o $out/${tag}_extend_${w} python tools/profile_run.py $w 2 > $out/${tag}_ncu_extend_${w}.log 2>&1
done
TB_FIF=1 timeout 900 ncu --set full $EXTRA --import-source on --clock-control none -k regex:k_walk -s 12 -c 3 -o $out/${tag}_walk_blobs20m python tools/profile_run.py blobs20m 2 > $out/${tag}_ncu_walk_blobs20m.log 2>&1
TB_FIF=1 timeout 900 ncu --set full $EXTRA --import-source on --clock-control none -k regex:k_shade -s 12 -c 4 -o $out/${tag}_shade_vwvan python tools/profile_run.py vwvan 2 > $out/${tag}_ncu_shade_vwvan.log 2>&1
# raw metric pages + the source page (SASS / source line counters) as text; the .ncu-rep files themselves (10-20 MB each)
# stay on the box: gpurun_out/ is capped at 64 MiB
for f in $out/${tag}_*.ncu-rep; do
  ncu -i $f --page raw --csv > ${f%.ncu-rep}_raw.csv 2>/dev/null
  rm -f $f
done
for f in teapot cornell dragon vwvan blobs20m blobs871k; do python - <<PY
import json
try:
    d = json.loads(open("$out/${tag}_bench_$f.json").read().strip().splitlines()[-1])
    inc = d.get("incoherent") or {}
    print("$f", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "incoherent", round(inc.get("mrays_per_s", 0), 1), "frac", round(d["roofline"]["frac"], 3),
          "share", round(d["roofline"]["kernel_share_of_step"], 3), "build ms", round(d["bvh_build_ms"], 2))
except Exception as e:
    print("$f", "failed", e)
PY
done
tail -c 300 $out/${tag}_bench_teapot_reference.json
ls -la $out | tail -40
