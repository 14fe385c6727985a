#!/bin/bash
# Round-end evidence, one GPU: bench lines, the ncu launch list, one `--set full` capture of the seven
# kernels of a call, and the per-block timeline.  Everything lands in gpurun_out/ (scratch); the
# summaries that are kept are copied into profiles/ by hand.
#   gpurun --timeout 1500 -- 'bash tools/collect_profiles.sh'
set -u
out=gpurun_out
mkdir -p $out
for w in c2 c1 c4 c5; do
  timeout 300 python bench.py --workload $w > $out/bench_$w.json 2> $out/bench_$w.err
done
timeout 300 python bench.py --impl reference > $out/bench_ref.json 2> $out/bench_ref.err
timeout 300 python tools/kernel_timeline.py c2 > $out/timeline_c2.txt 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 1 > $out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ -s 28 -c 7 -f -o $out/prof_all \
    python bench.py --steps 3 --warmup 3 > $out/ncu_full.log 2>&1
ls -la $out | tail -15
