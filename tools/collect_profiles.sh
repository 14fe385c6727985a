#!/bin/bash
# Round-end evidence, one GPU: bench lines, the ncu launch list, one `--set full` capture of the five
# kernels of a phasing call and of the clustering kernels, and the per-block timeline.  Everything lands in
# gpurun_out/prof/ (scratch); tools/summarize_profiles.py <tag> turns it into the tracked files under profiles/.
#   gpurun --timeout 2400 -- 'bash tools/collect_profiles.sh'
set -u
out=gpurun_out/prof
mkdir -p $out
timeout 900 python bench.py > $out/bench_c2.json 2> $out/bench_c2.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err
timeout 300 python tools/kernel_timeline.py c2 > $out/timeline_c2.txt 2>&1
timeout 300 python tools/kernel_timeline.py c5 > $out/timeline_c5.txt 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline --no-stage-wall > $out/ncu_launch.log 2>&1
# warm-up = 3 steps of 5 launches; the capture takes the 5 kernels of the first timed step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ -s 15 -c 5 -f -o $out/prof_all \
    python bench.py --steps 3 --warmup 3 --no-configs --no-cpu-baseline --no-stage-wall > $out/ncu_full.log 2>&1
# clustering: five launches per call (k_cl_max, k_cl_hist, k_cl_scatter, k_cl_bucket, k_cl_fix); 3 warm-up calls
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_cl_|k_rs_" -s 15 -c 5 -f -o $out/prof_c3 \
    python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > $out/ncu_c3.log 2>&1
timeout 600 python bench.py --workload c3 --steps 20 --warmup 5 > $out/bench_c3.json 2> $out/bench_c3.err
ls -la $out | tail -15
