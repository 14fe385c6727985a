set -u
out=gpurun_out/r2c
mkdir -p $out
timeout 600 python tools/kernel_timeline.py c2 0,64,128,192,196 > $out/timeline_c2.txt 2>&1
timeout 300 python tools/kernel_timeline.py c2 0,4 --no-flush > $out/timeline_c2_noflush.txt 2>&1
grep "===" $out/timeline_c2.txt $out/timeline_c2_noflush.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ -s 8 -c 4 -f -o $out/prof_c2_cold \
    python bench.py --steps 3 --warmup 3 --no-configs --no-cpu-baseline --no-stage-wall > $out/ncu_cold.log 2>&1
timeout 900 ncu --set full --clock-control none --cache-control none -k regex:k_ -s 8 -c 4 -f -o $out/prof_c2_warm \
    python bench.py --steps 3 --warmup 3 --no-configs --no-cpu-baseline --no-stage-wall > $out/ncu_warm.log 2>&1
tail -3 $out/ncu_cold.log
ls -la $out
