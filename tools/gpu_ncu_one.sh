# ncu --set full of one kernel of a bench workload: bash tools/gpu_ncu_one.sh <tag> <kernel regex> <workload> [skip]
set -u
out=gpurun_out/$1
mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" -s ${4:-3} -c 1 -f -o $out/prof_one \
    python bench.py --workload $3 --steps 3 --warmup 3 --no-cpu-baseline > $out/ncu_one.log 2>&1
tail -3 $out/ncu_one.log
