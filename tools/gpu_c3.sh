# clustering: parity tests + the C3 bench line.  usage: gpurun -- 'bash tools/gpu_c3.sh <tag> [ncu]'
set -u
out=gpurun_out/${1:-c3}
mkdir -p $out
timeout 300 python -m pytest tests/test_cluster.py -m gpu -x -q > $out/pytest_cluster.log 2>&1; echo "pytest rc=$?" >> $out/pytest_cluster.log
tail -15 $out/pytest_cluster.log
timeout 300 python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_c3.json 2> $out/bench_c3.err; echo "c3 rc=$?"
tail -5 $out/bench_c3.err
python - $out <<'PY'
import json, sys
d=json.loads(open(sys.argv[1] + "/bench_c3.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","clusters","gpu_launches")}, "e2e", d["e2e"]["ms_per_step"], d.get("kernel_ms"), d["roofline"]["kernel"], d["roofline"]["frac"], d["path_roofline"]["frac"])
PY
if [ "${2:-}" = "ncu" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_cl_|k_rs_" -s 9 -c 3 -f -o $out/prof_c3 \
      python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > $out/ncu_c3.log 2>&1
  ncu -i $out/prof_c3.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum 2>/dev/null | cut -d, -f5,12- | head -40
fi
