set -u
out=gpurun_out/r2n
mkdir -p $out
timeout 900 python -m pytest tests/test_cluster.py -m gpu -x -q > $out/pytest_cluster.log 2>&1; echo "pytest rc=$?" >> $out/pytest_cluster.log
tail -6 $out/pytest_cluster.log
timeout 600 python bench.py --workload c3 --steps 20 --warmup 5 > $out/bench_c3.json 2> $out/bench_c3.err; echo "c3 rc=$?"
tail -c 400 $out/bench_c3.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2n/bench_c3.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","clusters","gpu_launches")}, "e2e", d["e2e"]["ms_per_step"], d.get("kernel_ms"), d["roofline"]["kernel"], d["roofline"]["frac"], d["path_roofline"]["frac"], d.get("cpu_baseline",{}).get("value"))
PY
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_cluster.py -m gpu -x -q -k "not 2m and not config" > $out/racecheck_cluster.log 2>&1; tail -4 $out/racecheck_cluster.log
