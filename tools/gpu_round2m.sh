set -u
out=gpurun_out/r2m
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -4 $out/pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench_c2.json 2> $out/bench_c2.err; echo "bench rc=$?"
tail -c 600 $out/bench_c2.err
python - <<'PY'
import json
for f in ("bench_c2",):
    try:
        d=json.loads(open(f"gpurun_out/r2m/{f}.json").read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ("value","ms_per_step","stage_wall_s","host_decode_s","device_call_s","rows_s","n_gpus")}, "e2e", d.get("e2e",{}).get("ms_per_step"), d["roofline"]["kernel"], round(d["roofline"]["frac"],3), "path", round(d["path_roofline"]["frac"],3))
        print("  kernel_ms", {k:round(v*1e3,1) for k,v in d["kernel_ms"].items()})
        for k,v in (d.get("configs") or {}).items():
            print("   ",k, round(v["ms_per_step"],4), "e2e", round(v["e2e"]["ms_per_step"],3), "roof", round(v["roofline"]["frac"],3), v["roofline"]["kernel"], "path", round(v["path_roofline"]["frac"],3), v.get("cpu_baseline",{}).get("value"))
        if "cpu_baseline" in d: print("   cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"])
    except Exception as e:
        print(f, "ERR", e)
PY
