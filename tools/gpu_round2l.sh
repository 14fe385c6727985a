set -u
out=gpurun_out/r2l
mkdir -p $out
nvidia-smi -L > $out/gpus.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -4 $out/pytest.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench_c2.json 2> $out/bench_c2.err; echo "bench rc=$?"
tail -c 600 $out/bench_c2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_n2.json 2> $out/bench_n2.err; echo "n2 rc=$?"
tail -c 400 $out/bench_n2.err
python - <<'PY'
import json
for f in ("bench_ref","bench_c2","bench_n2"):
    try:
        d=json.loads(open(f"gpurun_out/r2l/{f}.json").read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ("value","ms_per_step","stage_wall_s","decode_s","post_decode_s","n_gpus")}, "e2e", d.get("e2e",{}).get("ms_per_step"))
        for k,v in (d.get("configs") or {}).items():
            print("   ",k, round(v["ms_per_step"],4), "e2e", round(v["e2e"]["ms_per_step"],3), "roof", round(v["roofline"]["frac"],3), v["roofline"]["kernel"], "path", round(v["path_roofline"]["frac"],3))
        if "strong" in d: print("   strong", {k:d["strong"][k] for k in ("value","ms_per_step","slices_equal_unsharded","load_share_max","n_emitted_gathered","n_emitted_unsharded")})
        if "cpu_baseline" in d: print("   cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"])
    except Exception as e:
        print(f, "ERR", e)
PY
