set -u
out=gpurun_out/r2a
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?" >> $out/smoke.log
timeout 300 python tools/kernel_timeline.py c2 > $out/timeline_c2.txt 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_c2.json 2> $out/bench_c2.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload c4 > $out/bench_c4.json 2> $out/bench_c4.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload c5 > $out/bench_c5.json 2> $out/bench_c5.err
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_phase.py -m gpu -x -q -k "multi_contig_small or last_row or tie_order or many_phase or empty_oneps or big_shard" > $out/memcheck.log 2>&1; echo "memcheck rc=$?" >> $out/memcheck.log
tail -3 $out/pytest.log; tail -2 $out/smoke.log; tail -5 $out/memcheck.log; cat $out/timeline_c2.txt
