set -u
out=gpurun_out/r2j
mkdir -p $out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?" >> $out/smoke.log; tail -3 $out/smoke.log
timeout 900 python -m pytest tests/test_gpu_phase.py -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -5 $out/pytest.log
timeout 600 python tools/kernel_timeline.py c2 0 --clean-flush > $out/timeline_c2.txt 2>&1
timeout 600 python tools/kernel_timeline.py c2 0 > $out/timeline_c2_dirty.txt 2>&1
timeout 300 python tools/kernel_timeline.py c5 0 --clean-flush > $out/timeline_c5.txt 2>&1
timeout 300 python tools/kernel_timeline.py c4 0 --clean-flush > $out/timeline_c4.txt 2>&1
timeout 300 python tools/kernel_timeline.py c1 0 --clean-flush > $out/timeline_c1.txt 2>&1
grep "===" $out/timeline_c*.txt
sed -n 2,20p $out/timeline_c2.txt | cut -c1-200
