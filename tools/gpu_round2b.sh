set -u
out=gpurun_out/r2b
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_phase.py -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -3 $out/pytest.log
timeout 600 python tools/kernel_timeline.py c2 0,4,8,1,2,3,16,32 > $out/timeline_c2.txt 2>&1
timeout 300 python tools/kernel_timeline.py c5 0,4,16 > $out/timeline_c5.txt 2>&1
timeout 300 python tools/kernel_timeline.py c4 0,4 > $out/timeline_c4.txt 2>&1
grep "===" $out/timeline_c2.txt $out/timeline_c5.txt $out/timeline_c4.txt
