# phasing: the GPU parity tests, then timelines.  usage: gpurun -- 'bash tools/gpu_phase.sh <tag> [workloads...]'
set -u
out=gpurun_out/${1:-ph}
mkdir -p $out
shift
timeout 1500 python -m pytest tests/test_gpu_phase.py -m gpu -x -q > $out/pytest_phase.log 2>&1; echo "pytest rc=$?" >> $out/pytest_phase.log
tail -4 $out/pytest_phase.log
for w in ${@:-c2 c5}; do
  timeout 300 python tools/kernel_timeline.py $w > $out/timeline_$w.txt 2>&1
  grep -E "^===" $out/timeline_$w.txt
  grep -A4 "^k_probe" $out/timeline_$w.txt
done
