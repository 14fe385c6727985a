set -u
for b in ${@:-14 15}; do
  DUET_CL_BITS=$b timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bits $b', d['ms_per_step'], d.get('kernel_ms'))"
done
