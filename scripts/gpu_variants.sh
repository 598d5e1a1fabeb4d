#!/bin/bash
# parity tests + a sweep of kernel variants / binning granularity
mkdir -p gpurun_out
if [ "$SKIP_TESTS" != "1" ]; then echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5; fi
for cfg in ${CFGS:-4 3}; do
for v in ${VARIANTS:-pruned:30 pruned:0 staged:30}; do
  a=${v%%:*}; b=${v##*:}
  echo "== cfg$cfg assoc=$a sort_bits=$b"
  KICP_ASSOC=$a KICP_SORT_BITS=$b timeout 600 python bench.py --workload $cfg --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.0f scans/s  ms/step %.3f  iters %d  e2e %.0f  kernel_us %.1f  frac %.2f  split %s  delta %s'%(d['value'],d['ms_per_step'],d['config']['iterations_per_registration'],d['e2e']['value'],d['roofline']['kernel_us'],d['roofline']['frac'],{k[:6]:round(v,3) for k,v in d['kernel_time_split_ms_per_step'].items()},d['pose_delta_vs_cpu']['translation_m']))
    else: print(l.rstrip())
"
done; done
