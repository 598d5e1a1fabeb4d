#!/bin/bash
for p in 1 0; do for a in group4 pruned; do
echo "== cfg1 persistent=$p assoc=$a"
KICP_PERSISTENT=$p KICP_ASSOC=$a timeout 300 python bench.py --workload 1 --steps 30 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step %.3f iters %d kernel_us/iter %.1f split %s'%(d['ms_per_step'],d['config']['iterations_per_registration'],d['roofline']['kernel_us'],{k[:6]:round(v,3) for k,v in d['kernel_time_split_ms_per_step'].items()}))"
done; done
