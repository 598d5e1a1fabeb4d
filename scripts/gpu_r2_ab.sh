#!/bin/bash
# same-box A/B of library builds (kinematic-icp_b200/lib/ab/*.so, same ABI) with the flushed bench loop
mkdir -p gpurun_out
export KICP_SPIN_TIMEOUT_MS=20000
run() { python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload ${1:-4} 2>/dev/null | python -c "
import json,sys
l=json.loads([x for x in sys.stdin.read().strip().split('\n') if x.startswith('{')][-1])
print(round(l['value']), 'scans/s', round(l['ms_per_step']*1e3),'us  e2e f32', round(l['e2e']['value']), ' anatomy', [round(x,1) for x in l['pass_anatomy_us']['median_over_passes_max_over_ranks']])"; }
{
for rep in 1 2; do
echo "== default (rep $rep)"; run
for so in kinematic-icp_b200/lib/ab/*.so; do echo "== $so (rep $rep)"; KICP_LIB=$PWD/$so run; done
done
} 2>&1 | tee gpurun_out/r2_ab.log
