#!/bin/bash
mkdir -p gpurun_out
export KICP_SPIN_TIMEOUT_MS=20000
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
run() { python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload $1 2>/dev/null | python -c "
import json,sys
l=json.loads([x for x in sys.stdin.read().strip().split('\n') if x.startswith('{')][-1])
print('cfg$1', round(l['value']), 'scans/s', round(l['ms_per_step']*1e3),'us  e2e f32', round(l['e2e']['value']), {k:round(v['value']) for k,v in l['e2e']['variants'].items()}, ' anatomy', [round(x,1) for x in l['pass_anatomy_us']['median_over_passes_max_over_ranks']])"; }
for wl in 4 3 2 1; do run $wl; done
