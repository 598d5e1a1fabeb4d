#!/bin/bash
# same-box A/B of library builds (kinematic-icp_b200/lib/ab/*.so, same ABI) with the flushed bench loop
mkdir -p gpurun_out
export KICP_SPIN_TIMEOUT_MS=20000
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "registration" 2>&1 | tail -2
KICP_LIB=$PWD/kinematic-icp_b200/lib/ab/libkicp_w10_g2_s1.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "registration" 2>&1 | tail -2
run() { python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload ${1:-4} 2>/dev/null | python -c "
import json,sys
l=json.loads([x for x in sys.stdin.read().strip().split('\n') if x.startswith('{')][-1])
print(round(l['value']), 'scans/s', round(l['ms_per_step']*1e3),'us  e2e f32', round(l['e2e']['value']), ' anatomy', [round(x,1) for x in l['pass_anatomy_us']['median_over_passes_max_over_ranks']])"; }
{
for rep in 1 2; do
echo "== default (rep $rep)"; run
for so in kinematic-icp_b200/lib/ab/*.so; do echo "== $so (rep $rep)"; KICP_LIB=$PWD/$so run; done
done
echo "== cfg3: default / top1 / w10g4"; run 3; KICP_LIB=$PWD/kinematic-icp_b200/lib/ab/libkicp_top1.so run 3; KICP_LIB=$PWD/kinematic-icp_b200/lib/ab/libkicp_w10_g2_s0.so run 3
} 2>&1 | tee gpurun_out/r2_ab.log
