#!/bin/bash
# GPU tests, then the bench line of every config on one GPU (no profiler), then the reference arm
mkdir -p gpurun_out
export KICP_SPIN_TIMEOUT_MS=20000
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r2_gputests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_n1.err
for wl in 3 2 1; do timeout 300 python bench.py --steps 20 --warmup 5 --workload $wl --no-cpu-baseline > gpurun_out/r2_bench_cfg$wl.json 2>> gpurun_out/r2_bench_n1.err; done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2>> gpurun_out/r2_bench_n1.err; echo "reference arm rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_n1.json','gpurun_out/r2_bench_cfg3.json','gpurun_out/r2_bench_cfg2.json','gpurun_out/r2_bench_cfg1.json','gpurun_out/r2_bench_ref.json'):
    try:
        l=json.loads([x for x in open(f).read().strip().split('\n') if x.startswith('{')][-1])
        print(f, round(l['value'],1), 'scans/s', round(l['ms_per_step']*1e3),'us e2e', round(l['e2e']['value'],1), {k:round(v['value']) for k,v in l['e2e'].get('variants',{}).items()}, 'traffic', l.get('roofline',{}).get('traffic'))
    except Exception as e: print(f, 'parse failed', e)
PY
