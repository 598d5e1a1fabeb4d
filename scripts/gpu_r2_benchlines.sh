#!/bin/bash
# bench lines of every config on one box (default run = cfg4 with the CPU arm, the sustained figure and the pipeline replay)
mkdir -p gpurun_out
export KICP_SPIN_TIMEOUT_MS=20000
T0=$(date +%s)
el() { echo "[+$(( $(date +%s) - T0 ))s] $*"; }
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/fin_bench_n1.json 2> gpurun_out/fin_bench_n1.err; el "bench cfg4 rc=$?"; tail -n 2 gpurun_out/fin_bench_n1.err
for wl in 3 2 1; do timeout 120 python bench.py --steps 20 --warmup 5 --workload $wl --no-cpu-baseline --no-replay > gpurun_out/fin_bench_cfg$wl.json 2>> gpurun_out/fin_bench_n1.err; el "bench cfg$wl rc=$?"; done
python - <<'PY'
import json
for f in ('gpurun_out/fin_bench_n1.json', 'gpurun_out/fin_bench_cfg3.json', 'gpurun_out/fin_bench_cfg2.json', 'gpurun_out/fin_bench_cfg1.json'):
    try:
        l = json.loads([x for x in open(f).read().strip().split('\n') if x.startswith('{')][-1])
        print(f, round(l['value']), 'scans/s e2e', round(l['e2e']['value']), 'sustained', l.get('sustained'), 'traffic', l['roofline']['traffic'], 'clocks', l['clocks']['sm_mhz'], l['clocks']['reasons'])
    except Exception as e:
        print(f, 'parse failed', e)
PY
el done
