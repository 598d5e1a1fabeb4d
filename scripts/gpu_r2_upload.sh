#!/bin/bash
# staged upload of pageable buffers: GPU suite with it on, then the e2e variants with it on and off on the same box
mkdir -p gpurun_out
export KICP_SPIN_TIMEOUT_MS=20000
T0=$(date +%s)
el() { echo "[+$(( $(date +%s) - T0 ))s] $*"; }
timeout 200 python -m pytest tests -x -q -m gpu > gpurun_out/up_gputests.log 2>&1; el "gpu tests rc=$?"; tail -n 3 gpurun_out/up_gputests.log
timeout 120 python bench.py --steps 20 --warmup 5 > gpurun_out/up_bench_on.json 2> gpurun_out/up_bench_on.err; el "bench (3 helpers) rc=$?"; tail -n 2 gpurun_out/up_bench_on.err
KICP_UPLOAD_THREADS=0 timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/up_bench_off.json 2> gpurun_out/up_bench_off.err; el "bench (driver staging) rc=$?"
KICP_UPLOAD_THREADS=7 timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustained 0 > gpurun_out/up_bench_7.json 2> gpurun_out/up_bench_7.err; el "bench (7 helpers) rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/up_bench_on.json', 'gpurun_out/up_bench_off.json', 'gpurun_out/up_bench_7.json'):
    try:
        l = json.loads([x for x in open(f).read().strip().split('\n') if x.startswith('{')][-1])
        print(f, round(l['value']), 'e2e', {k: round(v['value']) for k, v in l['e2e']['variants'].items()}, 'replay', (l.get('replay') or {}).get('value'), (l.get('replay') or {}).get('pageable_host_buffers'))
    except Exception as e:
        print(f, 'parse failed', e)
PY
el done
