#!/bin/bash
# full GPU test suite + the default bench line (N=1) + smoke
mkdir -p gpurun_out
export KICP_SPIN_TIMEOUT_MS=20000
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -8 gpurun_out/r2_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r2_bench.err; python - <<'PY'
import json
try:
    l=json.loads(open('gpurun_out/r2_bench.json').read().strip().split('\n')[-1])
    print({k:l[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', l['e2e']['value'], {k:round(v['value']) for k,v in l['e2e']['variants'].items()})
    print('roofline frac', l['roofline']['frac'], 'touched', l['roofline']['touched'])
    print('anatomy', l['pass_anatomy_us'])
    print('cpu', l['cpu_baseline'])
    print('pose', l['pose_delta_vs_cpu'])
except Exception as e: print('parse failed', e)
PY
