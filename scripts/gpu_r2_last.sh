#!/bin/bash
# the round's last GPU call, trimmed to the minutes left: the GPU suite on the final tree, one full-set ncu capture of a registration
# launch (DRAM traffic stamp for the final source), the ncu launch list, the default bench line
mkdir -p gpurun_out
export KICP_SPIN_TIMEOUT_MS=20000
T0=$(date +%s)
el() { echo "[+$(( $(date +%s) - T0 ))s] $*"; }
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/last_gputests.log 2>&1; el "gpu tests rc=$?"; tail -n 3 gpurun_out/last_gputests.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_register -s 3 -c 1 -f -o gpurun_out/reg_r02_last \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-replay > gpurun_out/last_ncu_bench.log 2>&1; el "ncu full rc=$?"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_last.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-replay > gpurun_out/last_launches_bench.log 2>&1; el "launch list rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/last_bench_n1.json 2> gpurun_out/last_bench_n1.err; el "bench rc=$?"; tail -n 2 gpurun_out/last_bench_n1.err
timeout 100 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/last_bench_ref.json 2> gpurun_out/last_bench_ref.err; el "reference arm rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/last_bench_n1.json', 'gpurun_out/last_bench_ref.json'):
    try:
        l = json.loads([x for x in open(f).read().strip().split('\n') if x.startswith('{')][-1])
        print(f, round(l['value'], 1), 'scans/s', 'e2e', round(l['e2e']['value'], 1), 'replay', (l.get('replay') or {}).get('frames_per_s'))
    except Exception as e:
        print(f, 'parse failed', e)
PY
el done
