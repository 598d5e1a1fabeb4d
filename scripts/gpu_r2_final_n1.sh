#!/bin/bash
# round 2 evidence on one GPU: full GPU suite, the bench line, ncu launch list + one full-set capture of the registration kernel
mkdir -p gpurun_out
export KICP_SPIN_TIMEOUT_MS=20000
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r2_gputests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n1.err
for wl in 3 2 1; do timeout 300 python bench.py --steps 20 --warmup 5 --workload $wl --no-cpu-baseline > gpurun_out/r2_bench_cfg$wl.json 2>> gpurun_out/r2_bench_n1.err; done
# launch list (cold-cache, serialised: the kernel's SHARE of the step is what counts)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
# one full-set capture of a registration launch right after an L2 flush (the 4th k_register launch of the bench run)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_register -s 3 -c 1 -f -o gpurun_out/reg_r02_final python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_n1.json','gpurun_out/r2_bench_cfg3.json','gpurun_out/r2_bench_cfg2.json','gpurun_out/r2_bench_cfg1.json'):
    try:
        l=json.loads([x for x in open(f).read().strip().split('\n') if x.startswith('{')][-1])
        print(f, round(l['value']), 'scans/s', round(l['ms_per_step']*1e3),'us e2e', round(l['e2e']['value']), {k:round(v['value']) for k,v in l['e2e']['variants'].items()}, 'iters', l['config']['iterations_per_registration'])
    except Exception as e: print(f, 'parse failed', e)
PY
