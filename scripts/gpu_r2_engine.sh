#!/bin/bash
# ONE GPU call for the voxel-sorted engine (the round's last GPU minutes): parity of both engines, same-box A/B, an ncu capture of
# the new kernel and the bench line with it.  Every step has its own timeout; logs land in gpurun_out/.
mkdir -p gpurun_out
export KICP_SPIN_TIMEOUT_MS=20000
T0=$(date +%s)
el() { echo "[+$(( $(date +%s) - T0 ))s] $*"; }
# (a) the whole GPU suite, library defaults; the registration matrix covers engine 0 / 1 / 2 explicitly
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/eng_gputests_default.log 2>&1; el "gpu tests (default engine) rc=$?"; tail -3 gpurun_out/eng_gputests_default.log
# (b) same-box A/B of the two engines on configs 4, 3, 2 (flushed L2, resident frame, CUDA events), results checked against each other
timeout 240 python scripts/ab_quick.py "4,3,2" "engine=0;engine=2" > gpurun_out/eng_ab.log 2>&1; el "ab rc=$?"; grep "cfg" gpurun_out/eng_ab.log
# (c) the whole GPU suite again with every single-GPU persistent registration on the sorted engine (frame path and pipeline included)
KICP_ENGINE=2 timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/eng_gputests_sorted.log 2>&1; el "gpu tests (KICP_ENGINE=2) rc=$?"; tail -3 gpurun_out/eng_gputests_sorted.log
# (d) build variants of the sorted kernel (resident CTAs per SM)
for so in kinematic-icp_b200/lib/ab/*.so; do
  [ -f "$so" ] || continue
  KICP_LIB=$PWD/$so timeout 120 python scripts/ab_quick.py "4" "engine=2" > gpurun_out/eng_ab_$(basename $so .so).log 2>&1; el "$so rc=$?"; grep "cfg" gpurun_out/eng_ab_$(basename $so .so).log
done
# (e) one full-set capture of a sorted-engine launch right after an L2 flush, then the launch list
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_register_sorted -s 3 -c 1 -f -o gpurun_out/reg_r02_sorted \
    python bench.py --engine 1 --steps 2 --warmup 1 --no-cpu-baseline --no-replay > gpurun_out/eng_ncu_bench.log 2>&1; el "ncu full rc=$?"
# (f) the bench lines: sorted engine, then the default
timeout 300 python bench.py --engine 1 --steps 20 --warmup 5 > gpurun_out/eng_bench_sorted.json 2> gpurun_out/eng_bench_sorted.err; el "bench engine 1 rc=$?"; tail -2 gpurun_out/eng_bench_sorted.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_sorted.csv \
    python bench.py --engine 1 --steps 2 --warmup 1 --no-cpu-baseline --no-replay > gpurun_out/eng_launches_bench.log 2>&1; el "launch list rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-replay > gpurun_out/eng_bench_default.json 2> gpurun_out/eng_bench_default.err; el "bench default rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/eng_bench_sorted.json', 'gpurun_out/eng_bench_default.json'):
    try:
        l = json.loads([x for x in open(f).read().strip().split('\n') if x.startswith('{')][-1])
        print(f, l['config'].get('engine'), round(l['value']), 'scans/s', round(l['ms_per_step'] * 1e3), 'us  e2e', round(l['e2e']['value']),
              {k: round(v['value']) for k, v in l['e2e']['variants'].items()}, 'anatomy', [round(x, 1) for x in l['pass_anatomy_us']['median_over_passes_max_over_ranks']])
    except Exception as e:
        print(f, 'parse failed', e)
PY
el done
