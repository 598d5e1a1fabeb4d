#!/bin/bash
# native replay harness: pipeline tests (incl. kicp_replay against the golden drive), then the default bench line with its `replay` key
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_frame.py -m gpu -x -q --timeout 300 2>&1 | tail -4 | tee gpurun_out/r2_replay_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
l=json.loads([x for x in open('gpurun_out/r2_bench_n1.json').read().strip().split('\n') if x.startswith('{')][-1])
print(round(l['value']), 'scans/s e2e', round(l['e2e']['value']), 'traffic', l['roofline']['traffic'])
print('replay', l.get('replay'))
PY
