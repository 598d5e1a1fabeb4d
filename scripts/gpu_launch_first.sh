#!/bin/bash
# A/B of the upload/launch order of the host-pointer entry point (kicp_register): e2e scans/s at cfg4 and cfg3
mkdir -p gpurun_out; rm -f gpurun_out/launch_first.log
run() { echo "== $*" | tee -a gpurun_out/launch_first.log; env "$@" 2>&1 | tail -${TAILN:-1} | tee -a gpurun_out/launch_first.log; }
TAILN=4 run timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 250
for lf in 0 1 0 1; do
  run KICP_LAUNCH_FIRST=$lf timeout 200 python bench.py --steps 60 --warmup 5 --no-cpu-baseline
done
for lf in 0 1; do
  run KICP_LAUNCH_FIRST=$lf timeout 200 python bench.py --workload 3 --steps 60 --warmup 5 --no-cpu-baseline
done
