#!/bin/bash
# round 2, first contact of the pooled registration kernel with the hardware: sanitizer on the smoke case, parity, bench, timings
mkdir -p gpurun_out
nvidia-smi -L
export KICP_SPIN_TIMEOUT_MS=5000
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/r2_sanitize.log 2>&1; echo "sanitize rc=$?"
tail -5 gpurun_out/r2_sanitize.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2_parity.log 2>&1; echo "parity rc=$?"
tail -15 gpurun_out/r2_parity.log
for c in 4 2 1; do timeout 120 python scripts/debug_timing.py $c; done 2>&1 | tee gpurun_out/r2_timing.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
cat gpurun_out/r2_bench.json | cut -c1-1500; tail -5 gpurun_out/r2_bench.err
