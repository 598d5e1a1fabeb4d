#!/bin/bash
# one development iteration on the GPU: registration parity subset, per-pass timings (default build and A/B builds), optional ncu
mkdir -p gpurun_out
export KICP_SPIN_TIMEOUT_MS=5000
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "registration" > gpurun_out/r2_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/r2_parity.log
{
for c in 4 3 2 1; do timeout 120 python scripts/debug_timing.py $c; done

} 2>&1 | tee gpurun_out/r2_timing.log
if [ -n "$1" ]; then bash scripts/gpu_r2_ncu.sh $1 4; fi
exit 0
