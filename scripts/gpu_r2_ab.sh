#!/bin/bash
# A/B of launch-bounds / loads-in-flight builds of the registration kernel (kinematic-icp_b200/lib/ab/*.so, same ABI)
mkdir -p gpurun_out
export KICP_SPIN_TIMEOUT_MS=5000
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "registration" > gpurun_out/r2_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/r2_parity.log
{
echo "== default"; for c in 4 2 1; do timeout 120 python scripts/debug_timing.py $c; done
for so in kinematic-icp_b200/lib/ab/*.so; do echo "== $so"; for c in 4 2; do KICP_LIB=$PWD/$so timeout 120 python scripts/debug_timing.py $c | grep -E "total|sum per pass|per point"; done; done
} 2>&1 | tee gpurun_out/r2_ab.log
