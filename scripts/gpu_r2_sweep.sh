#!/bin/bash
# scheduling-option sweep on the default build and on the occupancy variants under kinematic-icp_b200/lib/ab/
mkdir -p gpurun_out
export KICP_SPIN_TIMEOUT_MS=5000
echo "== parity"; timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 200 2>&1 | tail -3 | tee gpurun_out/r2_parity.log
{
timeout 200 python -u scripts/ab_quick.py "4,3,2" "0,0;0,1;2,0;3,0;3,1;5,0"
for so in kinematic-icp_b200/lib/ab/libkicp_w*.so; do
  KICP_LIB=$PWD/$so timeout 150 python -u scripts/ab_quick.py "4,3" "0,0;0,1;3,0;3,1"
done
} 2>&1 | grep -v "^$" | tee gpurun_out/r2_sweep.log
