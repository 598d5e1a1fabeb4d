#!/bin/bash
# option sweep on the default build (and on any variant under kinematic-icp_b200/lib/ab/); COMBOS / CFGS override the defaults
mkdir -p gpurun_out
export KICP_SPIN_TIMEOUT_MS=5000
COMBOS=${COMBOS:-"even_rounds=1;even_rounds=1,cert_margin=40;even_rounds=1,cert_margin=80;even_rounds=1,cert_margin=120;even_rounds=1,cert_margin=200"}
CFGS=${CFGS:-"4,3,2,1"}
if [ -n "$PARITY" ]; then echo "== parity"; timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 200 2>&1 | tail -3 | tee gpurun_out/r2_parity.log; fi
{
timeout 300 python -u scripts/ab_quick.py "$CFGS" "$COMBOS"
for so in kinematic-icp_b200/lib/ab/libkicp_*.so; do
  [ -f "$so" ] && KICP_LIB=$PWD/$so timeout 150 python -u scripts/ab_quick.py "4,3" "$COMBOS"
done
} 2>&1 | grep -v "^$" | tee gpurun_out/r2_sweep.log
