#!/bin/bash
# compute-sanitizer over the registration kernel (smoke-sized case through the public API) and then the GPU test suite
python -c "
import sys; sys.path.insert(0,'.')
from oracle import workloads as W
W.Workload(1)"
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  echo "== $tool"
  timeout 600 compute-sanitizer --tool $tool --print-limit 3 python scripts/debug_timing.py 1 2>&1 | grep -E "SUMMARY|Error|hazard" | head -4
done
timeout 400 python -m pytest tests -m gpu -x -q --timeout 150 2>&1 | tail -3
