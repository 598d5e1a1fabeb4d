#!/bin/bash
python -c "
import sys; sys.path.insert(0,'.')
from oracle import workloads as W
W.Workload(1)"
for tool in ${TOOLS:-racecheck}; do
  for a in pruned group4; do
    echo "== $tool assoc=$a"
    KICP_ASSOC=$a timeout 300 compute-sanitizer --tool $tool --print-limit 3 python scripts/repro_variants.py 1 $([ $a = group4 ] && echo 2 || echo 1) 0 2>&1 | grep -E "SUMMARY|Error|hazard" | head -4
  done
done
timeout 400 python -m pytest tests -m gpu -x -q --timeout 150 2>&1 | tail -3
