#!/bin/bash
python -c "
import sys; sys.path.insert(0,'.')
from oracle import workloads as W
W.Workload(4); W.Workload(3); W.Workload(2); W.Workload(1)"
echo "== hostsorted x8"
for i in 1 2 3 4 5 6 7 8; do timeout 40 python scripts/repro_exp.py hostsorted 2>&1 | grep -c "ok" || echo HANG; done | tr '\n' ' '; echo
echo "== session x2"
for i in 1 2; do timeout 120 python scripts/repro_session.py 2>&1 | grep -c "ok iters" || echo HANG; done | tr '\n' ' '; echo
echo "== pytest"
timeout 600 python -m pytest tests -m gpu -x -q --timeout 150 2>&1 | tail -4
CFGS="4 3" VARIANTS="pruned:30 pruned:0" SKIP_TESTS=1 bash scripts/gpu_variants.sh
