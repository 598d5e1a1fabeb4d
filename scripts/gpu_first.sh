#!/bin/bash
# first contact with the B200: smoke, GPU tests, short bench; logs under gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -20
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40
echo "== bench cfg2" ; timeout 600 python bench.py --workload 2 --steps 10 --warmup 3 2>&1 | tail -5
echo "== bench cfg4" ; timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tee gpurun_out/bench_cfg4.json | tail -5
