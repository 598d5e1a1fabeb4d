#!/bin/bash
# two GPUs: multi-GPU tests and the sharded bench line (fused NVLink exchange)
mkdir -p gpurun_out; rm -f gpurun_out/multi2.log
run() { echo "== $*" | tee -a gpurun_out/multi2.log; env "$@" 2>&1 | tail -${TAILN:-1} | tee -a gpurun_out/multi2.log; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
TAILN=6 run timeout 400 python -m pytest tests/test_multigpu.py -m gpu -x -q --timeout 200
run timeout 200 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline
