#!/bin/bash
# two GPUs: multi-GPU tests, offline replay with one sequence per GPU, sharded and replica bench lines
mkdir -p gpurun_out; rm -f gpurun_out/multi2.log
run() { echo "== $*" | tee -a gpurun_out/multi2.log; env "$@" 2>&1 | tail -${TAILN:-1} | tee -a gpurun_out/multi2.log; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
TAILN=6 run timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q --timeout 250
run REPLAY_HOST=pinned timeout 300 $TR tests/replay_offline.py 60 64 2048 f32 fused
run timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline
run timeout 300 $TR bench.py --gpus 2 --steps 30 --warmup 5
run timeout 300 $TR bench.py --gpus 2 --steps 30 --warmup 5 --mode replicas
