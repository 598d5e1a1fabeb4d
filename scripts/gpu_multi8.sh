#!/bin/bash
N=${1:-8}
python -c "
import sys; sys.path.insert(0,'.')
from oracle import workloads as W
W.Workload(4)"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29542 scripts/sharded_check.py 4 p2p 2>&1 | grep -E "SHARDED|rror" | head -5
timeout 300 $TR --master-port 29543 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | grep "^{" | tee gpurun_out/bench_sharded_$N.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('sharded p2p: value %.0f  ms/step %.3f  e2e %.0f  iters %d'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['config']['iterations_per_registration']))"
timeout 300 $TR --master-port 29544 bench.py --gpus $N --steps 20 --warmup 3 --mode replicas --no-cpu-baseline 2>&1 | grep "^{" | tee gpurun_out/bench_replicas_$N.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('replicas: value %.0f  ms/step %.3f  e2e %.0f'%(d['value'],d['ms_per_step'],d['e2e']['value']))"
