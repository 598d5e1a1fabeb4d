#!/bin/bash
N=${1:-2}
python -c "
import sys; sys.path.insert(0,'.')
from oracle import workloads as W
W.Workload(4); W.Workload(2)"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== sharded parity cfg2 / cfg4 on $N GPUs"
for mode in p2p nccl; do
timeout 300 $TR --master-port 29541 scripts/sharded_check.py 2 $mode 2>&1 | grep -E "SHARDED|Error|rror" | head -5
timeout 300 $TR --master-port 29542 scripts/sharded_check.py 4 $mode 2>&1 | grep -E "SHARDED|Error|rror" | head -5
done
echo "== bench sharded nccl $N"
timeout 600 $TR --master-port 29545 bench.py --gpus $N --steps 20 --warmup 3 --comm nccl 2>&1 | grep "^{" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('value %.0f  ms/step %.3f  e2e %.0f'%(d['value'],d['ms_per_step'],d['e2e']['value']))"
echo "== bench sharded p2p $N"
timeout 600 $TR --master-port 29543 bench.py --gpus $N --steps 20 --warmup 3 2>&1 | grep "^{" | tee gpurun_out/bench_sharded_$N.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('value %.0f  ms/step %.3f  e2e %.0f  iters %d  split %s delta %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['config']['iterations_per_registration'],d['kernel_time_split_ms_per_step'],d['pose_delta_vs_cpu']))"
echo "== bench replicas $N"
timeout 600 $TR --master-port 29544 bench.py --gpus $N --steps 20 --warmup 3 --mode replicas 2>&1 | grep "^{" | tee gpurun_out/bench_replicas_$N.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('value %.0f  ms/step %.3f  e2e %.0f scaling %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['scaling']))"
