#!/bin/bash
# N GPUs: sharded parity (p2p + nccl) and the bench line with the replicas key
N=${1:-2}
mkdir -p gpurun_out
export KICP_SPIN_TIMEOUT_MS=20000
nvidia-smi -L | head -8
timeout 400 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -2
for mode in p2p nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/sharded_check.py 2 $mode 2>&1 | grep -E "SHARDED|Error|error|Traceback" | head -5
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 scripts/sharded_check.py 4 p2p 2>&1 | grep -E "SHARDED|Error|error|Traceback" | head -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc=$?"
tail -5 gpurun_out/r2_bench_n$N.err; python - $N <<'PY'
import json,sys
try:
    l=json.loads([x for x in open('gpurun_out/r2_bench_n%s.json'%sys.argv[1]).read().strip().split('\n') if x.startswith('{')][-1])
    print({k:l.get(k) for k in ('value','ms_per_step','n_gpus','cross_rank_identical','replicas')}, 'e2e', l['e2e']['value'])
    print('anatomy', l['pass_anatomy_us'])
    print('pose', l['pose_delta_vs_cpu'])
except Exception as e: print('parse failed', e)
PY
