#!/bin/bash
# experiments: host buffer kind, ingest dtype, small-scan kernel choice; logs under gpurun_out/
mkdir -p gpurun_out; rm -f gpurun_out/replay2.log
run() { echo "== $*" | tee -a gpurun_out/replay2.log; env "$@" 2>&1 | tail -1 | tee -a gpurun_out/replay2.log; }
run REPLAY_HOST=pinned timeout 200 python tests/replay_offline.py 40 64 2048 f32 fused
run REPLAY_HOST=pinned timeout 200 python tests/replay_offline.py 40 64 2048 f64 fused
run REPLAY_HOST=pageable timeout 200 python tests/replay_offline.py 40 64 2048 f64 fused
run REPLAY_HOST=pageable timeout 200 python tests/replay_offline.py 40 64 2048 f32 fused
run REPLAY_HOST=pinned KICP_GROUP4_BELOW=65536 timeout 200 python tests/replay_offline.py 40 64 2048 f32 fused
for w in 1 2 3; do
  run KICP_GROUP4_BELOW=0 timeout 200 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline
  run KICP_GROUP4_BELOW=200000 timeout 200 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline
done
