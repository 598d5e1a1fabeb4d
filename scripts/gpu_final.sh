#!/bin/bash
# round-end style verification on one B200: GPU tests, smoke, replay, default bench; logs under gpurun_out/
mkdir -p gpurun_out; rm -f gpurun_out/final.log
run() { echo "== $*" | tee -a gpurun_out/final.log; env "$@" 2>&1 | tail -${TAILN:-1} | tee -a gpurun_out/final.log; }
TAILN=12 run timeout 900 python -m pytest tests -m gpu -x -q --timeout 250
TAILN=3 run timeout 200 python -c "import __graft_entry__ as g; g.smoke()"
run REPLAY_HOST=pinned timeout 200 python tests/replay_offline.py 60 64 2048 f32 fused
run REPLAY_HOST=pinned timeout 200 python tests/replay_offline.py 60 64 2048 f64 fused
run REPLAY_HOST=pageable timeout 200 python tests/replay_offline.py 60 64 2048 f64 fused
run REPLAY_HOST=pageable timeout 200 python tests/replay_offline.py 60 64 2048 f64 staged
run timeout 400 python bench.py
for w in 1 2 3; do run timeout 200 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline; done
