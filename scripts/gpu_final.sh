#!/bin/bash
# round-end style verification on one B200: GPU tests, smoke, both bench arms, one ncu capture of the small-scan kernel
mkdir -p gpurun_out; rm -f gpurun_out/final.log
run() { echo "== $*" | tee -a gpurun_out/final.log; env "$@" 2>&1 | tail -${TAILN:-1} | tee -a gpurun_out/final.log; }
TAILN=12 run timeout 900 python -m pytest tests -m gpu -x -q --timeout 250
TAILN=3 run timeout 200 python -c "import __graft_entry__ as g; g.smoke()"
run timeout 400 python bench.py
run timeout 400 python bench.py --impl reference --steps 3 --warmup 1
run REPLAY_HOST=pinned timeout 200 python tests/replay_offline.py 60 64 2048 f32 fused
echo "== ncu --set full, k_assoc_group4 at cfg2 (2nd registration)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assoc -s 1 -c 1 -f -o gpurun_out/assoc_group4_cfg2 \
    python bench.py --workload 2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_group4.log 2>&1
ls -la gpurun_out/assoc_group4_cfg2.ncu-rep
