#!/bin/bash
# single-synchronisation frame path: frame / pipeline / front-end tests, then the offline replay with the legacy order beside it;
# first the window timeline of the registration kernel (profile build)
mkdir -p gpurun_out
export KICP_SPIN_TIMEOUT_MS=5000
{
KICP_LIB=$PWD/kinematic-icp_b200/lib/ab/libkicp_prof.so timeout 240 python -u scripts/debug_timeline.py 4 flush
} 2>&1 | tee gpurun_out/r2_timeline.log
echo "== pytest gpu (frame + pipeline + front end)"; timeout 400 python -m pytest tests/test_gpu_frame.py tests/test_gpu_pipeline.py tests/test_gpu_frontend.py -m gpu -x -q --timeout 250 2>&1 | tail -15 | tee gpurun_out/r2_frame_tests.log
for fs in 0 1 0 1; do
  echo "== replay f32 fused, KICP_FRAME_SYNC=$fs"; KICP_FRAME_SYNC=$fs timeout 200 python tests/replay_offline.py 40 64 2048 f32 fused 2>&1 | tail -3 | tee -a gpurun_out/r2_replay_framesync.log
done
