#!/bin/bash
# fused RegisterFrame: GPU tests, then the offline replay (fused vs staged, f64 vs f32 ingest); logs under gpurun_out/
mkdir -p gpurun_out
echo "== pytest gpu (pipeline + frame)" ; timeout 600 python -m pytest tests/test_gpu_frame.py tests/test_gpu_pipeline.py tests/test_gpu_frontend.py -m gpu -x -q --timeout 250 2>&1 | tail -15
for v in "f64 fused" "f32 fused" "f64 staged"; do
  echo "== replay $v" ; timeout 200 python tests/replay_offline.py 40 64 2048 $v 2>&1 | tail -3 | tee -a gpurun_out/replay.log
done
echo "== ncu launch list of 6 fused frames"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/frame_launches.csv \
   python tests/replay_offline.py 6 64 2048 f32 fused > gpurun_out/frame_ncu.log 2>&1
tail -2 gpurun_out/frame_ncu.log
