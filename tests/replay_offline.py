"""Offline replay (BASELINE.json configs[4]): N independent synthetic 64-beam sequences, one per GPU, each driven frame
by frame through kinematic_icp::pipeline::KinematicICP::RegisterFrame of the C++ facade (front end, registration and
map update on the device), the way offline_node.cpp:99-149 replays a bag.  Prints aggregate frames/s as one JSON line,
writes the trajectories in TUM format, and checks rank 0 against the reference's own pipeline (oracle/_ref) when present.

    python tests/replay_offline.py [frames] [beams] [n_az] [f64|f32] [fused|staged]                       # 1 GPU
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 tests/replay_offline.py  # 8 independent sequences
"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "kinematic-icp_b200", "python")):
    sys.path.insert(0, p)
import numpy as np

from oracle import kicp_oracle_py as ko
from oracle import sequences as S

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 24
beams = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n_az = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
ingest = sys.argv[4] if len(sys.argv) > 4 else "f64"  # "f32": float32 x,y,z uploaded as is and widened on the device
# "fused": the facade's RegisterFrame = one kicp_register_frame call; "staged": the reference's own KinematicICP.cpp compiled over
# the facade classes (oracle/_ref/libkicp_ref_gpu.so) = one device call per stage with host round trips in between
variant = sys.argv[5] if len(sys.argv) > 5 else "fused"
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
os.environ["KICP_DEVICE"] = str(local)  # the facade's process-wide default context
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

seq = S.make_sequence(n_frames=frames, beams=beams, n_az=n_az, seed=4242 + 1000 * rank, deskew=True)
pipe = ko.facade_pipeline(deskew=True) if variant == "fused" else ko.ref_gpu_pipeline(deskew=True)
pipe.set_pose(seq["start"])
# warm-up frame (allocations, lazy module load), then reset
frames_in = [f.astype(np.float32) for f in seq["frames"]] if ingest == "f32" else seq["frames"]  # sequences are float32-representable
host = os.environ.get("REPLAY_HOST", "pinned")  # where the replay driver keeps the scans it feeds: page-locked or pageable memory
if host == "pinned":
    import kinematic_icp_b200 as kb  # kicp_host_alloc-backed numpy arrays
    def pin(a):
        b = kb.pinned_empty(a.shape, a.dtype)
        b[...] = a
        return b
    frames_in = [pin(f) for f in frames_in]
    stamps_in = [pin(np.asarray(t, dtype=np.float64)) for t in seq["stamps"]]
else:
    stamps_in = seq["stamps"]
step = pipe.register_frame_raw if variant == "fused" else pipe.register_frame  # raw: the caller's buffer, zero-copy
step(frames_in[0], stamps_in[0], seq["lidar_to_base"], seq["odoms"][0])
pipe.set_pose(seq["start"])
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
poses, stage_ms = [], []
KL = C.CDLL(os.path.join(ROOT, "tests", "hooks", "_build", "libkicp_facade_hooks.so"))
KL.kfac_debug_frame_timing.argtypes = [ko.c_dp]
tbuf = np.zeros(8)
for f, s, o in zip(frames_in, stamps_in, seq["odoms"]):
    p, _ = step(f, s, seq["lidar_to_base"], o)
    poses.append(p)
    if variant == "fused":
        KL.kfac_debug_frame_timing(tbuf.ctypes.data_as(ko.c_dp))
        stage_ms.append(np.diff(np.concatenate([[0.0], tbuf[:5]])))
elapsed = time.perf_counter() - t0
if world > 1:
    t = torch.tensor([elapsed], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed = float(t.item())
poses = np.array(poses)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
stamps = 0.1 * np.arange(1, frames + 1)
L = C.CDLL(os.path.join(ROOT, "tests", "hooks", "_build", "libkicp_facade_hooks.so"))
L.kfac_write_tum.argtypes = [C.c_char_p, ko.c_dp, ko.c_dp, C.c_int64]
tum = os.path.join(ROOT, "gpurun_out", "replay_rank%d.tum" % rank)
assert L.kfac_write_tum(tum.encode(), stamps.ctypes.data_as(ko.c_dp), np.ascontiguousarray(poses).ctypes.data_as(ko.c_dp), frames) == 0
if rank == 0 and stage_ms:
    np.savetxt(os.path.join(ROOT, "gpurun_out", "replay_stages_%s_%s.csv" % (ingest, host)), np.array(stage_ms), fmt="%.4f", delimiter=",",
               header="upload+front_end,enqueue_registration,registration,map_update,clouds (ms, per frame)")
if rank == 0:
    parity = None
    if ko.ref_available():
        ref = ko.ref_pipeline(deskew=True, max_num_threads=os.cpu_count() or 1)
        tr = time.perf_counter()
        rposes, _, _ = S.run_pipeline(ref, seq)
        cpu_fps = frames / (time.perf_counter() - tr)
        worst = [max(d) for d in zip(*[ko.pose_delta(a, b) for a, b in zip(poses, rposes)])]
        parity = {"translation_m": worst[0], "rotation_rad": worst[1], "cpu_reference_frames_per_s": cpu_fps,
                  "cpu_threads": os.cpu_count()}
    print(json.dumps({"metric": "offline replay: frames/s through KinematicICP::RegisterFrame (aggregate)",
                      "value": world * frames / elapsed, "unit": "frames/s", "n_gpus": world, "frames_per_sequence": frames, "ingest": ingest, "variant": variant, "host_buffers": host,
                      "points_per_frame": int(np.mean([len(f) for f in seq["frames"]])), "ms_per_frame": 1e3 * elapsed / frames,
                      "parity_rank0_vs_reference_pipeline": parity,
                      "host_stage_ms": dict(zip(["upload+front_end", "enqueue_registration", "registration", "map_update", "clouds"],
                                                np.round(np.mean(stage_ms, axis=0), 4).tolist())) if stage_ms else None, "tum_file": os.path.relpath(tum, ROOT)}), flush=True)
pipe.close()
if world > 1:
    dist.destroy_process_group()
