"""Launched by torchrun (one rank per GPU): sharded registration of cfg2 / cfg4 — every rank registers its contiguous index
range of the scan against a replicated map with one NCCL allreduce per iteration — and rank 0 checks the pose against
the CPU oracle.  Used by tests/test_multigpu.py and by hand:
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/sharded_check.py 2
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "kinematic-icp_b200", "python")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

import kinematic_icp_b200 as kb
from kinematic_icp_b200 import _capi
from oracle import kicp_oracle_py as ko
from oracle import workloads as W

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
if rank == 0:
    w = W.Workload(cfg)
dist.barrier()
if rank != 0:
    w = W.Workload(cfg)
ctx = kb.Context(local)
uid = torch.tensor(list(kb.comm_unique_id()), dtype=torch.uint8, device=dev) if rank == 0 else torch.empty(
    _capi.KICP_UNIQUE_ID_BYTES, dtype=torch.uint8, device=dev)
dist.broadcast(uid, 0)
mode = sys.argv[2] if len(sys.argv) > 2 else "p2p"
if mode == "nccl":
    ctx.comm_init(bytes(uid.cpu().tolist()), world, rank)
else:  # fused NVLink exchange: all-gather the CUDA-IPC handles of the mailboxes
    mine = torch.tensor(list(ctx.p2p_handle()), dtype=torch.uint8, device=dev)
    allh = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine)
    ctx.p2p_init([bytes(h.cpu().tolist()) for h in allh], world, rank)
gm = kb.VoxelHashMap(ctx, w.voxel_size, w.max_range, w.max_points_per_voxel)
gm.load_voxels(*w.map.export_voxels())
lo, hi = kb.shard_range(w.N, world, rank)
reg = kb.KinematicRegistration()
pose = reg.ComputeRobotMotionSharded(w.scan[lo:hi], gm, w.last_pose, w.rel_odom, w.tau)
res = reg.last_result
# every rank must hold the identical pose (the allreduced sums are identical, the solve is redundant)
t = torch.tensor(pose, dtype=torch.float64, device=dev)
gathered = [torch.empty_like(t) for _ in range(world)]
dist.all_gather(gathered, t)
ok = all(torch.equal(g, gathered[0]) for g in gathered)
if rank == 0:
    ref, st = w.map.register(w.scan, w.last_pose, w.rel_odom, w.tau, threads=os.cpu_count() or 1)
    dt, ang = ko.pose_delta(pose, ref)
    sums_ok = np.array_equal(res.sums_np()[:, 5], st.sums_np()[:, 5])
    print("SHARDED[%s] cfg%d world=%d iterations gpu=%d cpu=%d identical_on_all_ranks=%s N_match=%s pose delta %.3e m %.3e rad" %
          (mode, cfg, world, res.iterations, st.iterations, ok, sums_ok, dt, ang), flush=True)
    assert ok and sums_ok and res.iterations == st.iterations and dt <= 1e-6 and ang <= 1e-7
gm.close()
ctx.close()
dist.destroy_process_group()
