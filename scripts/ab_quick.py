"""Same-process sweep of the registration kernel's scheduling options on one library build (KICP_LIB selects it):
flushed-L2 timing of the resident path, CUDA events on the library's stream.
usage: python scripts/ab_quick.py "4,3" "opt=v,opt=v;opt=v;..."     (each ';'-separated entry is one configuration of kicp_ctx_set_option)"""
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "kinematic-icp_b200/python")
import numpy as np
import torch
import kinematic_icp_b200 as kb
from oracle import workloads as W
cfgs = [int(x) for x in sys.argv[1].split(",")]
combos = [[(kv.split("=")[0], int(kv.split("=")[1])) for kv in c.split(",") if kv] for c in sys.argv[2].split(";")]
allopts = sorted({k for c in combos for k, _ in c})
defaults = {k: int(os.environ.get("AB_DEFAULT_" + k.upper(), "0")) for k in allopts}
steps, warm = 20, 5
ctx = kb.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
print("lib", os.environ.get("KICP_LIB", "default"), flush=True)
for cfg in cfgs:
    w = W.Workload(cfg)
    gm = kb.VoxelHashMap(ctx, w.voxel_size, w.max_range, w.max_points_per_voxel)
    gm.load_voxels(*w.map.export_voxels())
    reg = kb.KinematicRegistration()
    scan = kb.Scan(ctx, w.N); scan.upload(w.scan)
    res = kb.pinned_result()
    ref = None
    for rep in range(2):
        for combo in combos:
            for k in allopts:
                ctx.set_option(k, dict(combo).get(k, defaults[k]))
            label = ",".join("%s=%d" % kv for kv in combo) or "defaults"
            for i in range(warm):
                with torch.cuda.stream(stream):
                    flush_buf.fill_(i)
                reg.enqueue(scan, gm, w.last_pose, w.rel_odom, w.tau, res)
            ctx.synchronize()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            for i in range(steps):
                with torch.cuda.stream(stream):
                    flush_buf.fill_(i)
                ev[i][0].record(stream)
                reg.enqueue(scan, gm, w.last_pose, w.rel_odom, w.tau, res)
                ev[i][1].record(stream)
            ctx.synchronize()
            us = np.array([a.elapsed_time(b) for a, b in ev]) * 1e3
            t = np.array(ctx.last_timing())[: res.iterations] / 1e3
            pose = np.array(res.pose)
            nsum = [res.sums[k][5] for k in range(res.iterations)]
            if ref is None:
                ref = (pose, nsum)
            ok = np.abs(pose - ref[0]).max() < 1e-9 and nsum == ref[1] and res.status == 0
            print("cfg %d %-28s : %.1f us (min %.1f) = %.0f scans/s | pass0 search+wait %.1f+%.1f  later: cert %.1f search %.1f wait %.1f | %s" %
                  (cfg, label, us.mean(), us.min(), 1e6 / us.mean(), t[0][2], t[0][3], np.median(t[1:, 0]) if len(t) > 1 else 0,
                   np.median(t[1:, 2]) if len(t) > 1 else 0, np.median(t[1:, 3]) if len(t) > 1 else 0, "ok" if ok else "MISMATCH"), flush=True)
    gm.close()
