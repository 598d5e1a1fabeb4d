"""Per-phase cycle split of the window loop (needs a -DKR_PROFILE build: KICP_LIB=kinematic-icp_b200/lib/ab/libkicp_prof.so)."""
import sys, ctypes as C
sys.path.insert(0, "."); sys.path.insert(0, "kinematic-icp_b200/python")
import numpy as np
import kinematic_icp_b200 as kb
from kinematic_icp_b200 import _capi
from oracle import workloads as W
cfg = int(sys.argv[1])
w = W.Workload(cfg)
ctx = kb.Context(0)
gm = kb.VoxelHashMap(ctx, w.voxel_size, w.max_range, w.max_points_per_voxel)
gm.load_voxels(*w.map.export_voxels())
reg = kb.KinematicRegistration()
scan = kb.Scan(ctx, w.N); scan.upload(w.scan)
res = kb.pinned_result()
for rep in range(3):
    reg.enqueue(scan, gm, w.last_pose, w.rel_odom, w.tau, res)
ctx.set_option("stats", 1)
reg.enqueue(scan, gm, w.last_pose, w.rel_odom, w.tau, res)
ctx.synchronize()
L = _capi.lib()
L.kicp_debug_last_prof.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
out = (C.c_uint64 * 24)()
L.kicp_debug_last_prof(ctx.h, out)
allp = [int(x) for x in out]
names = ["setup", "tasklist", "probe", "linemap", "decode+load", "math+reduce", "final", "next-ticket"]
print("cfg", cfg, "iters", res.iterations)
for label, p in (("pass 0 (every point searched)", allp[:12]), ("later passes (search of the uncertified points)", allp[12:])):
    nwin, nbatch, ngroup = p[8], p[9], p[10]
    tot = sum(p[:8])
    print(label, ": windows", nwin, "batches/window %.2f groups/window %.2f" % (nbatch / max(nwin, 1), ngroup / max(nwin, 1)))
    print("  cycles per window: total %.0f (%.1f us at 1.965 GHz)" % (tot / max(nwin, 1), tot / max(nwin, 1) / 1965.0))
    for k in range(8):
        print("    %-14s %8.0f cyc/window  %5.1f%%" % (names[k], p[k] / max(nwin, 1), 100.0 * p[k] / max(tot, 1)))
    print("    per batch: probe %.0f, linemap %.0f; per group: decode+load %.0f, math+reduce %.0f; outside windows per warp-pass %.0f cyc" %
          (p[2] / max(nbatch, 1), p[3] / max(nbatch, 1), p[4] / max(ngroup, 1), p[5] / max(ngroup, 1), p[11] / max(1, res.iterations)))
