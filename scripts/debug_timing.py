"""Per-pass device timings (CTA 0, %globaltimer) and work counters of the registration kernel on one workload.
usage: python scripts/debug_timing.py <cfg 1..4> [ctas_per_sm]"""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "kinematic-icp_b200/python")
import numpy as np
import kinematic_icp_b200 as kb
from oracle import workloads as W
cfg = int(sys.argv[1])
w = W.Workload(cfg)
ctx = kb.Context(0)
ctx.set_option("nn_cache", 2)
if len(sys.argv) > 2:
    ctx.set_option("ctas_per_sm", int(sys.argv[2]))
gm = kb.VoxelHashMap(ctx, w.voxel_size, w.max_range, w.max_points_per_voxel)
gm.load_voxels(*w.map.export_voxels())
reg = kb.KinematicRegistration()
scan = kb.Scan(ctx, w.N)
scan.upload(w.scan)
res = kb.pinned_result()
for rep in range(5):
    reg.enqueue(scan, gm, w.last_pose, w.rel_odom, w.tau, res)
ctx.synchronize()
out = ctx.last_timing()
it = res.iterations
print("cfg", cfg, "N", w.N, "iters", it, "[certificates, their barrier, search, barrier wait, reduce(+exchange), solve] us per pass (CTA 0):")
print(np.round(out[:it] / 1e3, 2))
print("sum per pass us:", np.round(out[:it].sum(1) / 1e3, 2), "total us", round(out[:it].sum() / 1e3, 1))
ctx.set_option("stats", 1)
reg.enqueue(scan, gm, w.last_pose, w.rel_odom, w.tau, res)
probes, cands, lines, _ = ctx.last_stats()
print("per point per pass: probes %.2f candidates %.2f lines %.2f" % (probes / it / w.N, cands / it / w.N, lines / it / w.N))
# the same with the neighbour cache off (every point searched in every pass)
ctx.set_option("stats", 0)
ctx.set_option("nn_cache", 0)
for rep in range(3):
    reg.enqueue(scan, gm, w.last_pose, w.rel_odom, w.tau, res)
ctx.synchronize()
out = ctx.last_timing()
print("nn_cache=0: sum per pass us:", np.round(out[:it].sum(1) / 1e3, 2), "total us", round(out[:it].sum() / 1e3, 1))
