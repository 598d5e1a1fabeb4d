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
L = _capi.lib()
L.kicp_debug_last_timing.argtypes = [C.c_void_p, _capi.c_dp]
for rep in range(3):
    reg.ComputeRobotMotion(w.scan, gm, w.last_pose, w.rel_odom, w.tau)
out = np.zeros((64, 4))
L.kicp_debug_last_timing(ctx.h, _capi.dp(out))
print("cfg", cfg, "iters", reg.last_result.iterations, "[windows_cta0, -, partial_sum, solve] us per iteration:")
print(np.round(out[:reg.last_result.iterations] / 1e3, 2))
