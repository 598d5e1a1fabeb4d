"""Timeline of one registration launch, window by window (needs a -DKR_PROFILE build: KICP_LIB=kinematic-icp_b200/lib/ab/libkicp_prof.so).
usage: python scripts/debug_timeline.py CFG [flush]   -> gpurun_out/r2_timeline_cfgN.npy + a printed summary"""
import sys, ctypes as C
sys.path.insert(0, "."); sys.path.insert(0, "kinematic-icp_b200/python")
import numpy as np
import kinematic_icp_b200 as kb
from kinematic_icp_b200 import _capi
from oracle import workloads as W
cfg = int(sys.argv[1])
w = W.Workload(cfg)
print("workload ready", flush=True)
ctx = kb.Context(0)
gm = kb.VoxelHashMap(ctx, w.voxel_size, w.max_range, w.max_points_per_voxel)
gm.load_voxels(*w.map.export_voxels())
reg = kb.KinematicRegistration()
scan = kb.Scan(ctx, w.N); scan.upload(w.scan)
res = kb.pinned_result()
for rep in range(3):
    reg.enqueue(scan, gm, w.last_pose, w.rel_odom, w.tau, res)
ctx.synchronize()
print("warm-up done, iterations", res.iterations, flush=True)
L = _capi.lib()
L.kicp_debug_window_log.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_int64, C.POINTER(C.c_int64)]
P = lambda x, q: np.percentile(x, q)


def capture(flush):
    if flush:
        import torch
        buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); buf.fill_(1); torch.cuda.synchronize()
    reg.enqueue(scan, gm, w.last_pose, w.rel_odom, w.tau, res)
    ctx.synchronize()
    cap = 65536
    out = np.zeros((cap, 4), dtype=np.uint64); n = C.c_int64(0)
    rc = L.kicp_debug_window_log(ctx.h, out.ctypes.data_as(C.POINTER(C.c_uint64)), cap, C.byref(n))
    assert rc == 0 and n.value > 0, (rc, n.value)
    log = out[: n.value]
    np.save("gpurun_out/r2_timeline_cfg%d%s.npy" % (cfg, "_flush" if flush else ""), log)
    t0 = log[:, 0].min()
    ts = (log[:, 0] - t0).astype(np.float64) / 1e3; te = (log[:, 1] - t0).astype(np.float64) / 1e3
    kind = (log[:, 2] >> np.uint64(56)).astype(int); it = ((log[:, 2] >> np.uint64(48)) & np.uint64(0xFF)).astype(int)
    sm = ((log[:, 2] >> np.uint64(32)) & np.uint64(0xFFFF)).astype(int); warp = (log[:, 2] & np.uint64(0xFFFFFFFF)).astype(int)
    a = (log[:, 3] >> np.uint64(32)).astype(np.int64); lines = (log[:, 3] & np.uint64(0xFFFFFFFF)).astype(np.int64)
    tasks = a >> 8; npts = a & 0xFF
    print("cfg", cfg, "flush" if flush else "warm", "iters", res.iterations, "records", n.value, "span %.1f us" % te.max(), flush=True)
    for p in range(int(it.max()) + 1):
        A = (kind == 1) & (it == p)
        if A.any():
            print("pass %d certificates: warps %d, end min/median/max %.1f / %.1f / %.1f us (from pass start %.1f)" %
                  (p, A.sum(), te[A].min(), np.median(te[A]), te[A].max(), ts[A].min()))
        Wn = (kind == 0) & (it == p)
        if not Wn.any(): continue
        d = te[Wn] - ts[Wn]
        print("pass %d search: windows %d  points/window %.1f tasks/window %.1f lines/window %.1f" % (p, Wn.sum(), npts[Wn].mean(), tasks[Wn].mean(), lines[Wn].mean()))
        print("   start %.1f .. last start %.1f ; end first %.1f  p50 %.1f  p90 %.1f  p99 %.1f  last %.1f us" %
              (ts[Wn].min(), ts[Wn].max(), te[Wn].min(), P(te[Wn], 50), P(te[Wn], 90), P(te[Wn], 99), te[Wn].max()))
        print("   window duration us: min %.1f p10 %.1f p50 %.1f p90 %.1f p99 %.1f max %.1f ; corr(duration, lines) %.2f corr(duration, tasks) %.2f" %
              (d.min(), P(d, 10), P(d, 50), P(d, 90), P(d, 99), d.max(), np.corrcoef(d, lines[Wn])[0, 1], np.corrcoef(d, tasks[Wn])[0, 1]))
        # per-warp finishing time (its last window's end) and windows per warp
        wl = warp[Wn]; order = np.argsort(wl, kind="stable")
        uw, cnt = np.unique(wl, return_counts=True)
        last_end = np.zeros(uw.size); np.maximum.at(last_end, np.searchsorted(uw, wl), te[Wn])
        print("   per warp: windows min %d max %d ; finish p10 %.1f p50 %.1f p90 %.1f max %.1f ; idle before the pass ends: mean %.1f us (%.0f%% of the phase)" %
              (cnt.min(), cnt.max(), P(last_end, 10), P(last_end, 50), P(last_end, 90), last_end.max(), (last_end.max() - last_end).mean(),
               100 * (last_end.max() - last_end).mean() / (last_end.max() - ts[Wn].min())))
        # the 10 last-finishing windows
        idx = np.where(Wn)[0][np.argsort(te[Wn])[-6:]]
        for i in idx:
            print("      late window: warp %5d sm %3d start %.1f end %.1f dur %.1f pts %d tasks %d lines %d" % (warp[i], sm[i], ts[i], te[i], te[i] - ts[i], npts[i], tasks[i], lines[i]))
        # per-SM mean window duration spread
        smd = np.zeros(160); smc = np.zeros(160); np.add.at(smd, sm[Wn], d); np.add.at(smc, sm[Wn], 1)
        m = smd[smc > 0] / smc[smc > 0]
        print("   per-SM mean window duration: min %.1f p50 %.1f max %.1f" % (m.min(), np.median(m), m.max()))


for fl in ([True, False] if len(sys.argv) > 2 else [False]):
    capture(fl)
