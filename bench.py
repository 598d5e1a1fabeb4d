#!/usr/bin/env python
"""bench.py — registrations/s of the kinematic-icp hot path on B200 (BASELINE.json metric).

A "step" is one full KinematicRegistration::ComputeRobotMotion (prior -> converged or max-iteration pose) of the
OS1-128-shape synthetic scan (~262 k points) against the 1 M-point voxel map (BASELINE.json configs[3], "cfg4").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload 1..4] [--mode sharded|replicas]

N > 1 is launched by torchrun, one rank per GPU.  In `sharded` mode (default, the north-star layout) the scan's
points are split by contiguous index range, the map is replicated, and every IRLS iteration ends with one NCCL
sum-allreduce of 8 doubles: total work is fixed, so "scaling" is "strong".  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "kinematic-icp_b200", "python")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "scans/sec (full ICP)"
UNIT = "scans/s"
L2_FLUSH_BYTES = 256 << 20  # > 126 MB L2


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", type=int, default=4, help="BASELINE.json config id 1..4 (default 4 = the quoted one)")
    ap.add_argument("--mode", default="sharded", choices=["sharded", "replicas"])
    ap.add_argument("--comm", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1 sharded mode: fused peer-memory exchange inside the persistent kernel (default) or NCCL allreduce")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="diagnostic only: keep L2 warm between steps")
    return ap.parse_args()


def workload_config(w, extra=None):
    d = w.describe()
    cfg = {"workload": "%s: %d-pt scan vs %d-pt voxel map (%d voxels), voxel_size %.2f, %d pts/voxel, tau %.4f, "
                       "max_iter 10, conv 1e-3, adaptive regularisation, prior = truth +5%% d +0.01 rad" %
                       (d["name"], d["N"], d["M"], d["voxels"], d["voxel_size"], d["max_points_per_voxel"], d["tau"]),
           "scan_points": d["N"], "map_points": d["M"], "map_voxels": d["voxels"]}
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


class NvmlSampler:
    """SM clock and throttle reasons sampled every millisecond by a thread, through NVML (nvidia_ml_py), only while the
    main thread is inside a timed region (`active`): the timed regions of this bench last milliseconds, shorter than one
    nvidia-smi start-up.  Falls back to the nvidia-smi loop above when NVML is unavailable."""
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, gpu_index, uuid=None):
        import threading
        self.fallback = None
        self.samples, self.active, self.stop_flag = [], False, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = None
            if uuid is not None:
                try:
                    self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
                except Exception:
                    self.h = None
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._run, daemon=True)
        except Exception:
            self.nv = None
            self.fallback = ClockSampler(gpu_index)

    def _reasons(self):
        for name in ("nvmlDeviceGetCurrentClocksEventReasons", "nvmlDeviceGetCurrentClocksThrottleReasons"):
            f = getattr(self.nv, name, None)
            if f is not None:
                try:
                    return int(f(self.h))
                except Exception:
                    continue
        return 0

    def _run(self):
        import time as _t
        while not self.stop_flag:
            if self.active:
                try:
                    self.samples.append((float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)), self._reasons()))
                except Exception:
                    pass
            _t.sleep(0.001)

    def start(self):
        if self.fallback is not None:
            self.fallback.start()
        else:
            self.thread.start()

    def stop(self):
        if self.fallback is not None:
            out = self.fallback.stop()
            out["source"] = "nvidia-smi -lms 100 over the whole run"
            return out
        self.stop_flag = True
        self.thread.join(timeout=2)
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": len(self.samples),
               "source": "NVML, 1 ms period, timed regions only"}
        if self.samples:
            out["sm_mhz"] = statistics.median(x[0] for x in self.samples)
            mask = 0
            for _, r in self.samples:
                mask |= r
            out["reasons"] = [name for bit, name in self.REASONS if mask & bit]
        return out


# -------------------------------------------------------------------------------------------------- CPU baseline
def cpu_registration_runner(w):
    """Returns (callable running one full registration on the host cores, kind, cores)."""
    from oracle import kicp_oracle_py as ko
    cores = os.cpu_count() or 1
    if ko.ref_available():
        _, _, pts = w.map.export_voxels()
        rm = ko.RefMap(w.voxel_size, w.max_range, w.max_points_per_voxel)
        rm.add_points(pts)

        def run():
            return rm.register(w.scan, w.last_pose, w.rel_odom, w.tau, threads=cores)
        return run, "reference", cores

    def run():
        return w.map.register(w.scan, w.last_pose, w.rel_odom, w.tau, threads=cores)[0]
    return run, "port", cores


def time_cpu(w, steps, warmup):
    run, kind, cores = cpu_registration_runner(w)
    for _ in range(warmup):
        run()
    ts = []
    for _ in range(steps):
        t = time.perf_counter()
        run()
        ts.append(time.perf_counter() - t)
    total = sum(ts)
    return {"value": steps / total, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d full registrations of the same workload (all %d scan points, all iterations), %s, "
                      "%d host threads, %.1f s of CPU work" %
                      (steps, w.N, "the reference's own Registration.cpp compiled against header shims (oracle/_ref)"
                       if kind == "reference" else "CPU oracle port", cores, total)}, total / steps


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import kicp_oracle_py as ko
    from oracle import workloads as W
    ko.build()
    w = W.Workload(args.workload)
    cb, sec_per_step = time_cpu(w, args.steps, args.warmup)
    line = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sec_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": workload_config(w), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------ GPU arm
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch

    import kinematic_icp_b200 as kb
    from kinematic_icp_b200 import _capi
    from oracle import kicp_oracle_py as ko
    from oracle import workloads as W

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun --nproc-per-node %d" % (args.gpus, args.gpus))
        args.gpus = world
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    ko.build()
    if rank == 0:
        w = W.Workload(args.workload)  # builds (and caches) once
    if world > 1:
        dist.barrier()
    if rank != 0:
        w = W.Workload(args.workload)

    ctx = kb.Context(local_rank)
    gm = kb.VoxelHashMap(ctx, w.voxel_size, w.max_range, w.max_points_per_voxel)
    gm.load_voxels(*w.map.export_voxels())  # replicated on every GPU
    reg = kb.KinematicRegistration()  # reference defaults: 10 iterations, 1e-3, adaptive regularisation

    sharded = world > 1 and args.mode == "sharded"
    if world > 1:
        if rank == 0:
            uid = torch.tensor(list(kb.comm_unique_id()), dtype=torch.uint8, device=dev)
        else:
            uid = torch.empty(_capi.KICP_UNIQUE_ID_BYTES, dtype=torch.uint8, device=dev)
        dist.broadcast(uid, 0)
        if args.comm == "nccl":
            ctx.comm_init(bytes(uid.cpu().tolist()), world, rank)
        else:  # fused exchange over NVLink peer memory: all-gather the CUDA-IPC handles of the mailboxes
            mine = torch.tensor(list(ctx.p2p_handle()), dtype=torch.uint8, device=dev)
            allh = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allh, mine)
            ctx.p2p_init([bytes(h.cpu().tolist()) for h in allh], world, rank)
    lo, hi = kb.shard_range(w.N, world, rank) if sharded else (0, w.N)
    shard = np.ascontiguousarray(w.scan[lo:hi])
    scan = kb.Scan(ctx, len(shard))
    scan.upload(shard)
    h_shard = kb.pinned_empty(shard.shape)
    h_shard[:] = shard

    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    flush_buf = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def flush_l2(i):
        if not args.no_flush:
            with torch.cuda.stream(stream):
                flush_buf.fill_(i & 0xFF)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    results = [kb.pinned_result() for _ in range(args.steps + args.warmup)]
    try:
        gpu_uuid = torch.cuda.get_device_properties(dev).uuid
    except Exception:
        gpu_uuid = None
    sampler = NvmlSampler(local_rank, gpu_uuid)
    if rank == 0:
        sampler.start()

    def timed_loop(enqueue, steps, warmup, profile=False):
        """W untimed warm-up steps, then K steps each bracketed by CUDA events on the launching stream, with an
        (untimed) L2 flush before every step.  Returns (sum of step ms as max over ranks, kernel profile)."""
        for i in range(warmup):
            flush_l2(i)
            enqueue(i)
        ctx.synchronize()
        barrier()
        if profile:
            ctx.profile_begin()
        sampler.active = True  # clocks are sampled only inside timed regions
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for i in range(steps):
            flush_l2(i)
            ev[i][0].record(stream)
            enqueue(warmup + i)
            ev[i][1].record(stream)
        prof = ctx.profile_end() if profile else None
        ctx.synchronize()
        sampler.active = False
        barrier()
        step_ms = [a.elapsed_time(b) for a, b in ev]
        return max_over_ranks(sum(step_ms)), prof, step_ms

    # ---- value: inputs already resident in HBM --------------------------------------------------------------
    def enqueue_resident(i):
        reg.enqueue(scan, gm, w.last_pose, w.rel_odom, w.tau, results[i], sharded=sharded)

    launches0 = ctx.launch_count
    total_ms, prof, step_ms = timed_loop(enqueue_resident, args.steps, args.warmup, profile=True)
    gpu_launches = ctx.launch_count - launches0
    # warm-up launches are included in launch_count; subtract them proportionally
    gpu_launches = int(round(gpu_launches * args.steps / float(args.steps + args.warmup)))
    jobs = world if (world > 1 and not sharded) else 1  # replicas: every rank finishes its own registrations
    value = jobs * args.steps / (total_ms * 1e-3)
    iters = results[args.warmup].iterations

    # ---- e2e: host buffers through the public synchronous call, copies inside the timed region ------------------
    out_pose = np.empty(7)
    res_host = kb.RegResult()
    params = reg._params()
    import ctypes as C
    fn = _capi.lib().kicp_register_sharded if sharded else _capi.lib().kicp_register

    def e2e_call(i):
        st = fn(gm.h, _capi.dp(h_shard), len(h_shard), _capi.dp(_capi.as_pose(w.last_pose)),
                _capi.dp(_capi.as_pose(w.rel_odom)), float(w.tau), C.byref(params), _capi.dp(out_pose), C.byref(res_host))
        assert st == 0, st

    e2e_ms, _, _ = timed_loop(e2e_call, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    e2e_value = jobs * args.steps / (e2e_ms * 1e-3)

    # ---- parity of what was just timed ------------------------------------------------------------------------
    pose_ref = pose_delta = None
    cpu_baseline = None
    if rank == 0:
        pose_ref, st_ref = w.map.register(w.scan, w.last_pose, w.rel_odom, w.tau, threads=os.cpu_count() or 1)
        dt, ang = ko.pose_delta(results[args.warmup].pose_np(), pose_ref)
        dt2, ang2 = ko.pose_delta(out_pose, pose_ref)
        pose_delta = {"translation_m": max(dt, dt2), "rotation_rad": max(ang, ang2), "iterations_gpu": int(iters),
                      "iterations_cpu": int(st_ref.iterations),
                      "tolerance": "1e-6 m / 1e-7 rad vs the CPU oracle (sequential-order FP64 restatement)"}
        if world == 1 and not args.no_cpu_baseline:
            cpu_baseline, _ = time_cpu(w, 5, 1)

    if rank == 0:
        # roofline of the association kernel: algorithmic bytes per point per launch
        cbar, kbar = w.map.neighbourhood_stats(w.scan, w.prior)
        a_pt = 16.0 + 27.0 * 16.0 + cbar * 16.0
        n_local = hi - lo
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
        else:
            peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        roofline = None
        if prof is not None and prof.assoc_launches > 0:
            # one launch of the association kernel = all IRLS iterations of one registration (persistent kernel) or one
            # iteration (multi-launch paths); algorithmic bytes per launch = passes per launch x N x A_pt
            passes_per_launch = prof.assoc_iterations / float(prof.assoc_launches)
            t_launch = prof.assoc_ms / prof.assoc_launches * 1e-3
            bytes_per_launch = passes_per_launch * n_local * a_pt
            achieved = bytes_per_launch / t_launch / 1e9
            traffic, traffic_note = None, None
            ncu_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
            if os.path.exists(ncu_path) and args.workload == 4 and world == 1:
                tj = json.load(open(ncu_path))
                traffic = tj.get("dram_bytes_per_launch")
                traffic_note = "dram__bytes_read+write per launch from %s" % tj.get("source")
            kernel_name = "k_register"
            roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": traffic, "traffic_note": traffic_note, "kernel": kernel_name,
                        "launch_us": t_launch * 1e6, "passes_per_launch": passes_per_launch,
                        "kernel_us": t_launch * 1e6 / max(passes_per_launch, 1e-9),
                        "launches_timed": int(prof.assoc_launches), "iterations_timed": int(prof.assoc_iterations),
                        "algorithmic_bytes_per_launch": bytes_per_launch,
                        "algorithmic_bytes_per_point": a_pt, "mean_candidates_per_point": cbar,
                        "mean_occupied_voxels_of_27": kbar, "peak_source": peak_src,
                        "note": "achieved = LOGICAL gather bytes (SURVEY.md 8(d): 16 + 27*16 + c*16 per point per pass) / "
                                "CUDA-event duration of the launch. The kernel prunes the 27-voxel neighbourhood exactly "
                                "(about 26 of the %.0f candidates per point are evaluated) and the map (%.0f MB) is served "
                                "from L2 after the first touch, so the logical figure exceeds what HBM carries: ncu "
                                "dram bytes per launch are in `traffic`. The kernel is latency/issue-bound, not "
                                "bandwidth-bound (profiles/)." %
                                (cbar, w.map.num_voxels() * w.max_points_per_voxel * 32 / 1e6)}
        ms_per_step = total_ms / args.steps
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if (world == 1 or sharded) else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(w, {
                "parallelism": ("1 GPU" if world == 1 else ("scan sharded by index range over %d GPUs, map replicated, "
                                "per-iteration exchange of 8 doubles: %s" % (world, "fused into the persistent kernel over NVLink peer memory"
                                if args.comm == "p2p" else "NCCL allreduce") if sharded else
                                "%d independent replicas" % world)),
                "l2": "flushed before every timed step (%d MiB write, untimed)" % (L2_FLUSH_BYTES >> 20)
                      if not args.no_flush else "NOT flushed (diagnostic run)",
                "iterations_per_registration": int(iters)}),
            "ms_per_iter": ms_per_step / max(int(iters), 1),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h_shard.nbytes) * (world if sharded else jobs),
                    "d2h_bytes_per_step": int(C.sizeof(kb.RegResult)) * world, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": gpu_launches,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "pose_delta_vs_cpu": pose_delta,
            "kernel_time_split_ms_per_step": None if prof is None else {
                "binning(init+keys+sort+gather)": prof.prep_ms / max(prof.registrations, 1),
                "association(active launches)": prof.assoc_ms / max(prof.registrations, 1),
                "launches_after_convergence": prof.idle_ms / max(prof.registrations, 1)},
        }
        print(json.dumps(line), flush=True)
    scan.close()
    gm.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
