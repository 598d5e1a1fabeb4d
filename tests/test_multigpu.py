"""N > 1 path.  CPU (gloo, world_size 2): the host-side logic of the sharded registration — contiguous index ranges,
additivity of the per-shard normal-equation sums under a sum-allreduce of 8 doubles, out-of-band broadcast of the NCCL
unique id.  GPU (only when >= 2 GPUs are visible): the real thing through torchrun + NCCL."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "kinematic-icp_b200", "python"))
    import kinematic_icp_b200 as kb
    from oracle import kicp_oracle_py as ko
    from oracle import workloads as W
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = W.Workload(1)
    # 1. the unique id travels as a 128-byte tensor broadcast from rank 0 (here a stand-in payload; NCCL itself needs GPUs)
    uid = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
    dist.broadcast(uid, 0)
    assert bytes(uid.tolist()) == bytes(range(128))
    # 2. per-shard sums of one association at the prior, all-reduced: must equal the whole-scan sums
    lo, hi = kb.shard_range(w.N, world, rank)
    _, st = w.map.register(w.scan[lo:hi], w.last_pose, w.rel_odom, w.tau, max_iter=1)
    part = torch.zeros(8, dtype=torch.float64)
    part[:7] = torch.from_numpy(st.sums_np()[0])
    dist.all_reduce(part, op=dist.ReduceOp.SUM)
    _, full = w.map.register(w.scan, w.last_pose, w.rel_odom, w.tau, max_iter=1)
    f = full.sums_np()[0]
    assert part[5].item() == f[5]  # N is an exact integer count
    assert np.allclose(part[:7].numpy(), f, rtol=1e-12, atol=1e-12)
    # 3. every rank solving from the identical all-reduced sums gets the identical update (no broadcast of T needed)
    N, beta = part[5].item(), 1.0 / (part[6].item() / part[5].item() + np.finfo(np.float64).tiny)
    a, b, d = part[0].item() / N + beta, part[1].item() / N, part[2].item() / N
    r0, r1 = part[3].item() / N, part[4].item() / N
    inv = 1.0 / (a * d - b * b)
    dx = torch.tensor([-(d * inv * r0 - b * inv * r1), -(-b * inv * r0 + a * inv * r1)], dtype=torch.float64)
    both = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(both, dx)
    assert all(torch.equal(x, both[0]) for x in both)
    assert np.allclose(dx.numpy(), full.dx_np()[0], rtol=1e-9, atol=1e-12)
    dist.destroy_process_group()
    q.put(rank)


def test_sharded_logic_gloo_world2(oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert sorted(q.get() for _ in range(2)) == [0, 1]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["p2p", "nccl"])
@pytest.mark.parametrize("cfg", [2])
def test_sharded_registration(cfg, mode):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    n = min(torch.cuda.device_count(), 8)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
                        "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "scripts", "sharded_check.py"), str(cfg), mode],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARDED" in r.stdout
