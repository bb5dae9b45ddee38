"""CPU tier: the multi-GPU host logic (batch sharding + gather of results / counters) with world_size 2 and 3 on the
gloo backend.  No data-path collective exists on this path: instances are independent."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mpcc_manipulator_b200.sharding import shard_range, shard_sizes, gather_results, reduce_counters


def test_shard_ranges_cover_batch():
    for total in (1, 7, 64, 4096, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_range(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            for a, b in zip(blocks[:-1], blocks[1:]):
                assert a[1] == b[0]
            sizes = shard_sizes(total, world)
            assert sum(sizes) == total and max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    idx = torch.arange(lo, hi)
    u0 = (idx[:, None] * 10 + torch.arange(8)[None]).to(torch.float64)       # stands for the controls of this shard
    status = (idx % 3).to(torch.int32); iters = (idx % 5 + 1).to(torch.int32)
    gu, gs, gi = gather_results(u0, status, iters, total, dist)
    solved, failed, mx = reduce_counters(int((status == 0).sum()), int((status != 0).sum()), int(iters.max()), dist)
    if rank == 0:
        out.put((gu.numpy(), gs.numpy(), gi.numpy(), solved, failed, mx))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,total", [(2, 64), (3, 50)])
def test_gather_gloo(world, total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    gu, gs, gi, solved, failed, mx = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    idx = np.arange(total)
    assert np.array_equal(gu, idx[:, None] * 10.0 + np.arange(8)[None])
    assert np.array_equal(gs, idx % 3) and np.array_equal(gi, idx % 5 + 1)
    assert solved == int((idx % 3 == 0).sum()) and failed == total - solved and mx == 5


def _worker2(rank, world, port, b, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mpcc_manipulator_b200.sharding import ResultGatherer
    g = ResultGatherer(b, world, "cpu")
    idx = torch.arange(rank * b, (rank + 1) * b)
    for rep in range(2):  # buffers are reused across cycles
        gu, gs, gi = g((idx[:, None] * 10 + torch.arange(8)[None] + rep).to(torch.float64), (idx % 3).to(torch.int32), (idx % 5 + 1).to(torch.int32), dist)
    if rank == 0:
        out.put((gu.numpy().copy(), gs.numpy().copy(), gi.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_result_gatherer_gloo():
    world, b = 2, 16
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker2, args=(r, world, port, b, q)) for r in range(world)]
    for p in procs:
        p.start()
    gu, gs, gi = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    idx = np.arange(world * b)
    assert np.array_equal(gu, idx[:, None] * 10.0 + np.arange(8)[None] + 1)
    assert np.array_equal(gs, idx % 3) and np.array_equal(gi, idx % 5 + 1)


@pytest.mark.gpu
def test_abi_gather_over_nccl_on_two_gpus():
    """The product ABI's own gather (mpcc_cuda_gather_results, NCCL on a side stream) on two GPUs of one box; skipped on a
    single-GPU box (the CPU tier covers the sharding logic with gloo above)."""
    import subprocess, sys, torch
    from pathlib import Path
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29731",
                        str(root / "tools" / "multi_gpu_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "multi-GPU gather ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
