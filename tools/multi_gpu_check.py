"""Multi-GPU check of the product ABI's result gather (mpcc_cuda_comm_init / gather_results / read_gathered over NCCL), run as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/multi_gpu_check.py
Every rank owns a shard of the batch on its own GPU; after each cycle the library's all-gather must hand every rank the
results of ALL ranks, in rank order, identical to what torch.distributed gathers from the ranks' own host copies."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import torch.distributed as dist
import mpcc_manipulator_b200 as M
from mpcc_manipulator_b200.sharding import shard_range

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
TOTAL, N = 64 * world, 10
lo, hi = shard_range(TOTAL, rank, world)
B = hi - lo
rng = np.random.default_rng(0)
q_home = np.array([0, 0, 0, -np.pi / 2, 0, np.pi / 2, np.pi / 4])
x_all = np.tile(np.r_[q_home, 0., 0.], (TOTAL, 1)); x_all[:, :7] += rng.uniform(-0.05, 0.05, (TOTAL, 7))
mpc = M.BatchMPC(B, N, device=local); mpc.load_nn(); mpc.set_params(M.load_default_params())
ee = mpc.eval_robot_data(q_home[None])[0, 7:10]
mpc.set_tracks(M.load_track_json(None, ee))
uid = torch.from_numpy(M.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8)).cuda()
dist.broadcast(uid, 0)
mpc.comm_init(uid.cpu().numpy(), rank, world)
x, u = x_all[lo:hi].copy(), np.zeros((B, 8))
for c in range(4):
    r = mpc.run_cycle(x, u, want_horizon=False)
    mpc.gather_results()
    g = mpc.read_gathered()
    # reference gather through torch.distributed from the ranks' own results
    tu = torch.from_numpy(r["u0"]).cuda(); gu = [torch.empty_like(tu) for _ in range(world)]; dist.all_gather(gu, tu)
    ts = torch.from_numpy(np.stack([r["status"], r["iters"]], 1)).cuda(); gs = [torch.empty_like(ts) for _ in range(world)]; dist.all_gather(gs, ts)
    ref_u = torch.cat(gu).cpu().numpy(); ref_s = torch.cat(gs).cpu().numpy()
    assert np.array_equal(g["u0"], ref_u) and np.array_equal(g["status"], ref_s[:, 0]) and np.array_equal(g["iters"], ref_s[:, 1]), (rank, c)
    assert np.array_equal(g["u0"][lo:hi], r["u0"])
    u = r["u0"]; x = mpc.sim_time_step(r["x0"], u)
# two cycles in flight on the device-resident path, gathers enqueued back to back: the double buffer keeps them apart
dx = torch.from_numpy(x).cuda(); du = torch.from_numpy(u).cuda()
mpc.run_cycle_device(dx.data_ptr(), du.data_ptr()); mpc.gather_results()
mpc.run_cycle_device(dx.data_ptr(), du.data_ptr()); mpc.gather_results()
g2 = mpc.read_gathered(); loc = mpc.read_results()
assert np.array_equal(g2["u0"][lo:hi], loc["u0"])
mpc.close()
dist.barrier()
if rank == 0:
    print(f"multi-GPU gather ok: world={world}, {TOTAL} instances, 4 host cycles + 2 device cycles")
dist.destroy_process_group()
