"""GPU check of the int8-split MLP kernel (k_mlp_oz, tcgen05) against the fp64 DMMA kernel (k_mlp) and the CPU oracle:
RobotData rows of both kernels on wide-range joint samples and obstacles, largest relative difference per block, and the
time of the robot-data launch pair of each kernel on the C2 sample count (4096 x 21)."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mpcc_manipulator_b200 as M  # noqa: E402


def main():
    import torch
    B, N = 4096, 20
    rng = np.random.default_rng(5)
    n = 4096
    q = rng.uniform(-2.8, 2.8, (n, 7))
    obs = np.c_[rng.uniform(-1, 1, (n, 3)), rng.uniform(0.02, 0.3, n)]
    out = {}
    import os
    for name, flags in (("dmma", 8),) + tuple((f"oz{f}", 16 + f) for f in map(int, os.environ.get("OZ_EXP", "0").split(","))):
        mpc = M.BatchMPC(B, N, flags=flags)
        mpc.load_nn()
        out[name] = mpc.eval_robot_data(q, obs)
        # timing: the launch pair (k_kin + MLP kernel) through the same entry, device time by CUDA events around repeated calls of the
        # smallest wrapper there is (eval_robot_data copies; so time the difference of 1 vs 5 calls is not clean) -> use stats from a cycle
        t = []
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            mpc.eval_robot_data(q, obs)
            torch.cuda.synchronize()
            t.append(time.perf_counter() - t0)
        print(f"{name}: eval_robot_data wall (n={n}, launches over {B * (N + 1)} samples) {min(t) * 1e3:.2f} ms")
        mpc.close()
    a, b = out["dmma"], out["oz0"]
    names = [("sel", 39 + 21 + 1 + 7, 1), ("dsel", 39 + 21 + 1 + 7 + 1, 7), ("env", 78 + 0, 9), ("denv", 87, 63)]
    print("max |dmma|:", np.abs(a).max(), " nan:", np.isnan(b).sum())
    # RobotData layout: q7|p3|R9|Jv21|Jw21|manip|dmanip7|sel|dsel7|obs_r|env9|denv63
    off = {"sel": 69, "dsel": 70, "obs_r": 77, "env": 78, "denv": 87}
    for k, (o, w) in {"sel": (69, 1), "dsel": (70, 7), "env": (78, 9), "denv": (87, 63)}.items():
        d = np.abs(a[:, o:o + w] - b[:, o:o + w])
        ref = np.abs(a[:, o:o + w]).max()
        i = np.unravel_index(np.argmax(d), d.shape)
        print(f"{k:5s}: max abs diff {d.max():.3e}  (scale {ref:.3e}, rel {d.max() / ref:.3e}) at sample {i[0]} col {i[1]}: dmma {a[i[0], o + i[1]]:.15e} oz {b[i[0], o + i[1]]:.15e}")
    other = np.abs(a[:, :69] - b[:, :69]).max()
    print("kinematics block identical:", other == 0.0)
    try:
        from oracle import oracle as O
        nn = O.OracleNN()
        m = 64
        ref = np.stack([nn.robot_data(q[i], obs[i]) for i in range(m)])
        for nm, x in (("dmma", a), ("oz0", b)):
            rel = np.abs(x[:m, 69:] - ref[:, 69:]).max() / np.abs(ref[:, 69:]).max()
            print(f"{nm} vs oracle (64 samples): rel {rel:.3e}")
    except Exception as ex:  # noqa: BLE001
        print("oracle comparison skipped:", ex)


if __name__ == "__main__":
    main()
