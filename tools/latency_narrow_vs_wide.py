import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
from bc_mpc_b200 import Planner
from oracle import mpc_oracle as O
rng, norm, w, state = O.synthetic_problem(0, (256, 256))
pw, logstd, rms_mean, rms_std = O.synthetic_policy(rng, hidden=(128, 128), norm=norm)
ln = [(np.ones(256, np.float32), np.zeros(256, np.float32))] * 2
for K, H, prop, use_ln in ((400, 7, "policy_mix", True), (1000, 15, "uniform", False), (4096, 15, "uniform", False)):
    for mode in ("1", "0"):
        os.environ["BCMPC_NARROW"] = mode
        pl = Planner(20, 6, 2, 256, "relu", num_paths=K, horizon=H, precision="bf16", proposal=prop, explore=0.5)
        pl.set_dynamics(w)
        if use_ln: pl.set_layer_norm(ln)
        pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
        if prop != "uniform": pl.set_policy(pw, logstd, rms_mean, rms_std, clip=5.0)
        for i in range(20): pl.plan(state, seed=i)
        n = 300
        t0 = time.perf_counter()
        for i in range(n): pl.plan(state, seed=100 + i)
        dt = (time.perf_counter() - t0) / n
        print(f"K={K} H={H} {prop} ln={use_ln} {'narrow' if mode == '1' else 'wide  '}: {dt * 1e6:.1f} us per plan")
        pl.close()
