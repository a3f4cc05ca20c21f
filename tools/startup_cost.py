import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
from bc_mpc_b200 import Planner
from oracle import mpc_oracle as O
rng, norm, weights, state = O.synthetic_problem(0, (500, 500))
for B in (4, 32):
    for H in (10, 20, 40):
        pl = Planner(20, 6, 2, 500, "relu", num_paths=4096, horizon=H, precision="bf16", max_states=B)
        pl.set_dynamics(weights)
        pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
        states = norm["mean_obs"] + norm["std_obs"] * rng.standard_normal((B, 20))
        pl.enable_timing(True)
        ms = []
        for i in range(8):
            pl.plan(states, seed=i); ms.append(pl.last_kernel_ms())
        print(f"B={B} H={H}: {np.median(ms[3:]) * 1e3:.1f} us")
        pl.close()
