"""Single-plan latency of the host-buffer C-ABI call (configs[0] shape) split into kernel time and everything else."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bc_mpc_b200 import Planner  # noqa: E402
from oracle import mpc_oracle as O  # noqa: E402

rng, norm, weights, state = O.synthetic_problem(0, (500, 500))
for K, H in ((1000, 15), (1000, 1), (4096, 15)):
    pl = Planner(20, 6, 2, 500, "relu", num_paths=K, horizon=H, precision="bf16")
    pl.set_dynamics(weights)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
    pl.enable_timing(True)
    for i in range(20):
        pl.plan(state, seed=i)
    n = 200
    t0 = time.perf_counter()
    for i in range(n):
        pl.plan(state, seed=100 + i)
    wall = (time.perf_counter() - t0) / n
    k_ms = pl.last_kernel_ms()
    print(f"K={K} H={H}: wall {wall * 1e6:.1f} us per plan, rollout+reduce kernels {k_ms * 1e3:.1f} us, rest {wall * 1e6 - k_ms * 1e3:.1f} us")
    pl.close()
