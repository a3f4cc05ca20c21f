"""Throughput of the fast kernel for several hidden widths (32 states x K=4096 x H=20, uniform shooting)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bc_mpc_b200 import Planner  # noqa: E402
from oracle import mpc_oracle as O  # noqa: E402

B, K, H = int(os.environ.get("SWEEP_B", "32")), 4096, 20
for S in [int(x) for x in os.environ.get("SWEEP_S", "64,128,256,300,384,500").split(",")]:
    rng, norm, weights, state = O.synthetic_problem(0, (S, S))
    pl = Planner(20, 6, 2, S, "relu", num_paths=K, horizon=H, precision="bf16", max_states=B)
    pl.set_dynamics(weights)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
    states = norm["mean_obs"] + norm["std_obs"] * rng.standard_normal((B, 20))
    pl.enable_timing(True)
    ms = []
    for i in range(8):
        pl.plan(states, seed=i)
        ms.append(pl.last_kernel_ms())
    ms = float(np.median(ms[3:]))
    flop = 2 * (26 * S + S * S + S * 20)
    print(f"S={S}: {ms:.3f} ms  {B * K * H / ms / 1e6:.3f} G transitions/s  {B * K * H * flop / ms / 1e9:.0f} TFLOP/s algorithmic")
    pl.close()
