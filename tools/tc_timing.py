"""Kernel-only timing of the tcgen05 rollout kernel at several grid sizes (separates per-CTA latency from
chip-wide contention).  Usage: python tools/tc_timing.py [S]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bc_mpc_b200 import Planner  # noqa: E402
from oracle import mpc_oracle as O  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 500
H = int(os.environ.get("TIMING_H", "20"))
rng, norm, weights, state = O.synthetic_problem(0, (S, S))
flop = 2 * (26 * S + S * S + S * 20)
for tiles in (1, 8, 37, 74, 148, 296, 1024):
    K = 128 * tiles
    pl = Planner(20, 6, 2, S, "relu", num_paths=K, horizon=H, precision="bf16")
    pl.set_dynamics(weights)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
    pl.enable_timing(True)
    ms = []
    for i in range(6):
        pl.plan(state, seed=i)
        ms.append(pl.last_kernel_ms())
    m = float(np.median(ms[2:]))
    waves = -(-tiles // 148)
    print(f"S={S} tiles={tiles:5d} H={H} kernel {m * 1e3:9.1f} us  per tile-step {m * 1e3 / (waves * H):7.2f} us  "
          f"{K * H / m / 1e6:8.1f} Mtr/s  {K * H * flop / m / 1e9:7.1f} TFLOP/s (cg {os.environ.get('BCMPC_TC_CG', '2')})")
    pl.close()
