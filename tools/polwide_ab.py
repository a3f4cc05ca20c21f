"""A/B: BASELINE configs[3] (16 states x K=8192 x H=15, 2x500 dynamics, 128-unit proposal network) on the regular policy
variant (four 32 KB weight stages) and, forced with BCMPC_POL_WIDE=1, on the kVarPolWide variant (eight 8 KB stages):
what a small weight ring costs.  Also times the 256-unit network."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bc_mpc_b200 import Planner  # noqa: E402
from oracle import mpc_oracle as O  # noqa: E402  (input generation only)

B, K, H = 16, 8192, 15
rng, norm, weights, state = O.synthetic_problem(0, (500, 500))
states = norm["mean_obs"] + norm["std_obs"] * np.random.default_rng(2).standard_normal((B, 20))


def run(tag, hid, env):
    if env:
        os.environ["BCMPC_POL_WIDE"] = "1"
    else:
        os.environ.pop("BCMPC_POL_WIDE", None)
    pw, logstd, rms_mean, rms_std = O.synthetic_policy(np.random.default_rng(4), hidden=hid, norm=norm)
    pl = Planner(20, 6, 2, 500, "relu", num_paths=K, horizon=H, max_states=B, precision="bf16", proposal="policy_gaussian")
    pl.set_dynamics(weights)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
    pl.set_policy(pw, logstd, rms_mean, rms_std, clip=5.0)
    pl.enable_timing(True)
    ms = []
    for i in range(13):
        r = pl.plan(states, seed=i)
        ms.append(pl.last_kernel_ms())
    k = float(np.median(ms[3:]))
    print(f"{tag}: kernel {k:.4f} ms, {B * K * H / k / 1e3:.1f} M transitions/s, winner {int(r['index'][0])} cost {float(r['cost'][0]):.4f}")
    pl.close()


run("policy 128/128, regular variant (4 x 32 KB stages)", (128, 128), False)
run("policy 128/128, wide variant forced (8 x 8 KB stages)", (128, 128), True)
run("policy 256/256, wide variant", (256, 256), False)
