"""One batched plan of the bench workload (32 states x K=4096 x H=20, bf16) for ncu."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bc_mpc_b200 import Planner  # noqa: E402
from oracle import mpc_oracle as O  # noqa: E402

B = int(os.environ.get("PROF_B", "32"))
H = int(os.environ.get("PROF_H", "20"))
S = int(os.environ.get("PROF_S", "500"))
K = int(os.environ.get("PROF_K", "4096"))
proposal = os.environ.get("PROF_PROPOSAL", "uniform")      # e.g. policy_gaussian: BASELINE configs[3]'s kernel (PROF_B=1 PROF_K=8192 PROF_H=15)
rng, norm, weights, state = O.synthetic_problem(0, (S, S))
pl = Planner(20, 6, 2, S, "relu", num_paths=K, horizon=H, precision="bf16", max_states=B, proposal=proposal)
pl.set_dynamics(weights)
if proposal != "uniform":
    pw, logstd, rms_mean, rms_std = O.synthetic_policy(rng, hidden=(128, 128), norm=norm)
    pl.set_policy(pw, logstd, rms_mean, rms_std, clip=5.0)
pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
states = norm["mean_obs"] + norm["std_obs"] * rng.standard_normal((B, 20))
for i in range(3):
    rec = pl.plan(states, seed=i)
print("ok", rec[0]["index"], rec[0]["cost"])
