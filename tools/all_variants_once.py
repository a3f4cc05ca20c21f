"""Small fast-mode and verification-mode plans, every kernel variant once (quick GPU smoke of a new build)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bc_mpc_b200 import Planner  # noqa: E402
from oracle import mpc_oracle as O  # noqa: E402

rng, norm, weights, state = O.synthetic_problem(1, (500, 500))
pw, logstd, rms_mean, rms_std = O.synthetic_policy(rng, hidden=(128, 128), norm=norm)
for precision in ("bf16", "fp32"):
    for proposal in ("uniform", "policy_gaussian"):
        for ln in (False, True):
            if ln and proposal != "uniform" and precision == "bf16":
                continue
            pl = Planner(20, 6, 2, 500, "relu", num_paths=300, horizon=3, max_states=2, precision=precision, proposal=proposal)
            pl.set_dynamics(weights)
            pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"],
                                 norm["mean_deltas"], norm["std_deltas"])
            if ln:
                pl.set_layer_norm([(np.ones(500, np.float32), np.zeros(500, np.float32))] * 2)
            if proposal != "uniform":
                pl.set_policy(pw, logstd, rms_mean, rms_std, clip=5.0)
            r = pl.plan(np.stack([state, state * 1.01]), seed=4)
            print(precision, proposal, "ln" if ln else "--", r["index"].tolist(), r["cost"].tolist())
            pl.close()
print("done")
