"""A/B of rollout_duo_kernel (two tiles per CTA taking turns on one hidden-1 buffer) against rollout_tmem_kernel on the bench
workload (32 states x K = 4096 x H = 20, 2 x 500 ReLU, uniform shooting) and a few other shapes: kernel time from the handle's
CUDA events, records compared bit for bit.  DUO_SHAPES="B,K,H;..." overrides the shapes."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bc_mpc_b200 import Planner  # noqa: E402
from oracle import mpc_oracle as O  # noqa: E402

S = int(os.environ.get("DUO_S", "500"))
shapes = [tuple(int(x) for x in sh.split(",")) for sh in os.environ.get("DUO_SHAPES", "32,4096,20;37,4096,20;256,4096,20;1,131072,20").split(";")]
rng, norm, weights, state = O.synthetic_problem(0, (S, S))
flop = 2 * (26 * S + S * S + S * 20)
for B, K, H in shapes:
    states = norm["mean_obs"] + norm["std_obs"] * np.random.default_rng(1).standard_normal((B, 20))
    out = {}
    for mode in ("0", "2"):
        os.environ["BCMPC_DUO"] = mode
        pl = Planner(20, 6, 2, S, "relu", num_paths=K, horizon=H, precision="bf16", max_states=B)
        pl.set_dynamics(weights)
        pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
        pl.enable_timing(True)
        ms = []
        for i in range(10):
            rec = pl.plan(states, seed=i)
            ms.append(pl.last_kernel_ms())
        out[mode] = (float(np.median(ms[3:])), rec)
        pl.close()
    same = bool(np.array_equal(out["0"][1]["index"], out["2"][1]["index"]) and np.array_equal(out["0"][1]["cost"], out["2"][1]["cost"]))
    for mode, name in (("0", "rollout_tmem_kernel"), ("2", "rollout_duo_kernel ")):
        ms = out[mode][0]
        print(f"B={B} K={K} H={H} S={S} {name}: {ms:.4f} ms  {B * K * H / ms / 1e6:.3f} G transitions/s  {B * K * H * flop / ms / 1e9:.0f} TFLOP/s", flush=True)
    print(f"    identical records: {same}   speed-up {out['0'][0] / out['2'][0]:.3f}x", flush=True)
