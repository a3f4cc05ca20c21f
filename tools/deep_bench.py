#!/usr/bin/env python
"""Times the cluster-split tcgen05 kernel (rollout_deep.cuh) on BASELINE.json configs[4] or a scaled-down variant:
  python tools/deep_bench.py [--K 1048576] [--H 30] [--hidden 1024,1024,1024] [--plans 5]
Prints one JSON line (kernel ms from the handle's CUDA events, TFLOP/s on the algorithmic FLOPs)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--K", type=int, default=1 << 20)
    ap.add_argument("--H", type=int, default=30)
    ap.add_argument("--hidden", default="1024,1024,1024")
    ap.add_argument("--plans", type=int, default=5)
    ap.add_argument("--states", type=int, default=1)
    ap.add_argument("--ln", action="store_true", help="LayerNorm after every hidden activation (FLAGS.LAYER_NORM)")
    ap.add_argument("--proposal", default="uniform", help="uniform | policy_gaussian | policy_mix (proposal network 20-128-128-6)")
    ap.add_argument("--act", default="relu")
    args = ap.parse_args()
    from bc_mpc_b200 import Planner
    from oracle import mpc_oracle as O      # synthetic inputs only
    hidden = tuple(int(x) for x in args.hidden.split(","))
    rng, norm, weights, state = O.synthetic_problem(5, hidden)
    pl = Planner(20, 6, len(hidden), hidden[0], args.act, num_paths=args.K, horizon=args.H, max_states=args.states, precision="bf16",
                 proposal=args.proposal, explore=0.3)
    pl.set_dynamics(weights)
    pol_flop = 0
    if args.proposal != "uniform":
        pw, logstd, rms_mean, rms_std = O.synthetic_policy(rng, hidden=(128, 128), norm=norm)
        pl.set_policy(pw, logstd, rms_mean, rms_std, clip=5.0)
        pol_flop = 2 * (20 * 128 + 128 * 128 + 128 * 6)
    if args.ln:
        S = hidden[0]
        pl.set_layer_norm([(1.0 + 0.1 * rng.standard_normal(S).astype(np.float32), 0.1 * rng.standard_normal(S).astype(np.float32)) for _ in hidden])
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
    states = norm["mean_obs"] + norm["std_obs"] * rng.standard_normal((args.states, 20))
    pl.enable_timing(True)
    ms = []
    for i in range(args.plans + 1):
        pl.plan(states, seed=i)
        ms.append(pl.last_kernel_ms())
    ms = ms[1:]
    flop = 2 * sum(a * b for a, b in zip((26,) + hidden, hidden + (20,))) + pol_flop
    tr = args.states * args.K * args.H
    k = float(np.median(ms))
    print(json.dumps({"hidden": hidden, "proposal": args.proposal, "activation": args.act, "layer_norm": bool(args.ln), "K": args.K, "H": args.H, "states": args.states, "kernel_ms_median": k, "kernel_ms_min": min(ms),
                      "transitions_per_s": tr / (k * 1e-3), "tflops_algorithmic": tr * flop / (k * 1e-3) / 1e12,
                      "flop_per_transition": flop}))


if __name__ == "__main__":
    main()
