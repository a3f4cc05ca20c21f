"""Diagnostics for the tcgen05 rollout kernel: structured weights that isolate each layer so a wrong
descriptor / swizzle shows up as a recognisable pattern.  Run on the GPU box."""
import sys
import os
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bc_mpc_b200 import Planner  # noqa: E402
from oracle import mpc_oracle as O  # noqa: E402


def run(weights, norm, state, actions, K, H, tag):
    pl = Planner(20, 6, 2, weights[0][0].shape[1], "relu", num_paths=K, horizon=H, precision="bf16")
    pl.set_dynamics(weights)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"],
                         norm["mean_deltas"], norm["std_deltas"])
    so, ao, co = pl.rollout_debug(state, actions)
    ref = O.OracleMPC(O.OracleDynamics(weights, norm), H, K).rollout(state, actions)
    sd = np.where(norm["std_deltas"] > 0, norm["std_deltas"], 1.0)
    e = (so - ref[0]) / sd
    print(f"[{tag}] K={K} H={H} max|err|/std_delta per step:", np.round(np.abs(e).max(axis=(1, 2)), 4))
    bad = np.argwhere(np.abs(e[1]) > 0.2)
    if len(bad):
        print(f"   step-1 bad entries: {len(bad)} of {e[1].size}; rows {sorted(set(bad[:, 0]))[:16]} dims {sorted(set(bad[:, 1]))}")
        r = bad[0][0]
        print("   got ", np.round(((so[1, r] - so[0, r]) / sd), 3))
        print("   want", np.round(((ref[0][1, r] - ref[0][0, r]) / sd), 3))
    print(f"   cost corr {np.corrcoef(co, ref[2])[0, 1]:.5f}  argmin gpu {np.argmin(co)} ref {np.argmin(ref[2])}")
    pl.close()
    return np.abs(e).max()


def main():
    S = int(os.environ.get("DIAG_S", "500"))
    K, H = 256, 2
    rng, norm, weights, state = O.synthetic_problem(0, (S, S))
    actions = rng.uniform(-1, 1, (H, K, 6)).astype(np.float32).astype(np.float64)
    # (1) identity-ish middle layer: y = relu(relu(xW1+b1)) W3 -> isolates L1 + A1 swizzle + L3
    W1, b1 = weights[0]
    W3, b3 = weights[2]
    W2 = np.eye(S, dtype=np.float32)
    run([(W1, b1), (W2, np.zeros(S, np.float32)), (W3, b3)], norm, state, actions, K, H, "W2=I")
    # (2) one-hot W1 (hidden j copies input j%26), identity W2, W3 picks hidden 0..19
    W1h = np.zeros((26, S), np.float32)
    for j in range(S):
        W1h[j % 26, j] = 1.0
    W3h = np.zeros((S, 20), np.float32)
    for d in range(20):
        W3h[26 * 3 + d, d] = 1.0
    run([(W1h, np.zeros(S, np.float32)), (W2, np.zeros(S, np.float32)), (W3h, np.zeros(20, np.float32))], norm, state,
        actions, K, H, "one-hot")
    # (3) random everything
    run(weights, norm, state, actions, K, H, "random")
    # (4) longer horizon, partial tile
    K2, H2 = 300, 15
    actions2 = rng.uniform(-1, 1, (H2, K2, 6)).astype(np.float32).astype(np.float64)
    run(weights, norm, state, actions2, K2, H2, "random-long")


if __name__ == "__main__":
    main()
