"""Randomised soak test of the fast-mode kernels: random shapes and variants, every plan run twice (determinism) and
replayed through the debug flavour (production == debug).  Usage: python tools/stress.py [seconds]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bc_mpc_b200 import Planner  # noqa: E402
from oracle import mpc_oracle as O  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(os.environ.get("STRESS_SEED", "1")))
t0, n, by_kind = time.time(), 0, {}
while time.time() - t0 < budget:
    S = int(rng.choice([7, 64, 100, 128, 200, 256, 300, 384, 500, 510]))
    kind = str(rng.choice(["uniform", "policy_gaussian", "policy_mix", "bc_uniform", "ln", "ln+policy", "reward", "reward+ln+policy", "reward+policy",
                           "tanh", "tanh+policy", "tanh+ln", "deep", "deep+tanh", "deep+ln", "deep+tanh+ln", "deep+policy", "deep+ln+policy", "deep+tanh+policy", "deep+tanh+ln+policy"]))
    rew, ln, deep = "reward" in kind, "ln" in kind, "deep" in kind
    act = "tanh" if (rew or "tanh" in kind) else "relu"
    if rew:
        S = int(rng.choice([100, 256, 330, 500]))
    hidden = (S, S)
    if deep:      # the cluster-split kernel: more than two hidden layers or more than 510 units
        S = int(rng.choice([100, 300, 600, 768, 1024]))
        if "policy" in kind:      # the proposal network's activations take 128 TMEM columns: at least 2 chunks per K-half
            S = int(rng.choice([200, 300, 500, 600, 768, 1024]))
        hidden = (S,) * int(rng.integers(3 if S <= 510 else 2, 5))
    prop = kind if kind in ("policy_gaussian", "policy_mix", "bc_uniform") else ("policy_gaussian" if "policy" in kind else "uniform")
    K, H, B = int(rng.integers(1, 3000)), int(rng.integers(1, 12)), int(rng.integers(1, 5))
    if rng.random() < 0.15:
        K, B = int(rng.integers(3000, 20000)), int(rng.integers(4, 40))       # multi-wave launches (narrow-net kernel where it applies)
    if deep:
        K = min(K, 4000)
        prng, norm, weights, state = O.synthetic_problem(int(rng.integers(1 << 30)), hidden)
        rweights = None
    else:
        prng, norm, weights, rweights, state = O.synthetic_reward_problem(int(rng.integers(1 << 30)), S)
    try:
        pl = Planner(20, 6, len(hidden), S, act, num_paths=K, horizon=H, max_states=B, precision="bf16", proposal=prop, explore=0.3)
        pl.set_dynamics(weights)
        pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
        lns = [(1.0 + 0.1 * prng.standard_normal(S).astype(np.float32), 0.1 * prng.standard_normal(S).astype(np.float32)) for _ in range(max(3, len(hidden)))]
        if ln:
            pl.set_layer_norm(lns[:len(hidden)])
        if rew:
            pl.set_reward_head(rweights, norm["mean_reward"], norm["std_reward"], 0.97, layer_norm=lns[2] if ln else None)
        if prop == "bc_uniform":
            pl.set_policy(O.glorot_mlp(prng, (20, 128, 64, 6), 0.05))
        elif prop != "uniform":
            pw, logstd, rms_mean, rms_std = O.synthetic_policy(prng, hidden=((int(rng.choice([32, 64, 128])), int(rng.choice([48, 128]))) if deep else (int(rng.choice([32, 64, 128, 129, 200, 256])), int(rng.choice([48, 128, 192, 256])))), norm=norm)
            pl.set_policy(pw, logstd, rms_mean, rms_std, clip=5.0)
    except RuntimeError as e:            # combinations without a kernel in this mode (e.g. BCMPC_TC_CG=1: reward heads + policy)
        if "weight ring" in str(e) or "hidden sizes" in str(e):
            continue
        raise
    states = norm["mean_obs"] + norm["std_obs"] * prng.standard_normal((B, 20))
    seed = int(rng.integers(1 << 40))
    try:
        a, b = pl.plan(states, seed=seed), pl.plan(states, seed=seed)
    except RuntimeError:
        print("FAILED configuration:", dict(kind=kind, hidden=hidden, K=K, H=H, B=B, seed=seed), flush=True)
        raise
    assert np.array_equal(a["index"], b["index"]) and np.array_equal(a["cost"], b["cost"], equal_nan=True), ("not deterministic", kind, S, K, H, B)
    if K * H * B <= 400000:
        c = pl.plan(states, inj=pl.sample_injection(seed, B))
        assert np.array_equal(a["index"], c["index"]) and np.array_equal(a["cost"], c["cost"], equal_nan=True), ("production != debug", kind, S, K, H, B)
    assert np.all((a["index"] >= 0) & (a["index"] < K)) and np.all(np.isfinite(a["cost"])), (kind, S, K, H, B, a)
    pl.close()
    n += 1
    by_kind[kind] = by_kind.get(kind, 0) + 1
print(f"stress ok: {n} random configurations in {time.time() - t0:.0f} s", by_kind)
