"""Bit-level fingerprint of the fast-mode planner over several network widths, with and without LayerNorm.
`python tools/fast_fingerprint.py save FILE` writes it, `... check FILE` compares bit for bit (used when the weight
packing code changes: the images, hence every cost, must not move)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bc_mpc_b200 import Planner  # noqa: E402
from oracle import mpc_oracle as O  # noqa: E402


def fingerprint():
    out = {}
    for S in (500, 510, 300, 256, 64, 7):
        for ln in (False, True):
            rng, norm, weights, state = O.synthetic_problem(100 + S, (S, S))
            pl = Planner(20, 6, 2, S, "relu", num_paths=512, horizon=4, precision="bf16")
            pl.set_dynamics(weights)
            pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"],
                                 norm["mean_deltas"], norm["std_deltas"])
            if ln:
                pl.set_layer_norm([(1.0 + 0.1 * rng.standard_normal(S).astype(np.float32),
                                    0.1 * rng.standard_normal(S).astype(np.float32)) for _ in range(2)])
            so, _, co = pl.rollout_debug(state, seed=9)
            out[f"s{S}_ln{int(ln)}_states"] = so
            out[f"s{S}_ln{int(ln)}_costs"] = co
    # policy-guided variants (proposal network images and biases), incl. the LayerNorm and learned-reward combinations
    for S, prop, rew, ln in ((500, "policy_gaussian", False, False), (300, "policy_mix", False, False), (64, "bc_uniform", False, False),
                             (256, "policy_gaussian", False, True), (500, "policy_mix", True, False), (256, "policy_gaussian", True, True)):
        rng, norm, weights, rweights, state = O.synthetic_reward_problem(200 + S, S)
        pl = Planner(20, 6, 2, S, "tanh" if rew else "relu", num_paths=512, horizon=4, precision="bf16", proposal=prop, explore=0.3)
        pl.set_dynamics(weights)
        pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"],
                             norm["mean_deltas"], norm["std_deltas"])
        lns = [(1.0 + 0.1 * rng.standard_normal(S).astype(np.float32), 0.1 * rng.standard_normal(S).astype(np.float32)) for _ in range(3)]
        if ln:
            pl.set_layer_norm(lns[:2])
        try:
            if rew:
                pl.set_reward_head(rweights, norm["mean_reward"], norm["std_reward"], 0.97, layer_norm=lns[2] if ln else None)
        except RuntimeError as e:          # single-CTA debug mode cannot hold the reward heads and a proposal network
            print("skipped", S, prop, rew, ln, "-", e)
            continue
        if prop == "bc_uniform":
            pl.set_policy(O.glorot_mlp(rng, (20, 128, 64, 6), 0.05))
        else:
            pw, logstd, rms_mean, rms_std = O.synthetic_policy(rng, hidden=(128, 96), norm=norm)
            pl.set_policy(pw, logstd, rms_mean, rms_std, clip=5.0)
        so, ao, co = pl.rollout_debug(state, seed=9)
        tag = f"pol_s{S}_{prop}_r{int(rew)}_ln{int(ln)}"
        out[tag + "_states"], out[tag + "_actions"], out[tag + "_costs"] = so, ao, co
    return out


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    fp = fingerprint()
    if mode == "save":
        np.savez_compressed(path, **fp)
        print("saved", len(fp), "arrays")
    else:
        ref = np.load(path)
        bad = [k for k in fp if k in ref.files and not np.array_equal(fp[k], ref[k], equal_nan=True)]
        print("compared", sum(k in ref.files for k in fp), "of", len(fp), "arrays")
        for k in bad:
            print("MISMATCH", k, float(np.nanmax(np.abs(fp[k] - ref[k]))))
        print("fingerprint", "DIFFERS" if bad else "identical", f"({len(fp)} arrays)")
        sys.exit(1 if bad else 0)
