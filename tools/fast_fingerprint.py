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
    return out


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    fp = fingerprint()
    if mode == "save":
        np.savez_compressed(path, **fp)
        print("saved", len(fp), "arrays")
    else:
        ref = np.load(path)
        bad = [k for k in fp if not np.array_equal(fp[k], ref[k], equal_nan=True)]
        for k in bad:
            print("MISMATCH", k, float(np.nanmax(np.abs(fp[k] - ref[k]))))
        print("fingerprint", "DIFFERS" if bad else "identical", f"({len(fp)} arrays)")
        sys.exit(1 if bad else 0)
