#!/usr/bin/env python
"""CPU study for the FP8 question of VERDICT r1 item 8: would an e4m3 tensor-core mode (kind::f8f6f4: BOTH operands 8-bit,
fp32 accumulate, fp32 state and cost) keep the planner's arg-min?  Emulates it on the oracle's rollout: weights rounded to
e4m3 with one power-of-two scale per output column, hidden activations rounded to e4m3 with one power-of-two scale per layer
(set from the observed maximum), inputs likewise; everything else as the bf16 fast mode.  Reports arg-min agreement and regret
against the float32 oracle over 24 seeded states, next to the same emulation with bf16 operands.
Uses the oracle: developer tool, not product code.  Output: profiles/r2_fp8_feasibility.json"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mpc_oracle as O  # noqa: E402


def q_e4m3(x, scale):
    t = torch.from_numpy(np.ascontiguousarray(x / scale, dtype=np.float32))
    return t.clamp(-448, 448).to(torch.float8_e4m3fn).to(torch.float32).numpy() * scale


def q_bf16(x, scale=None):
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


def pow2_scale(amax, top):
    return 2.0 ** np.ceil(np.log2(np.maximum(amax, 1e-30) / top))


def rollout(weights, norm, state, actions, mode):
    H, K, A = actions.shape
    s = np.tile(state.astype(np.float32), (K, 1))
    cost = np.zeros(K, np.float32)
    q = {"fp8": q_e4m3, "bf16": q_bf16}.get(mode)
    Ws = []
    for w, b in weights:
        if mode == "fp8":
            sc = pow2_scale(np.abs(w).max(axis=0, keepdims=True), 448.0)
            Ws.append((q_e4m3(w, sc), b))
        elif mode == "bf16":
            Ws.append((q_bf16(w), b))
        else:
            Ws.append((w, b))
    for h in range(H):
        x = np.concatenate([(s - norm["mean_obs"]) / (norm["std_obs"] + 1e-10), (actions[h] - norm["mean_action"]) / (norm["std_action"] + 1e-10)], 1).astype(np.float32)
        for li, (w, b) in enumerate(Ws):
            if q is not None:
                x = q(x, pow2_scale(np.abs(x).max(), 448.0) if mode == "fp8" else None)
            x = x @ w + b
            if li + 1 < len(Ws):
                x = np.maximum(x, 0)
        delta = x * norm["std_deltas"] + norm["mean_deltas"]
        cost += 10.0 * (s[:, 5] >= 0.2) + 10.0 * (s[:, 6] >= 0) + 10.0 * (s[:, 7] >= 0) - delta[:, 17] * 100.0
        s = (s + delta).astype(np.float32)
    return cost


def study(tag, weights, norm, K, H, n, rng):
    out = {}
    rows = {"bf16": [], "fp8": []}
    for i in range(n):
        state = O.synthetic_state(rng, norm)
        actions = rng.uniform(-1, 1, (H, K, 6)).astype(np.float32)
        ref = rollout(weights, norm, state, actions, "f32")
        for mode in rows:
            c = rollout(weights, norm, state, actions, mode)
            w, best = int(np.argmin(c)), int(np.argmin(ref))
            rows[mode].append((w == best, float((ref[w] - ref[best]) / ref.std()), float(np.corrcoef(c, ref)[0, 1]), int((ref < ref[w]).sum())))
    for mode, r in rows.items():
        out[mode] = dict(argmin_agreement=float(np.mean([x[0] for x in r])), mean_regret_in_cost_std=float(np.mean([x[1] for x in r])),
                         max_regret_in_cost_std=float(np.max([x[1] for x in r])), min_cost_correlation=float(np.min([x[2] for x in r])),
                         median_rank_of_chosen_path=float(np.median([x[3] for x in r])), worst_rank_of_chosen_path=int(np.max([x[3] for x in r])))
    return {"config": tag, "K": K, "H": H, "seeds": n, **out}


def main():
    rng = np.random.default_rng(7)
    report = {"what": __doc__.split("\n\n")[0].replace("\n", " ")}
    _, norm, weights, _ = O.synthetic_problem(0, (500, 500))
    report["configs0_synthetic_2x500"] = study("BASELINE configs[0] shape, synthetic Glorot weights", weights, norm, 1000, 15, 24, rng)
    report["configs1_synthetic_2x500"] = study("BASELINE configs[1] shape, synthetic Glorot weights", weights, norm, 4096, 15, 24, rng)
    g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "vars_ppo_mpc.npz"))
    tw = [(g[f"dyn_w{i}"], g[f"dyn_b{i}"]) for i in range(3)]
    report["trained_2x256"] = study("saved_data/ppo-mpc/vars.pkl dynamics 26-256-256-20", tw, O.synthetic_normalization(), 1000, 15, 24, rng)
    print(json.dumps(report, indent=1))
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_fp8_feasibility.json"), "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
