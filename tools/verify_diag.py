"""Prints the fp32 verification-mode error figures the parity tests assert on (run on the GPU box)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mpc_oracle as O  # noqa: E402
from tests._util import f32_uniform, make_pair, rel_err  # noqa: E402


def report(tag, so, ref_s, co, ref_c, rec=None, dbg=None):
    e = rel_err(so, ref_s)
    i = np.unravel_index(np.argmax(e), e.shape)
    print(f"[{tag}] max rel err {e.max():.3e} at step/path/dim {i}: got {so[i]:.9g} want {ref_s[i]:.9g}; "
          f"per-step max {np.array2string(e.max(axis=(1, 2)), precision=1)}")
    print(f"    cost max abs err {np.abs(co - ref_c).max():.3e}; argmin gpu {np.argmin(co)} ref {np.argmin(ref_c)}; "
          f"best two ref costs {np.sort(ref_c)[:2]}")


g = np.load("tests/golden/vars_ppo_mpc.npz")
from bc_mpc_b200 import Planner  # noqa: E402
weights = [(g[f"dyn_w{i}"], g[f"dyn_b{i}"]) for i in range(3)]
norm = O.synthetic_normalization()
H, K = g["plan_actions"].shape[:2]
pl = Planner(20, 6, 2, 256, "relu", num_paths=K, horizon=H)
pl.set_dynamics(weights)
pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
so, ao, co = pl.rollout_debug(g["plan_state"], g["plan_actions"])
report("trained", so, g["plan_states"], co, g["plan_costs"])

rng, norm, state, pl, mpc = make_pair(0, (500, 500), 4096, 15)
actions = f32_uniform(rng, (15, 4096, 6))
act, dbg = mpc.get_action(state, actions, return_debug=True)
so, ao, co = pl.rollout_debug(state, actions)
report("c2 seed0", so, dbg["states"], co, dbg["costs"])
rec = pl.plan(state, inj=actions)[0]
print("    plan idx", rec["index"], "oracle", dbg["argmin"], "cost", rec["cost"], dbg["costs"][dbg["argmin"]])

for proposal, bc in (("policy_gaussian", False), ("policy_mix", False), ("bc_uniform", True)):
    K, H = 512, 6
    rng, norm, state, pl, mpc = make_pair(21, (500, 500), K, H, proposal=proposal, bc=bc)
    inj = rng.standard_normal((H, K, 6)).astype(np.float32).astype(np.float64) if proposal == "policy_gaussian" else f32_uniform(rng, (H, K, 6))
    act, dbg = mpc.get_action(state, inj, return_debug=True)
    so, ao, co = pl.rollout_debug(state, inj)
    ea = rel_err(ao, dbg["actions"])
    print(f"[{proposal}] action rel err per step {np.array2string(ea.max(axis=(1, 2)), precision=1)}")
    report(proposal, so, dbg["states"], co, dbg["costs"])
    n = 64
    S = norm["mean_obs"] + norm["std_obs"] * rng.standard_normal((n, 20))
    got = pl.policy_act(S)
    want = mpc.policy.act(S, False)[0] if not bc else mpc.policy.predict(S)
    print("    policy_act max abs err", np.abs(got - want).max(), "max |want|", np.abs(want).max())

rng, norm, state, pl, mpc = make_pair(5, (500, 500), 8, 2)
S = norm["mean_obs"] + norm["std_obs"] * rng.standard_normal((333, 20))
A = f32_uniform(rng, (333, 6))
e = rel_err(pl.predict(S, A), mpc.dyn_model.predict(S, A))
print("predict max rel err", e.max(), np.unravel_index(np.argmax(e), e.shape))
