"""Helpers shared by the GPU parity tests: build a Planner and the matching oracle from one problem."""
import numpy as np

from oracle import mpc_oracle as O


def make_pair(seed, hidden=(500, 500), K=1000, H=15, precision="fp32", proposal="uniform", policy_hidden=(128, 128),
              explore=0.5, max_states=1, bias_scale=0.1, path_offset=0, total_paths=0, weights=None, bc=False):
    from bc_mpc_b200 import Planner
    rng, norm, w, state = O.synthetic_problem(seed, hidden, bias_scale=bias_scale)
    if weights is not None:
        w = weights
    dyn = O.OracleDynamics(w, norm)
    pl = Planner(20, 6, len(hidden), hidden[0], "relu", num_paths=K, horizon=H, max_states=max_states,
                 precision=precision, proposal=proposal, explore=explore, path_offset=path_offset,
                 total_paths=total_paths)
    pl.set_dynamics(w)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"],
                         norm["mean_deltas"], norm["std_deltas"])
    pl.set_action_bounds(-np.ones(6), np.ones(6))
    pol = None
    if proposal != "uniform":
        if bc:
            pw = O.glorot_mlp(rng, (20, 128, 64, 6), 0.05)
            pol = O.OracleBC(pw)
            pl.set_policy(pw)
        else:
            pw, logstd, rms_mean, rms_std = O.synthetic_policy(rng, hidden=policy_hidden, norm=norm)
            pol = O.OraclePolicy(pw, logstd, rms_mean, rms_std)
            pl.set_policy(pw, logstd, rms_mean, rms_std, clip=5.0)
    mpc = O.OracleMPC(dyn, H, K, proposal=proposal, policy=pol, explore=explore)
    return rng, norm, state, pl, mpc


def rel_err(got, ref):
    """per-entry error relative to the entry's magnitude (floored at 1e-3)"""
    return np.abs(got - ref) / (np.abs(ref) + 1e-3)


def state_scale(norm):
    """natural magnitude of every observation dimension: |mean_obs| + std_obs (idx 18 is identically 0)"""
    return np.abs(norm["mean_obs"]) + norm["std_obs"]


def scaled_rel_err(got, ref, scale):
    """per-entry error relative to max-ish(|entry|, magnitude of its dimension): |got-ref| / (|ref| + scale + 1e-3).
    Two float32 GEMMs that sum in different orders (numpy/OpenBLAS vs FFMA k-ascending vs TF/Eigen) differ by
    ~1e-6 of the DIMENSION's magnitude, so an entry that happens to pass near zero has an unbounded
    entry-relative error in any float32 implementation; the dimension scale is the meaningful denominator."""
    return np.abs(got - ref) / (np.abs(ref) + scale + 1e-3)


def f32_uniform(rng, shape):
    """U[-1,1) draws that are exactly representable in float32 (so every precision mode sees the same numbers)"""
    return rng.uniform(-1, 1, shape).astype(np.float32).astype(np.float64)


def check_winner_replay(rec, one_bf16, oracle_mpc, state, inj1, norm, worst_sigma=1.0):
    """Full-size property check of a bf16 plan whose K is far beyond what the oracle rolls in seconds: replay ONLY the
    winning path (its device-drawn action sequence `inj1` [H,1,A]) through (i) a one-path bf16 planner - must reproduce
    the recorded cost bit for bit - and (ii) the oracle: the bf16 trajectory stays within `worst_sigma` delta-sigmas of
    the oracle's at every step (the metric of the trajectory-tracking tests), and the cost difference is exactly what
    that drift explains: cost = sum of threshold penalties - (x17[H] - x17[0]) / 0.01 (cost_functions.py:16-28 summed over
    the steps), so |cost_bf16 - cost_oracle| <= 100 * |x17[H] error| + 10 * (threshold decisions that differ)."""
    so, ao, co = one_bf16.rollout_debug(state, inj1)
    assert co[0] == rec["cost"], (co[0], rec["cost"])
    ref_s, _, ref_c = oracle_mpc.rollout(state, inj1)
    sd = np.where(norm["std_deltas"] > 0, norm["std_deltas"], 1.0)
    e = np.abs((so - ref_s) / sd)
    assert e.max() < worst_sigma, e.max()
    flips = 0
    for d, thr in ((5, 0.2), (6, 0.0), (7, 0.0)):
        flips += int(np.sum((so[:-1, 0, d] >= thr) != (ref_s[:-1, 0, d] >= thr)))
    bound = 100.0 * abs(so[-1, 0, 17] - ref_s[-1, 0, 17]) + 10.0 * flips + 1e-2 * max(1.0, abs(ref_c[0]))
    assert abs(co[0] - ref_c[0]) <= bound, (co[0], ref_c[0], bound, flips)
    return float(e.max()), float(abs(co[0] - ref_c[0])), flips
