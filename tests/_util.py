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
