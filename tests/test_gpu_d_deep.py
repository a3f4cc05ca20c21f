"""The cluster-split tcgen05 kernel (rollout_deep.cuh): hidden layers wider than 510 units or more than two of them -
BASELINE.json configs[4] is 26 -> 1024 -> 1024 -> 1024 -> 20, K = 2**20 paths x H = 30.  Accuracy against the oracle at
sizes it finishes in seconds, bit-exact self-consistency (batching, sharding, device sampling replay, production ==
debug flavour) and the full-size properties."""
import json
import os

import numpy as np
import pytest

from oracle import mpc_oracle as O
from tests._util import check_winner_replay, f32_uniform, make_pair

pytestmark = pytest.mark.gpu


def _norm_err(got, ref, norm):
    sd = np.where(norm["std_deltas"] > 0, norm["std_deltas"], 1.0)
    return (got - ref) / sd


@pytest.mark.parametrize("hidden", [(1024, 1024, 1024), (1024, 1024), (600, 600), (700, 700, 700), (1000, 1000, 1000, 1000),
                                    (100, 100, 100), (300, 300, 300), (384, 384, 384, 384)])
def test_deep_one_step_accuracy(hidden):
    """One predict step (dynamics.py:106-119) of every supported shape class: all 1..8 chunks per K-half, 2..4 hidden
    layers, ragged widths (zero-padded units), K not a multiple of the 256-path unit."""
    K = 1000
    rng, norm, state, pl, mpc = make_pair(3, hidden, K, 1, precision="bf16")
    actions = f32_uniform(rng, (1, K, 6))
    so, ao, co = pl.rollout_debug(state, actions)
    ref_s, ref_a, ref_c = mpc.rollout(state, actions)
    assert np.array_equal(ao, actions)
    e = _norm_err(so[1], ref_s[1], norm)            # error of the network output in normalised-delta units
    assert np.abs(e).max() < 0.1, np.abs(e).max()
    assert np.sqrt((e ** 2).mean()) < 0.02
    assert np.abs(so[0] - ref_s[0]).max() < 1e-5
    # predict entry point (per-row states) agrees with the rollout's first step bit for bit
    nxt = pl.predict(np.tile(state, (K, 1)), actions[0])
    assert np.array_equal(nxt, so[1])
    pl.close()
    # a few steps (barrier phases carry across steps and work units): tracks the oracle, plan == argmin of its own costs
    H = 4
    rng, norm, state, pl, mpc = make_pair(4, hidden, K, H, precision="bf16")
    actions = f32_uniform(rng, (H, K, 6))
    so, ao, co = pl.rollout_debug(state, actions)
    ref_s, ref_a, ref_c = mpc.rollout(state, actions)
    assert np.abs(_norm_err(so, ref_s, norm)).max() < 1.0
    assert np.corrcoef(co, ref_c)[0, 1] > 0.98
    rec = pl.plan(state, inj=actions)[0]
    assert rec["index"] == int(np.argmin(co)) and rec["cost"] == co[rec["index"]]
    prod = pl.plan(state, seed=11)[0]
    again = pl.plan(state, inj=pl.sample_injection(11)[0])[0]
    assert prod["index"] == again["index"] and prod["cost"] == again["cost"]


def test_deep_rollout_tracks_oracle_and_argmin_agreement():
    """configs[4]'s network at an oracle-sized problem: trajectories, costs, argmin agreement over seeded problems."""
    K, H, hidden = 512, 6, (1024, 1024, 1024)
    agree, regrets, worst = 0, [], 0.0
    seeds = range(8)
    for seed in seeds:
        rng, norm, state, pl, mpc = make_pair(seed, hidden, K, H, precision="bf16")
        actions = f32_uniform(rng, (H, K, 6))
        act, dbg = mpc.get_action(state, actions, return_debug=True)
        so, ao, co = pl.rollout_debug(state, actions)
        e = _norm_err(so, dbg["states"], norm)
        worst = max(worst, float(np.abs(e).max()))
        rec = pl.plan(state, inj=actions)[0]
        assert rec["index"] == int(np.argmin(co))
        assert np.array_equal(rec["action"][:6], actions[0, rec["index"]])
        agree += int(rec["index"] == dbg["argmin"])
        spread = dbg["costs"].std()
        regrets.append(float((dbg["costs"][rec["index"]] - dbg["costs"][dbg["argmin"]]) / spread))
        assert np.corrcoef(co, dbg["costs"])[0, 1] > 0.98
        pl.close()
    report = dict(net="26-1024-1024-1024-20", K=K, H=H, seeds=len(seeds), argmin_agreement=agree / len(seeds),
                  max_regret_in_cost_std=max(regrets), worst_state_err_in_delta_std=worst)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/bf16_agreement_deep_3x1024.json", "w") as f:
        json.dump(report, f)
    print(report)
    assert max(regrets) < 0.5 and worst < 1.0


def test_deep_production_flavour_batching_and_sharding_are_bit_consistent():
    """Device-sampled plans (production flavour) equal the plans replayed from the drawn tensor (debug flavour); a batch
    equals its single plans; the union of K-shards is the whole plan (Philox counter = global path index)."""
    K, H, B, hidden = 2048, 5, 3, (1024, 1024, 1024)
    rng, norm, state, pl, mpc = make_pair(9, hidden, K, H, precision="bf16", max_states=B)
    w = mpc.dyn_model.weights
    states = norm["mean_obs"] + norm["std_obs"] * rng.standard_normal((B, 20))
    prod = pl.plan(states, seed=77)
    inj = pl.sample_injection(77, B)
    dbg = pl.plan(states, inj=inj)
    assert np.array_equal(prod["index"], dbg["index"]) and np.array_equal(prod["cost"], dbg["cost"])
    assert np.array_equal(prod["action"], dbg["action"])
    assert not np.array_equal(pl.plan(states, seed=78)["action"], prod["action"])
    for b in range(B):
        _, _, _, p1, _ = make_pair(9, hidden, K, H, precision="bf16", weights=w)
        r = p1.plan(states[b], inj=inj[b])[0]
        assert r["index"] == prod[b]["index"] and r["cost"] == prod[b]["cost"]
        p1.close()
    want = make_pair(9, hidden, K, H, precision="bf16", weights=w)[3].plan(states[0], seed=5)[0]
    best = None
    bounds = [0, 300, 1111, K]
    for g in range(3):
        sp = make_pair(9, hidden, bounds[g + 1] - bounds[g], H, precision="bf16", weights=w, path_offset=bounds[g], total_paths=K)[3]
        r = sp.plan(states[0], seed=5)[0]
        if best is None or r["cost"] < best["cost"] or (r["cost"] == best["cost"] and r["index"] < best["index"]):
            best = r
        sp.close()
    assert best["index"] == want["index"] and best["cost"] == want["cost"] and np.array_equal(best["action"], want["action"])


def test_deep_full_size_properties():
    """BASELINE.json configs[4] at its own size, K = 2**20 x H = 30 on the tensor cores: the winner's recorded cost is
    reproduced by replaying that single path through the ORACLE with the device-drawn actions (delta-sigma metric), and
    K-shards (as on 8 GPUs) reduce to the same winner bit for bit."""
    import time
    K, H, hidden = 1 << 20, 30, (1024, 1024, 1024)
    rng, norm, weights, state = O.synthetic_problem(5, hidden)
    _, _, _, whole, _ = make_pair(5, hidden, K, H, precision="bf16", weights=weights)
    whole.plan(state, seed=98)
    t0 = time.perf_counter()
    full = whole.plan(state, seed=99)[0]
    dt = time.perf_counter() - t0
    print(f"configs[4] full size, cluster-split tcgen05 kernel: {dt * 1e3:.1f} ms per plan, {K * H / dt:.3e} transitions/s, "
          f"{K * H * 4288512 / dt / 1e12:.0f} TFLOP/s")
    whole.close()
    best, inj = None, None
    for g in range(8):
        _, _, _, shard, _ = make_pair(5, hidden, K // 8, H, precision="bf16", weights=weights, path_offset=g * (K // 8), total_paths=K)
        r = shard.plan(state, seed=99)[0]
        if g == int(full["index"]) // (K // 8):
            one = make_pair(5, hidden, 1, H, precision="fp32", weights=weights, path_offset=int(full["index"]), total_paths=K)[3]
            inj = one.sample_injection(99)[0]           # [H, 1, A]: the winner's action sequence
            one.close()
        if best is None or r["cost"] < best["cost"] or (r["cost"] == best["cost"] and r["index"] < best["index"]):
            best = r
        shard.close()
    assert best["index"] == full["index"] and best["cost"] == full["cost"] and np.array_equal(best["action"], full["action"])
    assert np.array_equal(full["action"][:6], inj[0, 0])
    one = make_pair(5, hidden, 1, H, precision="bf16", weights=weights, path_offset=int(full["index"]), total_paths=K)
    print("winner replay (worst delta-sigma, |cost diff|, threshold flips):",
          check_winner_replay(full, one[3], one[4], state, inj, norm, worst_sigma=1.5))


def test_deep_rejects_what_it_does_not_implement():
    from bc_mpc_b200 import Planner
    with pytest.raises(RuntimeError, match="2..4 hidden layers"):
        Planner(20, 6, 1, 1024, "relu", num_paths=256, horizon=2, precision="bf16")
    rng, norm, weights, state = O.synthetic_problem(5, (600, 600))
    pl = Planner(20, 6, 2, 600, "relu", num_paths=256, horizon=2, precision="bf16")
    pl.set_dynamics(weights)
    with pytest.raises(RuntimeError, match="n_hidden"):
        pl.set_layer_norm([(np.ones(600, np.float32), np.zeros(600, np.float32))] * 3)


@pytest.mark.parametrize("hidden", [(1024, 1024, 1024), (600, 600), (300, 300, 300, 300)])
def test_deep_tanh_dynamics(hidden):
    """tanh hidden layers - the reference builder's default activation (dynamics.py:60) - on the cluster-split kernel:
    same schedule, tanh.approx.bf16x2 in the layer-1 and owner epilogues instead of ReLU."""
    from bc_mpc_b200 import Planner
    K, H = 1000, 5
    rng, norm, weights, state = O.synthetic_problem(43, hidden)
    dyn = O.OracleDynamics(weights, norm, activation="tanh")
    mpc = O.OracleMPC(dyn, H, K)
    pl = Planner(20, 6, len(hidden), hidden[0], "tanh", num_paths=K, horizon=H, precision="bf16")
    pl.set_dynamics(weights)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
    actions = f32_uniform(rng, (H, K, 6))
    act, dbg = mpc.get_action(state, actions, return_debug=True)
    so, ao, co = pl.rollout_debug(state, actions)
    e = _norm_err(so, dbg["states"], norm)
    assert np.abs(e[1]).max() < 0.1 and np.abs(e).max() < 1.0, (np.abs(e[1]).max(), np.abs(e).max())
    assert np.corrcoef(co, dbg["costs"])[0, 1] > 0.98
    rec = pl.plan(state, inj=actions)[0]
    assert rec["index"] == int(np.argmin(co)) and rec["cost"] == co[rec["index"]]
    assert (dbg["costs"][rec["index"]] - dbg["costs"].min()) / dbg["costs"].std() < 0.5
    prod = pl.plan(state, seed=11)[0]
    again = pl.plan(state, inj=pl.sample_injection(11)[0])[0]
    assert prod["index"] == again["index"] and prod["cost"] == again["cost"]
    # and it is not the ReLU network
    pr = Planner(20, 6, len(hidden), hidden[0], "relu", num_paths=K, horizon=H, precision="bf16")
    pr.set_dynamics(weights)
    pr.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
    assert np.abs(pr.rollout_debug(state, actions)[0] - so).max() > 1e-3
    pl.close(); pr.close()


@pytest.mark.parametrize("activation", ["relu", "tanh"])
@pytest.mark.parametrize("hidden", [(1024, 1024, 1024), (600, 600), (300, 300, 300, 300), (100, 100, 100), (768, 768)])
def test_deep_layer_norm_dynamics(hidden, activation):
    """FLAGS.LAYER_NORM (the reference's default, dynamics.py:68-69) with more than two hidden layers or more than 510 units:
    the cluster-split kernel multiplies un-normalised ReLU outputs by diag(gamma) W and corrects in the owner's epilogue; a
    path's statistics are summed over both K-halves (8 bytes per path and layer through distributed shared memory)."""
    from bc_mpc_b200 import Planner
    K, H, S = 1000, 5, hidden[0]
    rng, norm, weights, state = O.synthetic_problem(47, hidden)
    lns = [(1.0 + 0.1 * rng.standard_normal(S).astype(np.float32), 0.1 * rng.standard_normal(S).astype(np.float32)) for _ in hidden]
    dyn = O.OracleDynamics(weights, norm, activation=activation, layer_norm=lns)
    mpc = O.OracleMPC(dyn, H, K)
    pl = Planner(20, 6, len(hidden), S, activation, num_paths=K, horizon=H, precision="bf16")
    pl.set_dynamics(weights)
    pl.set_layer_norm(lns)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
    actions = f32_uniform(rng, (H, K, 6))
    act, dbg = mpc.get_action(state, actions, return_debug=True)
    so, ao, co = pl.rollout_debug(state, actions)
    e = _norm_err(so, dbg["states"], norm)
    print("LayerNorm", hidden, activation, "one-step err", np.abs(e[1]).max(), "max err", np.abs(e).max())
    assert np.abs(e[1]).max() < 0.15 and np.abs(e).max() < 1.0, (np.abs(e[1]).max(), np.abs(e).max())
    assert np.corrcoef(co, dbg["costs"])[0, 1] > 0.98
    rec = pl.plan(state, inj=actions)[0]
    assert rec["index"] == int(np.argmin(co)) and rec["cost"] == co[rec["index"]]
    assert (dbg["costs"][rec["index"]] - dbg["costs"].min()) / dbg["costs"].std() < 0.5
    # production flavour == debug flavour, several work units per cluster, and determinism
    prod = pl.plan(state, seed=11)[0]
    again = pl.plan(state, inj=pl.sample_injection(11)[0])[0]
    assert prod["index"] == again["index"] and prod["cost"] == again["cost"]
    assert pl.plan(state, seed=11)[0]["cost"] == prod["cost"]
    # switching LayerNorm off again gives the plain network back
    pl.set_layer_norm(None)
    plain = pl.rollout_debug(state, actions)[0]
    ref_plain = O.OracleMPC(O.OracleDynamics(weights, norm, activation=activation), H, K).rollout(state, actions)[0]
    assert np.abs(_norm_err(plain, ref_plain, norm)).max() < 1.0 and np.abs(plain - so).max() > 1e-3
    pl.close()


def test_deep_layer_norm_many_units_and_batched_states():
    """More work units than clusters (barrier phases carry across units) and several root states: batched == one by one."""
    from bc_mpc_b200 import Planner
    hidden, K, H, B = (1024, 1024, 1024), 256 * 70, 3, 2
    S = hidden[0]
    rng, norm, weights, state = O.synthetic_problem(48, hidden)
    lns = [(1.0 + 0.1 * rng.standard_normal(S).astype(np.float32), 0.1 * rng.standard_normal(S).astype(np.float32)) for _ in hidden]
    pl = Planner(20, 6, 3, S, "relu", num_paths=K, horizon=H, max_states=B, precision="bf16")
    pl.set_dynamics(weights)
    pl.set_layer_norm(lns)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
    states = norm["mean_obs"] + norm["std_obs"] * rng.standard_normal((B, 20))
    both = pl.plan(states, seed=5)
    # (the Philox counter includes the root-state index, so a single-state plan replays state 0 only)
    one = pl.plan(states[:1], seed=5)
    assert both["index"][0] == one["index"][0] and both["cost"][0] == one["cost"][0]
    assert np.array_equal(pl.plan(states, seed=5)["cost"], both["cost"])
    # the winner's cost replayed through the oracle on the drawn actions
    inj = pl.sample_injection(5, B)
    dyn = O.OracleDynamics(weights, norm, layer_norm=lns)
    j = int(both["index"][1])
    s = states[1:2].copy()
    cost = 0.0
    for h in range(H):
        nxt = dyn.predict(s, inj[1, h, j:j + 1])
        cost += float(O.cheetah_cost_fn(s, inj[1, h, j:j + 1], nxt)[0])
        s = nxt
    assert abs(cost - both["cost"][1]) < 0.05 * abs(cost) + 5.0, (cost, both["cost"][1])
    pl.close()


@pytest.mark.parametrize("hidden,proposal,ln,activation", [((500, 500, 500), "policy_gaussian", False, "relu"), ((256, 256, 256), "policy_mix", True, "relu"),
                                                    ((600, 600), "policy_mix", False, "relu"), ((320, 320, 320, 320), "bc_uniform", False, "relu"),
                                                    ((640, 640, 640), "policy_gaussian", True, "relu"), ((200, 200, 200), "policy_gaussian", False, "relu"),
                                                    ((1024, 1024, 1024), "policy_gaussian", False, "relu"), ((1024, 1024, 1024), "policy_mix", True, "relu"),
                                                    ((768, 768), "bc_uniform", True, "tanh"), ((400, 400, 400), "policy_mix", False, "tanh"),
                                                    ((896, 896, 896, 896), "policy_gaussian", False, "tanh")])
def test_deep_policy_guided(hidden, proposal, ln, activation):
    """Policy-guided shooting (controllers.py:189-237, the legacy BC planner) on nets with more than two hidden layers (or 511..640
    units): the proposal phase of rollout_tmem_kernel at the top of every step of the cluster-split kernel, both K-half pairs
    computing the same actions from their bit-identical states."""
    from bc_mpc_b200 import Planner
    K, H, S = 1000, 5, hidden[0]
    rng, norm, weights, state = O.synthetic_problem(49, hidden)
    lns = [(1.0 + 0.1 * rng.standard_normal(S).astype(np.float32), 0.1 * rng.standard_normal(S).astype(np.float32)) for _ in hidden] if ln else None
    dyn = O.OracleDynamics(weights, norm, layer_norm=lns, activation=activation)
    pl = Planner(20, 6, len(hidden), S, activation, num_paths=K, horizon=H, precision="bf16", proposal=proposal, explore=0.4)
    pl.set_dynamics(weights)
    if ln:
        pl.set_layer_norm(lns)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
    pl.set_action_bounds(-np.ones(6), np.ones(6))
    if proposal == "bc_uniform":
        pw = O.glorot_mlp(rng, (20, 128, 64, 6), 0.05)
        pol = O.OracleBC(pw)
        pl.set_policy(pw)
    else:
        pw, logstd, rms_mean, rms_std = O.synthetic_policy(rng, hidden=(128, 96), norm=norm)
        pol = O.OraclePolicy(pw, logstd, rms_mean, rms_std)
        pl.set_policy(pw, logstd, rms_mean, rms_std, clip=5.0)
    mpc = O.OracleMPC(dyn, H, K, proposal=proposal, policy=pol, explore=0.4)
    inj = rng.standard_normal((H, K, 6)).astype(np.float32).astype(np.float64) if proposal == "policy_gaussian" else f32_uniform(rng, (H, K, 6))
    act, dbg = mpc.get_action(state, inj, return_debug=True)
    so, ao, co = pl.rollout_debug(state, inj)
    assert np.abs(ao[0] - dbg["actions"][0]).max() < 0.03           # first proposals: same state, bf16 policy net
    e = _norm_err(so, dbg["states"], norm)
    print("deep policy", hidden, proposal, ln, activation, "one-step", np.abs(e[1]).max(), "max", np.abs(e).max())
    assert np.abs(e[1]).max() < 0.15 and np.abs(e).max() < 1.0, (np.abs(e[1]).max(), np.abs(e).max())
    assert np.corrcoef(co, dbg["costs"])[0, 1] > 0.98
    rec = pl.plan(state, inj=inj)[0]
    assert rec["index"] == int(np.argmin(co)) and np.allclose(rec["action"][:6], ao[0, rec["index"]])
    assert (dbg["costs"][rec["index"]] - dbg["costs"].min()) / dbg["costs"].std() < 0.5
    prod = pl.plan(state, seed=11)[0]
    again = pl.plan(state, inj=pl.sample_injection(11)[0])[0]
    assert prod["index"] == again["index"] and prod["cost"] == again["cost"] and np.array_equal(prod["action"], again["action"])
    assert not np.array_equal(pl.plan(state, seed=12)[0]["action"], prod["action"])
    pl.close()


def test_deep_policy_guided_shapes_outside_the_fast_mode():
    from bc_mpc_b200 import Planner
    with pytest.raises(RuntimeError, match="129..1024"):
        Planner(20, 6, 3, 100, "relu", num_paths=256, horizon=2, precision="bf16", proposal="policy_mix")
    rng, norm, weights, state = O.synthetic_problem(5, (400, 400, 400))
    pl = Planner(20, 6, 3, 400, "relu", num_paths=256, horizon=2, precision="bf16", proposal="policy_mix")
    pw, logstd, rms_mean, rms_std = O.synthetic_policy(rng, hidden=(256, 64), norm=norm)
    with pytest.raises(RuntimeError, match="at most 128 units"):
        pl.set_policy(pw, logstd, rms_mean, rms_std, clip=5.0)
    pl.close()


@pytest.mark.parametrize("hidden,first_stage", [((300, 300, 300), "uniform"), ((768, 768), "policy_gaussian")])
def test_deep_two_stage_proposals(hidden, first_stage):
    """The tree-search planner's proposal schedule (controllers.py:409-416,429: a random or stochastic first-stage action, then
    the deterministic policy) on the cluster-split kernel: first_proposal switches the mode after step 0."""
    from bc_mpc_b200 import Planner
    K, H, S = 777, 5, hidden[0]
    rng, norm, weights, state = O.synthetic_problem(53, hidden)
    dyn = O.OracleDynamics(weights, norm)
    pw, logstd, rms_mean, rms_std = O.synthetic_policy(rng, hidden=(128, 128), norm=norm)
    pol = O.OraclePolicy(pw, logstd, rms_mean, rms_std)
    pl = Planner(20, 6, len(hidden), S, "relu", num_paths=K, horizon=H, precision="bf16", proposal="policy_mean", first_proposal=first_stage)
    pl.set_dynamics(weights)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
    pl.set_policy(pw, logstd, rms_mean, rms_std, clip=5.0)
    z = f32_uniform(rng, (K, 6)) if first_stage == "uniform" else rng.standard_normal((K, 6)).astype(np.float32).astype(np.float64)
    inj = np.zeros((H, K, 6))
    inj[0] = z
    # the oracle, step by step
    s, states, acts = np.tile(state, [K, 1]), [], []
    for h in range(H):
        a = z if (h == 0 and first_stage == "uniform") else pol.act(s, stochastic=(h == 0), noise=z)[0]
        states.append(s)
        acts.append(np.asarray(a, np.float64))
        s = dyn.predict(s, a)
    states.append(s)
    states, acts = np.asarray(states), np.asarray(acts)
    costs = O.trajectory_cost_fn(O.cheetah_cost_fn, states[:-1], acts, states[1:])
    so, ao, co = pl.rollout_debug(state, inj)
    assert np.abs(ao[0] - acts[0]).max() < (1e-6 if first_stage == "uniform" else 0.03)
    assert np.abs(ao[1:] - acts[1:]).max() < 0.05                        # deterministic policy on (almost) the same states
    e = _norm_err(so, states, norm)
    assert np.abs(e[1]).max() < 0.15 and np.abs(e).max() < 1.0, (np.abs(e[1]).max(), np.abs(e).max())
    assert np.corrcoef(co, costs)[0, 1] > 0.98
    rec = pl.plan(state, inj=inj)[0]
    assert rec["index"] == int(np.argmin(co)) and np.allclose(rec["action"][:6], ao[0, rec["index"]])
    assert (costs[rec["index"]] - costs.min()) / costs.std() < 0.5
    prod = pl.plan(state, seed=3)[0]
    again = pl.plan(state, inj=pl.sample_injection(3)[0])[0]
    assert prod["index"] == again["index"] and prod["cost"] == again["cost"] and np.array_equal(prod["action"], again["action"])
    pl.close()
