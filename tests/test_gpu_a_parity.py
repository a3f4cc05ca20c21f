"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes -> libbcmpc.so), against
the oracle on the same seeded inputs and against the golden fixtures made by the reference's own
controllers.py / cost_functions.py.  FP32 verification mode: per-step states within 1e-4 relative,
identical argmin and first action (BASELINE.json north_star)."""
import os

import numpy as np
import pytest

from oracle import mpc_oracle as O
from tests._util import f32_uniform, make_pair, rel_err, scaled_rel_err, state_scale

pytestmark = pytest.mark.gpu

TOL = 1e-4   # north_star: predicted states within 1e-4 relative per step


def _planner_for(weights, norm, K, H, **kw):
    from bc_mpc_b200 import Planner
    hidden = weights[0][0].shape[1]
    pl = Planner(20, 6, len(weights) - 1, hidden, "relu", num_paths=K, horizon=H, **kw)
    pl.set_dynamics(weights)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"],
                         norm["mean_deltas"], norm["std_deltas"])
    return pl


def test_golden_uniform_plans(golden_dir):
    g = np.load(os.path.join(golden_dir, "mpc_uniform.npz"))
    for tag in "abc":
        hidden = tuple(int(x) for x in g[f"{tag}_hidden"])
        rng, norm, weights, state = O.synthetic_problem(int(g[f"{tag}_seed"]), hidden)
        actions = g[f"{tag}_actions"]
        H, K = actions.shape[:2]
        pl = _planner_for(weights, norm, K, H)
        so, ao, co = pl.rollout_debug(state, actions)
        assert np.array_equal(ao, actions)
        assert scaled_rel_err(so, g[f"{tag}_states"], state_scale(norm)).max() < TOL
        assert np.abs(co - g[f"{tag}_costs"]).max() < 1e-3
        rec = pl.plan(state, inj=actions)[0]
        assert rec["index"] == int(g[f"{tag}_argmin"])
        assert np.array_equal(rec["action"][:6], g[f"{tag}_first_action"])


def test_golden_trained_weights(golden_dir):
    g = np.load(os.path.join(golden_dir, "vars_ppo_mpc.npz"))
    weights = [(g[f"dyn_w{i}"], g[f"dyn_b{i}"]) for i in range(3)]
    norm = O.synthetic_normalization()
    H, K = g["plan_actions"].shape[:2]
    pl = _planner_for(weights, norm, K, H)
    so, _, co = pl.rollout_debug(g["plan_state"], g["plan_actions"])
    assert scaled_rel_err(so, g["plan_states"], state_scale(norm)).max() < TOL
    rec = pl.plan(g["plan_state"], inj=g["plan_actions"])[0]
    assert rec["index"] == int(np.argmin(g["plan_costs"]))
    assert np.array_equal(rec["action"][:6], g["plan_first_action"])


def test_golden_policy_plans(golden_dir):
    g = np.load(os.path.join(golden_dir, "mpc_policy.npz"))
    hidden = tuple(int(x) for x in g["hidden"])
    H, K = g["eps"].shape[:2]
    for proposal, inj, want in (("policy_gaussian", g["eps"], g["selfexp_first_action"]),
                                ("policy_mix", g["uni"], g["mix_first_action"])):
        rng, norm, state, pl, mpc = make_pair(int(g["seed"]), hidden, K, H, proposal=proposal, policy_hidden=(32, 32),
                                              explore=float(g["explore"]), bias_scale=0.1)
        rec = pl.plan(state, inj=inj)[0]
        assert np.allclose(rec["action"][:6], want, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("K,H", [(1000, 15), (4096, 15)])   # BASELINE.json configs[0], configs[1]
def test_config_1_2_verify_mode(seed, K, H):
    rng, norm, state, pl, mpc = make_pair(seed, (500, 500), K, H)
    actions = f32_uniform(rng, (H, K, 6))
    act, dbg = mpc.get_action(state, actions, return_debug=True)
    so, ao, co = pl.rollout_debug(state, actions)
    err = scaled_rel_err(so, dbg["states"], state_scale(norm))
    assert err.max() < TOL, err.max()
    # one-step error with the drift removed: predict() fed the ORACLE's own trajectory states, entry-relative
    h = H // 2
    one = pl.predict(dbg["states"][h], actions[h])
    assert rel_err(one, dbg["states"][h + 1]).max() < TOL
    rec = pl.plan(state, inj=actions)[0]
    order = np.sort(dbg["costs"])
    assert rec["index"] == dbg["argmin"], (rec["index"], dbg["argmin"], order[:2])
    assert np.array_equal(rec["action"][:6], act)
    assert abs(rec["cost"] - dbg["costs"][dbg["argmin"]]) < 1e-3


def test_config_5_three_by_1024_net():
    """BASELINE.json configs[4]: dynamics MLP 3 x 1024 (verification kernel).  Oracle-size parity, then the
    size-independent properties at a larger K: device sampling replay, K-sharding with the min-loc reduce."""
    import time
    K, H = 512, 6
    rng, norm, state, pl, mpc = make_pair(5, (1024, 1024, 1024), K, H)
    actions = f32_uniform(rng, (H, K, 6))
    act, dbg = mpc.get_action(state, actions, return_debug=True)
    so, ao, co = pl.rollout_debug(state, actions)
    assert scaled_rel_err(so, dbg["states"], state_scale(norm)).max() < TOL
    assert np.abs(co - dbg["costs"]).max() < 1e-3
    rec = pl.plan(state, inj=actions)[0]
    assert rec["index"] == dbg["argmin"] and np.array_equal(rec["action"][:6], act)
    # larger K, sharded 4 ways (path_offset / total_paths as on 4 GPUs): the union of the shards is the single plan
    K2, H2 = 1 << 14, 30
    w = mpc.dyn_model.weights
    _, _, _, whole, _ = make_pair(5, (1024, 1024, 1024), K2, H2, weights=w)
    t0 = time.perf_counter()
    full = whole.plan(state, seed=77)[0]
    dt = time.perf_counter() - t0
    print(f"3x1024 verification kernel: {K2 * H2 / dt:.3e} transitions/s")
    best = None
    for g in range(4):
        _, _, _, shard, _ = make_pair(5, (1024, 1024, 1024), K2 // 4, H2, weights=w, path_offset=g * K2 // 4, total_paths=K2)
        r = shard.plan(state, seed=77)[0]
        if best is None or r["cost"] < best["cost"] or (r["cost"] == best["cost"] and r["index"] < best["index"]):
            best = r
        shard.close()
    assert best["index"] == full["index"] and best["cost"] == full["cost"] and np.array_equal(best["action"], full["action"])
    inj = whole.sample_injection(77)[0]
    j = int(full["index"])
    states = np.tile(state, [1, 1])
    cost = 0.0
    for h in range(H2):      # replay only the winning path through the oracle
        nxt = mpc.dyn_model.predict(states, inj[h, j:j + 1])
        cost += float(O.cheetah_cost_fn(states, inj[h, j:j + 1], nxt)[0])
        states = nxt
    assert abs(cost - full["cost"]) < 1e-3 * max(1.0, abs(cost))


def test_config_5_full_size_properties():
    """BASELINE.json configs[4] at its full size - K = 2**20 paths x H = 30, MLP 3 x 1024 - through size-independent
    properties: the union of 8 K-shards (as on 8 GPUs, min-loc over the shard records) is the single-handle plan, and
    the winning path replayed through the oracle reproduces its cost."""
    import time
    K, H = 1 << 20, 30
    rng, norm, weights, state = O.synthetic_problem(5, (1024, 1024, 1024))
    dyn = O.OracleDynamics(weights, norm)
    _, _, _, whole, _ = make_pair(5, (1024, 1024, 1024), K, H, weights=weights)
    t0 = time.perf_counter()
    full = whole.plan(state, seed=99)[0]
    dt = time.perf_counter() - t0
    print(f"configs[4] full size, verification kernel: {dt:.2f} s per plan, {K * H / dt:.3e} transitions/s")
    whole.close()
    best = None
    for g in range(8):
        _, _, _, shard, _ = make_pair(5, (1024, 1024, 1024), K // 8, H, weights=weights, path_offset=g * (K // 8), total_paths=K)
        r = shard.plan(state, seed=99)[0]
        if g == int(full["index"]) // (K // 8):
            inj = shard.sample_injection(99)[0]           # [H, K/8, A] of the shard that holds the winner
            assert r["index"] == full["index"]
        if best is None or r["cost"] < best["cost"] or (r["cost"] == best["cost"] and r["index"] < best["index"]):
            best = r
        shard.close()
    assert best["index"] == full["index"] and best["cost"] == full["cost"] and np.array_equal(best["action"], full["action"])
    j = int(full["index"]) % (K // 8)
    s, cost = state[None], 0.0
    for h in range(H):
        nxt = dyn.predict(s, inj[h, j:j + 1])
        cost += float(O.cheetah_cost_fn(s, inj[h, j:j + 1], nxt)[0])
        s = nxt
    assert abs(cost - full["cost"]) < 1e-3 * max(1.0, abs(cost))
    assert np.array_equal(full["action"][:6], inj[0, j])


def test_layer_norm_dynamics_variant():
    """FLAGS.LAYER_NORM = True, the default of train_mpc_ppo.py:52: LayerNorm after every hidden ReLU (dynamics.py:68-69)."""
    K, H, S = 1000, 15, 256
    rng, norm, state, pl, mpc = make_pair(17, (S, S), K, H)
    ln = [(1.0 + 0.1 * rng.standard_normal(S).astype(np.float32), 0.1 * rng.standard_normal(S).astype(np.float32)) for _ in range(2)]
    pl.set_layer_norm(ln)
    mpc.dyn_model.layer_norm = ln
    actions = f32_uniform(rng, (H, K, 6))
    act, dbg = mpc.get_action(state, actions, return_debug=True)
    so, ao, co = pl.rollout_debug(state, actions)
    assert scaled_rel_err(so, dbg["states"], state_scale(norm)).max() < TOL
    rec = pl.plan(state, inj=actions)[0]
    assert rec["index"] == dbg["argmin"] and np.array_equal(rec["action"][:6], act)
    pl.set_layer_norm(None)                      # switching it off gives the plain network again
    mpc.dyn_model.layer_norm = None
    so2, _, _ = pl.rollout_debug(state, actions)
    assert scaled_rel_err(so2, mpc.rollout(state, actions)[0], state_scale(norm)).max() < TOL
    assert np.abs(so2 - so).max() > 1e-3
    # fast mode: LayerNorm + proposal network runs at every hidden width in CTA-pair mode (the single-CTA developer mode,
    # BCMPC_TC_CG=1, has it for hidden <= 256 and 449..512 only and must say so)
    import os
    wide_ln = [(np.ones(320, np.float32), np.zeros(320, np.float32))] * 2
    _, _, st320, pl320, _ = make_pair(17, (320, 320), 128, 2, precision="bf16", proposal="policy_mix")
    if os.environ.get("BCMPC_TC_CG") == "1":
        with pytest.raises(RuntimeError, match="no tensor-core variant"):
            pl320.set_layer_norm(wide_ln)
    else:
        pl320.set_layer_norm(wide_ln)
        prod = pl320.plan(st320, seed=5)[0]
        replay = pl320.plan(st320, inj=pl320.sample_injection(5)[0])[0]
        assert prod["index"] == replay["index"] >= 0 and prod["cost"] == replay["cost"]


@pytest.mark.parametrize("tag", ["plain", "ln"])
def test_golden_reward_and_tree_search_plans(golden_dir, tag):
    """tests/golden/mpc_reward.npz: first actions chosen by the REFERENCE's MPCcontrollerReward,
    MPCcontrollerPolicyNetReward and MCTScontrollerPolicyNetReward; the drop-in classes must choose the same."""
    from bc_mpc_b200 import controllers, dynamics, policy
    from tests.test_oracle_golden import _reward_fixture

    class Space:
        def __init__(self, n):
            self.low, self.high, self.shape = -np.ones(n), np.ones(n), (n,)

    class Env:
        action_space, observation_space = Space(6), Space(20)

    g, odyn, opol, state = _reward_fixture(golden_dir, tag)
    S, ln = int(g["hidden"]), tag == "ln"
    H, K = g["actions"].shape[:2]
    N, R = int(g["N"]), int(g["R"])
    n10 = (odyn.mean_obs, odyn.std_obs, odyn.mean_action, odyn.std_action, odyn.mean_reward, odyn.std_reward, None, None,
           odyn.mean_deltas, odyn.std_deltas)
    dm = dynamics.NNDynamicsRewardModel(Env(), n10, size=S, layer_norm=ln)
    lns = [(g["ln_gamma"][i], g["ln_beta"][i]) for i in range(3)]
    dm.set_weights(odyn.weights, odyn.reward_weights, lns[:2] if ln else None, lns[2] if ln else None)
    pol = policy.MlpPolicy(None, Env(), 32, 2)
    pol.set_weights(opol.weights, opol.logstd, opol.rms_mean, opol.rms_std)
    for gamma in (1.0, 0.9):
        c = controllers.MPCcontrollerReward(Env(), dm, horizon=H, num_simulated_paths=K, gamma=gamma)
        c.sample_random_actions = lambda: g["actions"]
        assert np.array_equal(c.get_action(state), g[f"{tag}_uniform_g{gamma}_first_action"])
        if gamma == 0.9:
            _, _, co = c._planner.rollout_debug(state, g["actions"])
            assert np.abs(-co - g[f"{tag}_uniform_g0.9_rewards"]).max() < 1e-3
    c = controllers.MPCcontrollerPolicyNetReward(Env(), dm, pol, explore=0.4, self_exp=False, horizon=H, num_simulated_paths=K, gamma=0.5)
    c.sample_random_actions = lambda: g["actions"]
    assert np.allclose(c.get_action(state), g[f"{tag}_mix_first_action"], rtol=1e-5, atol=1e-6)
    c = controllers.MPCcontrollerPolicyNetReward(Env(), dm, pol, explore=0.4, self_exp=True, horizon=H, num_simulated_paths=K)
    c._ensure_planner()
    rec = c._planner.plan(state, inj=g["eps"].astype(np.float64))[0]           # the stochastic branch with the reference's own eps
    assert np.allclose(rec["action"][:6], g[f"{tag}_selfexp_first_action"], rtol=1e-5, atol=1e-6)
    for mode, z, key in (("uniform", g["first_uni"], "tree_uniform_action"), ("policy_gaussian", g["first_eps"], "tree_selfexp_action")):
        t = controllers.MCTScontrollerPolicyNetReward(Env(), dm, pol, self_exp=mode == "policy_gaussian", horizon=H, num_first_stage_actions=N,
                                                      random_path_per_action=R, random_first_stage_action=mode == "uniform")
        t.sample_first_stage = lambda z=z: np.asarray(z, np.float64)
        got = t.get_action(state)
        assert got.shape == (1, 6) and np.allclose(got, g[f"{tag}_{key}"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("proposal,gamma,ln", [("uniform", 0.9, False), ("uniform", 1.0, False), ("policy_mix", 1.0, False),
                                               ("uniform", 0.9, True), ("policy_gaussian", 1.0, True)])
def test_learned_reward_planners(proposal, gamma, ln):
    """NNDynamicsRewardModel + MPCcontrollerReward / MPCcontrollerPolicyNetReward (dynamics.py:121-238,
    controllers.py:90-158,289-363): shared tanh layer, delta head, reward head; argmax of the (discounted) reward sum."""
    from bc_mpc_b200 import Planner
    K, H, S = 1000, 10, 500
    rng, norm, weights, rweights, state = O.synthetic_reward_problem(29, S)
    dyn = O.OracleRewardDynamics(weights, rweights, norm)
    pl = Planner(20, 6, 2, S, "tanh", num_paths=K, horizon=H, precision="fp32", proposal=proposal, explore=0.4)
    pl.set_dynamics(weights)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
    lns = None
    if ln:        # FLAGS.LAYER_NORM (the default): LayerNorm after the shared layer and after each head's hidden layer
        lns = [(1.0 + 0.1 * rng.standard_normal(S).astype(np.float32), 0.1 * rng.standard_normal(S).astype(np.float32)) for _ in range(3)]
        dyn = O.OracleRewardDynamics(weights, rweights, norm, layer_norm=lns[:2], reward_layer_norm=lns[2])
        pl.set_layer_norm(lns[:2])
    pl.set_reward_head(rweights, norm["mean_reward"], norm["std_reward"], gamma, layer_norm=lns[2] if ln else None)
    pol = None
    if proposal != "uniform":
        pw, logstd, rms_mean, rms_std = O.synthetic_policy(rng, hidden=(128, 128), norm=norm)
        pol = O.OraclePolicy(pw, logstd, rms_mean, rms_std)
        pl.set_policy(pw, logstd, rms_mean, rms_std, clip=5.0)
    mpc = O.OracleMPC(dyn, H, K, proposal=proposal, policy=pol, explore=0.4, learned_reward=True, gamma=gamma)
    inj = rng.standard_normal((H, K, 6)).astype(np.float32).astype(np.float64) if proposal == "policy_gaussian" else f32_uniform(rng, (H, K, 6))
    act, dbg = mpc.get_action(state, inj, return_debug=True)
    so, ao, co = pl.rollout_debug(state, inj)
    assert scaled_rel_err(so, dbg["states"], state_scale(norm)).max() < TOL
    assert np.abs(co - dbg["costs"]).max() < 1e-3 * max(1.0, np.abs(dbg["costs"]).max())
    rec = pl.plan(state, inj=inj)[0]
    assert rec["index"] == dbg["argmin"]
    assert np.allclose(rec["action"][:6], act, rtol=1e-5, atol=1e-6)
    # single-step predict with reward (dynamics.py:221-238)
    n = 200
    Sx = norm["mean_obs"] + norm["std_obs"] * rng.standard_normal((n, 20))
    Ax = f32_uniform(rng, (n, 6))
    nxt, rew = pl.predict_reward(Sx, Ax)
    wn, wr = dyn.predict(Sx, Ax)
    assert scaled_rel_err(nxt, wn, state_scale(norm)).max() < 1e-5
    assert np.abs(rew - wr).max() < 1e-4
    pl.set_reward_head(None)          # back to cheetah_cost_fn
    _, _, co2 = pl.rollout_debug(state, inj)
    assert np.abs(co2 - co).max() > 1e-3


def test_learned_reward_controllers_drop_in():
    """The reference-shaped classes end to end: NNDynamicsRewardModel + MPCcontrollerReward /
    MPCcontrollerPolicyNetReward with sample_random_actions overridden (controllers.py:119,325)."""
    from bc_mpc_b200 import controllers, dynamics, policy

    class Space:
        def __init__(self, n):
            self.low, self.high, self.shape = -np.ones(n), np.ones(n), (n,)

    class Env:
        action_space, observation_space = Space(6), Space(20)

    K, H, S = 300, 6, 500
    rng, norm, weights, rweights, state = O.synthetic_reward_problem(31, S)
    n10 = (norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_reward"],
           norm["std_reward"], None, None, norm["mean_deltas"], norm["std_deltas"])
    dm = dynamics.NNDynamicsRewardModel(Env(), n10)
    dm.set_weights(weights, rweights)
    od = O.OracleRewardDynamics(weights, rweights, norm)
    inj = f32_uniform(rng, (H, K, 6))
    c = controllers.MPCcontrollerReward(Env(), dm, horizon=H, cost_fn=None, num_simulated_paths=K, gamma=0.95)
    c.sample_random_actions = lambda: inj
    want = O.OracleMPC(od, H, K, learned_reward=True, gamma=0.95).get_action(state, inj)
    assert np.allclose(c.get_action(state), want, rtol=1e-5, atol=1e-6)
    nxt, rew = dm.predict(state[None], inj[0, :1])
    wn, wr = od.predict(state[None], inj[0, :1])
    assert rew.shape == (1, 1) and np.abs(rew - wr).max() < 1e-4
    assert scaled_rel_err(nxt, wn, state_scale(norm)).max() < 1e-5

    pw, logstd, rms_mean, rms_std = O.synthetic_policy(rng, hidden=(128, 128), norm=norm)
    pol = policy.MlpPolicy(None, Env(), 128, 2)
    pol.set_weights(pw, logstd, rms_mean, rms_std)
    pc = controllers.MPCcontrollerPolicyNetReward(Env(), dm, pol, explore=0.3, self_exp=False, horizon=H, cost_fn=None,
                                                  num_simulated_paths=K, gamma=0.5)
    pc.sample_random_actions = lambda: inj
    op = O.OraclePolicy(pw, logstd, rms_mean, rms_std)
    want = O.OracleMPC(od, H, K, proposal="policy_mix", policy=op, explore=0.3, learned_reward=True, gamma=1.0).get_action(state, inj)
    assert np.allclose(pc.get_action(state), want, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("first_stage", ["uniform", "policy_gaussian", "policy_mean"])
def test_tree_search_controller(first_stage):
    """MCTScontrollerPolicyNetReward.get_action (controllers.py:397-457) as one N x (H+1) launch with the step-0
    proposal switched; scores and chosen first action vs the oracle restatement (which replays the R copies)."""
    from bc_mpc_b200 import controllers, dynamics, policy

    class Space:
        def __init__(self, n):
            self.low, self.high, self.shape = -np.ones(n), np.ones(n), (n,)

    class Env:
        action_space, observation_space = Space(6), Space(20)

    H, N, R, S = 5, 64, 4, 500
    rng, norm, weights, rweights, state = O.synthetic_reward_problem(37, S)
    n10 = (norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_reward"],
           norm["std_reward"], None, None, norm["mean_deltas"], norm["std_deltas"])
    dm = dynamics.NNDynamicsRewardModel(Env(), n10)
    dm.set_weights(weights, rweights)
    od = O.OracleRewardDynamics(weights, rweights, norm)
    pw, logstd, rms_mean, rms_std = O.synthetic_policy(rng, hidden=(128, 128), norm=norm)
    pol = policy.MlpPolicy(None, Env(), 128, 2)
    pol.set_weights(pw, logstd, rms_mean, rms_std)
    op = O.OraclePolicy(pw, logstd, rms_mean, rms_std)
    z = f32_uniform(rng, (N, 6)) if first_stage == "uniform" else rng.standard_normal((N, 6)).astype(np.float32).astype(np.float64)
    c = controllers.MCTScontrollerPolicyNetReward(Env(), dm, pol, self_exp=first_stage == "policy_gaussian", horizon=H,
                                                  num_first_stage_actions=N, random_path_per_action=R,
                                                  random_first_stage_action=first_stage == "uniform")
    c.sample_first_stage = lambda: z
    got = c.get_action(state)
    want, dbg = O.OracleMCTS(od, op, H, N, R, first_stage).get_action(state, z, return_debug=True)
    assert got.shape == (1, 6)
    inj = np.zeros((H + 1, N, 6))
    inj[0] = z
    _, ao, co = c._planner.rollout_debug(state, inj)
    assert np.allclose(ao[0], dbg["actions"], rtol=1e-5, atol=1e-6)
    assert np.abs(-co - dbg["total"]).max() < 1e-3 * max(1.0, np.abs(dbg["total"]).max())
    if first_stage != "policy_mean":       # with the deterministic first stage all N candidates tie
        assert c.last_record["index"] == dbg["best"]
    assert np.allclose(got, want, rtol=1e-5, atol=1e-6)
    # device-drawn first stage: replay the drawn tensor through the oracle
    c2 = controllers.MCTScontrollerPolicyNetReward(Env(), dm, pol, self_exp=first_stage == "policy_gaussian", horizon=H,
                                                   num_first_stage_actions=N, random_path_per_action=R,
                                                   random_first_stage_action=first_stage == "uniform")
    np.random.seed(5)
    got2 = c2.get_action(state)
    np.random.seed(5)
    seed = int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 31)
    zd = c2._planner.sample_injection(seed)[0, 0]
    want2 = O.OracleMCTS(od, op, H, N, R, first_stage).get_action(state, zd)
    assert np.allclose(got2, want2, rtol=1e-5, atol=1e-6)


def test_batched_controller_and_vectorized_sampler():
    """get_actions(states[B,D]) = B independent plans in one launch (grows the handle on demand); the vectorised
    sampler drives B model-envs with it.  Every plan is replayed through the oracle with the device-drawn tensor."""
    from bc_mpc_b200 import controllers, cost_functions, dynamics, sampling

    class Space:
        def __init__(self, n):
            self.low, self.high, self.shape = -np.ones(n), np.ones(n), (n,)

    class ModelEnv:            # the learned model as the environment: deterministic, CPU
        action_space, observation_space = Space(6), Space(20)

        def __init__(self, dyn, s0):
            self.dyn, self.s0 = dyn, s0

        def reset(self):
            self.s = self.s0.copy()
            return self.s

        def step(self, a):
            self.s = self.dyn.predict(self.s[None], np.asarray(a)[None])[0]
            return self.s, 0.0, False, {}

    K, H, B = 500, 5, 5
    rng, norm, weights, state = O.synthetic_problem(41, (64, 64))
    od = O.OracleDynamics(weights, norm)
    n10 = (norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], None, None, None, None,
           norm["mean_deltas"], norm["std_deltas"])
    dm = dynamics.NNDynamicsModel(ModelEnv, 2, 64, "relu", None, n10)
    dm.set_weights(weights)
    c = controllers.MPCcontroller(ModelEnv, dm, horizon=H, cost_fn=cost_functions.cheetah_cost_fn, num_simulated_paths=K)
    assert c.get_action(state).shape == (6,) and c._planner.max_states == 1
    S = norm["mean_obs"] + norm["std_obs"] * rng.standard_normal((B, 20))
    np.random.seed(11)
    acts = c.get_actions(S)
    assert acts.shape == (B, 6) and c._planner.max_states == 8
    np.random.seed(11)
    seed = int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 31)
    inj = c._planner.sample_injection(seed, n_states=B)
    for b in range(B):
        assert np.allclose(acts[b], O.OracleMPC(od, H, K).get_action(S[b], inj[b]), rtol=1e-6, atol=1e-7)
    envs = [ModelEnv(od, S[b]) for b in range(B)]
    np.random.seed(12)
    paths = sampling.sample_vectorized(envs, c, horizon=3)
    assert len(paths) == B and len(paths[0]["actions"]) == 3
    np.random.seed(12)
    for t in range(3):
        seed = int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 31)
        inj = c._planner.sample_injection(seed, n_states=B)
        for b in range(B):
            want = O.OracleMPC(od, H, K).get_action(paths[b]["observations"][t], inj[b])
            assert np.allclose(paths[b]["actions"][t], want, rtol=1e-6, atol=1e-7)


def test_threshold_and_tie_cases():
    K, H = 256, 4
    rng, norm, state, pl, mpc = make_pair(11, (64, 64), K, H)
    state[5], state[6], state[7] = 0.2, -0.0, -1e-12   # exactly on / next to the >= thresholds
    actions = f32_uniform(rng, (H, K, 6))
    actions[:, 200] = actions[:, 17]                      # identical sequences -> lowest index must win
    actions[:, 5] = actions[:, 17]
    act, dbg = mpc.get_action(state, actions, return_debug=True)
    so, ao, co = pl.rollout_debug(state, actions)
    assert co[5] == co[17] == co[200]
    assert np.array_equal(np.round(co[:8] - dbg["costs"][:8], 3), np.zeros(8))
    rec = pl.plan(state, inj=actions)[0]
    assert rec["index"] == dbg["argmin"]


@pytest.mark.parametrize("proposal,bc", [("policy_gaussian", False), ("policy_mix", False), ("bc_uniform", True)])
def test_policy_proposals_vs_oracle(proposal, bc):
    K, H = 8192 if proposal == "policy_gaussian" else 512, 15   # BASELINE.json configs[3] shape for P1
    rng, norm, state, pl, mpc = make_pair(21, (500, 500), K, H, proposal=proposal, bc=bc)
    inj = rng.standard_normal((H, K, 6)).astype(np.float32).astype(np.float64) if proposal == "policy_gaussian" \
        else f32_uniform(rng, (H, K, 6))
    act, dbg = mpc.get_action(state, inj, return_debug=True)
    so, ao, co = pl.rollout_debug(state, inj)
    assert scaled_rel_err(ao, dbg["actions"], 1.0).max() < TOL          # scale = half-width of the action box
    assert scaled_rel_err(so, dbg["states"], state_scale(norm)).max() < TOL
    rec = pl.plan(state, inj=inj)[0]
    assert rec["index"] == dbg["argmin"]
    assert np.allclose(rec["action"][:6], act, rtol=1e-5, atol=1e-6)


def test_predict_policy_cost_entry_points(golden_dir):
    rng, norm, state, pl, mpc = make_pair(5, (500, 500), 8, 2, proposal="policy_gaussian")
    n = 333
    S = norm["mean_obs"] + norm["std_obs"] * rng.standard_normal((n, 20))
    A = f32_uniform(rng, (n, 6))
    assert scaled_rel_err(pl.predict(S, A), mpc.dyn_model.predict(S, A), state_scale(norm)).max() < 1e-5
    assert rel_err(pl.predict(S, A), mpc.dyn_model.predict(S, A)).max() < TOL
    eps = rng.standard_normal((n, 6)).astype(np.float32)
    assert np.allclose(pl.policy_act(S), mpc.policy.act(S, False)[0], rtol=1e-5, atol=1e-6)
    assert np.allclose(pl.policy_act(S, eps), mpc.policy.act(S, True, noise=eps)[0], rtol=1e-5, atol=1e-6)
    g = np.load(os.path.join(golden_dir, "cost_kat.npz"))
    assert np.array_equal(pl.cheetah_cost(g["S"], g["N"]), g["batched"])     # bit-exact: float64 compare/sub/div


def test_device_sampling_replays_through_oracle():
    K, H = 1000, 15
    rng, norm, state, pl, mpc = make_pair(31, (500, 500), K, H)
    rec = pl.plan(state, seed=1234)[0]
    inj = pl.sample_injection(1234)[0]
    assert inj.shape == (H, K, 6) and (inj >= -1).all() and (inj < 1).all()
    act, dbg = mpc.get_action(state, inj, return_debug=True)
    assert rec["index"] == dbg["argmin"] and np.array_equal(rec["action"][:6], act)
    rec2 = pl.plan(state, seed=1235)[0]
    assert not np.array_equal(rec2["action"], rec["action"])
    # statistics of U[-1,1): mean 0, var 1/3
    assert abs(inj.mean()) < 0.01 and abs(inj.var() - 1 / 3) < 0.01
    cols = inj.reshape(-1, 6)
    assert np.abs(np.corrcoef(cols.T) - np.eye(6)).max() < 0.04


def test_normal_sampling_statistics():
    rng, norm, state, pl, mpc = make_pair(32, (64, 64), 4096, 8, proposal="policy_gaussian")
    z = pl.sample_injection(7)[0].reshape(-1)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    assert abs((z ** 3).mean()) < 0.03 and abs((z ** 4).mean() - 3) < 0.1


def test_batched_states_equal_single_plans():
    K, H, B = 300, 6, 5
    rng, norm, state, pl, mpc = make_pair(41, (64, 64), K, H, max_states=B)
    states = norm["mean_obs"] + norm["std_obs"] * rng.standard_normal((B, 20))
    inj = f32_uniform(rng, (B, H, K, 6))
    recs = pl.plan(states, inj=inj)
    for b in range(B):
        act, dbg = mpc.get_action(states[b], inj[b], return_debug=True)
        assert recs[b]["index"] == dbg["argmin"] and np.array_equal(recs[b]["action"][:6], act)
    # device sampling: root-state index is part of the Philox counter, so states do not share draws
    recs = pl.plan(np.tile(states[0], [B, 1]), seed=3)
    assert len({int(r["index"]) for r in recs}) > 1


def test_sharded_paths_match_single_gpu_result():
    import torch
    from bc_mpc_b200 import _lib as L
    K, H, G = 1000, 8, 4
    rng, norm, state, pl, mpc = make_pair(51, (64, 64), K, H)
    want = pl.plan(state, seed=99)[0]
    bounds = [0, 250, 500, 777, 1000]     # ragged shards
    gathered = torch.zeros(G * 80, dtype=torch.uint8, device="cuda")
    d_state = torch.tensor(state, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()    # the handles launch on their own streams
    shards = []
    for g in range(G):
        _, _, _, sp, _ = make_pair(51, (64, 64), bounds[g + 1] - bounds[g], H, path_offset=bounds[g], total_paths=K)
        sp.plan_device(d_state.data_ptr(), 1, gathered.data_ptr() + g * 80, seed=99)
        shards.append(sp)
    torch.cuda.synchronize()
    out = torch.zeros(80, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    pl.reduce_records_device(gathered.data_ptr(), G, 1, out.data_ptr())
    torch.cuda.synchronize()
    got = np.frombuffer(out.cpu().numpy().tobytes(), L.RECORD_DTYPE)[0]
    assert got["index"] == want["index"] and got["cost"] == want["cost"]
    assert np.array_equal(got["action"], want["action"])


def test_edge_shapes_and_errors():
    from bc_mpc_b200 import Planner
    for K, H in [(1, 1), (33, 2), (129, 3)]:
        rng, norm, state, pl, mpc = make_pair(61, (40, 40), K, H)
        actions = f32_uniform(rng, (H, K, 6))
        act, dbg = mpc.get_action(state, actions, return_debug=True)
        rec = pl.plan(state, inj=actions)[0]
        assert rec["index"] == dbg["argmin"] and np.array_equal(rec["action"][:6], act)
    pl = Planner(num_paths=4, horizon=2, hidden_size=8)
    with pytest.raises(RuntimeError, match="dynamics weights not set"):
        pl.plan(np.zeros(20))
    with pytest.raises(ValueError):
        pl.set_dynamics([(np.zeros((26, 9), np.float32), np.zeros(9, np.float32))] * 3)
    with pytest.raises(RuntimeError, match="obs_dim"):
        Planner(obs_dim=64)
    with pytest.raises(RuntimeError, match="n_states"):
        rng, norm, state, pl, mpc = make_pair(61, (40, 40), 4, 2)
        pl.plan(np.zeros((2, 20)))


def test_error_paths_of_the_widened_api():
    """Every unsupported or mis-ordered request fails loudly with a message; nothing falls back silently."""
    import ctypes as C
    from bc_mpc_b200 import Planner, _lib as L
    rng, norm, weights, rweights, state = O.synthetic_reward_problem(3, 40)
    pl = Planner(20, 6, 2, 40, "tanh", num_paths=8, horizon=2)
    with pytest.raises(RuntimeError, match="dynamics weights not set"):
        pl.set_layer_norm([(np.ones(40), np.zeros(40))] * 2)
    with pytest.raises(RuntimeError, match="reward head not set"):
        pl.predict_reward(state[None], np.zeros((1, 6)))
    pl.set_dynamics(weights)
    with pytest.raises(ValueError, match="reward head must be"):
        pl.set_reward_head(rweights[:1])
    with pytest.raises(ValueError, match="LayerNorm gamma/beta"):
        pl.set_layer_norm([(np.ones(41), np.zeros(41))] * 2)
    with pytest.raises(RuntimeError, match="n_hidden"):
        pl.set_layer_norm([(np.ones(40), np.zeros(40))])
    rc = pl._lib.bcmpc_set_reward_layer_norm(pl._h, None, None, C.c_float(1e-12))
    assert rc < 0 and b"reward head not set" in pl._lib.bcmpc_last_error(pl._h)
    with pytest.raises(RuntimeError, match="unknown layout|layout"):
        L.check(pl._lib.bcmpc_set_dynamics_device(pl._h, (C.c_void_p * 3)(1, 1, 1), (C.c_void_p * 3)(1, 1, 1), 3, 7, None), pl._h)
    Planner(num_paths=4, horizon=2, hidden_size=8, first_proposal="uniform", precision="bf16").close()   # (tree search runs in fast mode since round 2)
    three = Planner(20, 6, 3, 40, "relu", num_paths=8, horizon=2)
    three.set_dynamics(O.synthetic_problem(3, (40, 40, 40))[2])
    with pytest.raises(RuntimeError, match="n_hidden must be 2"):
        three.set_reward_head(rweights)
    import os
    fast = Planner(20, 6, 2, 40, "tanh", num_paths=8, horizon=2, precision="bf16")
    fast.set_dynamics(weights)
    if os.environ.get("BCMPC_TC_CG") == "1":                                     # single-CTA developer mode: reward model widths 193..256 / 449..512 only
        with pytest.raises(RuntimeError, match="no tensor-core variant"):
            fast.set_reward_head(rweights)
    else:
        fast.set_reward_head(rweights)                                          # CTA pairs: every width
    with pytest.raises(RuntimeError, match="tanh layers"):                       # ... and tanh only
        fast = Planner(20, 6, 2, 500, "relu", num_paths=8, horizon=2, precision="bf16")
        fast.set_dynamics(O.synthetic_problem(3, (500, 500))[2])
        fast.set_reward_head(O.synthetic_reward_problem(3, 500)[3])


def test_nan_cost_wins_like_numpy():
    K, H = 64, 2
    rng, norm, state, pl, mpc = make_pair(71, (32, 32), K, H)
    actions = f32_uniform(rng, (H, K, 6))
    actions[1, 40, 2] = np.nan          # a NaN action poisons path 40's last step
    with np.errstate(invalid="ignore"):
        act, dbg = mpc.get_action(state, actions, return_debug=True)
    assert dbg["argmin"] == 40
    rec = pl.plan(state, inj=actions)[0]
    assert rec["index"] == 40 and np.isnan(rec["cost"])
