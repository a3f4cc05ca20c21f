"""Oracle pinned against the golden vectors produced by the reference's own
controllers.py / cost_functions.py (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import mpc_oracle as O


def test_cost_kats(golden_dir):
    g = np.load(os.path.join(golden_dir, "cost_kat.npz"))
    got = O.cheetah_cost_fn(g["S"], g["A"], g["N"])
    assert np.array_equal(got, g["batched"])
    sc = np.array([O.cheetah_cost_fn(g["S"][i], g["A"][i], g["N"][i]) for i in range(len(g["S"]))])
    assert np.array_equal(sc, g["scalar"])
    assert np.array_equal(O.trajectory_cost_fn(O.cheetah_cost_fn, g["St"], g["At"], g["Nt"]), g["traj"])


def test_cost_known_answers():
    # SURVEY.md section 4, produced by running the reference function.
    z = np.zeros((1, 20))
    assert O.cheetah_cost_fn(z, None, z)[0] == 20.0
    s = z.copy(); s[0, 5] = 0.2
    assert O.cheetah_cost_fn(s, None, s)[0] == 30.0
    s[0, 5] = 0.19999
    assert O.cheetah_cost_fn(s, None, s)[0] == 20.0
    s = z.copy(); s[0, 6] = -0.0
    assert O.cheetah_cost_fn(s, None, s)[0] == 20.0
    s = z.copy(); s[0, 6] = -1; s[0, 7] = -1
    n = s.copy(); n[0, 17] = 0.05
    assert O.cheetah_cost_fn(s, None, n)[0] == -(0.05 / 0.01)
    s = z.copy(); s[0, 5] = 1
    n = s.copy(); n[0, 17] = -0.02
    assert O.cheetah_cost_fn(s, None, n)[0] == 32.0
    assert np.argmin([3, 1, 1, 2]) == 1 and np.argmin([3, np.nan, 1]) == 1


def test_uniform_plans(golden_dir):
    g = np.load(os.path.join(golden_dir, "mpc_uniform.npz"))
    for tag in "abc":
        hidden = tuple(int(x) for x in g[f"{tag}_hidden"])
        rng, norm, weights, state = O.synthetic_problem(int(g[f"{tag}_seed"]), hidden)
        assert np.array_equal(state, g[f"{tag}_state"])
        actions = g[f"{tag}_actions"]
        H, K = actions.shape[:2]
        mpc = O.OracleMPC(O.OracleDynamics(weights, norm), H, K)
        act, dbg = mpc.get_action(state, actions, return_debug=True)
        assert np.array_equal(act, g[f"{tag}_first_action"])
        assert np.array_equal(dbg["costs"], g[f"{tag}_costs"])
        assert np.array_equal(dbg["states"], g[f"{tag}_states"])
        assert dbg["argmin"] == int(g[f"{tag}_argmin"])
    assert int(g["a_argmin"]) != 7  # the duplicated path never wins over its lower-index twin


def test_policy_plans(golden_dir):
    g = np.load(os.path.join(golden_dir, "mpc_policy.npz"))
    hidden = tuple(int(x) for x in g["hidden"])
    rng, norm, weights, state = O.synthetic_problem(int(g["seed"]), hidden)
    dyn = O.OracleDynamics(weights, norm)
    pol = O.OraclePolicy(*O.synthetic_policy(rng, hidden=(32, 32), norm=norm))
    H, K = g["eps"].shape[:2]
    a = O.OracleMPC(dyn, H, K, proposal="policy_gaussian", policy=pol).get_action(state, g["eps"])
    assert np.array_equal(a, g["selfexp_first_action"])
    a = O.OracleMPC(dyn, H, K, proposal="policy_mix", policy=pol, explore=float(g["explore"])).get_action(
        state, g["uni"])
    assert np.array_equal(a, g["mix_first_action"])


def _reward_fixture(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "mpc_reward.npz"))
    S = int(g["hidden"])
    rng, norm, weights, rweights, state = O.synthetic_reward_problem(int(g["seed"]), S)
    lns = [(g["ln_gamma"][i], g["ln_beta"][i]) for i in range(3)]
    for _ in range(3):      # the generator drew the LayerNorm parameters before the policy: keep the stream aligned
        rng.standard_normal(S), rng.standard_normal(S)
    pol = O.OraclePolicy(*O.synthetic_policy(rng, hidden=(32, 32), norm=norm))
    dyn = (O.OracleRewardDynamics(weights, rweights, norm) if tag == "plain" else
           O.OracleRewardDynamics(weights, rweights, norm, layer_norm=lns[:2], reward_layer_norm=lns[2]))
    assert np.array_equal(state, g["state"])
    return g, dyn, pol, state


@pytest.mark.parametrize("tag", ["plain", "ln"])
def test_reward_and_tree_search_plans(golden_dir, tag):
    """mpc_reward.npz: the reference's MPCcontrollerReward / MPCcontrollerPolicyNetReward / MCTScontrollerPolicyNetReward
    (controllers.py:90-158, 289-363, 365-457) run on the oracle reward model."""
    g, dyn, pol, state = _reward_fixture(golden_dir, tag)
    H, K = g["actions"].shape[:2]
    N, R = int(g["N"]), int(g["R"])
    for gamma in (1.0, 0.9):
        a, dbg = O.OracleMPC(dyn, H, K, learned_reward=True, gamma=gamma).get_action(state, g["actions"], return_debug=True)
        assert np.array_equal(a, g[f"{tag}_uniform_g{gamma}_first_action"])
        if gamma == 0.9:
            assert np.allclose(-dbg["costs"], g[f"{tag}_uniform_g0.9_rewards"], rtol=1e-12, atol=1e-12)
    a = O.OracleMPC(dyn, H, K, proposal="policy_mix", policy=pol, explore=0.4, learned_reward=True).get_action(state, g["actions"])
    assert np.array_equal(a, g[f"{tag}_mix_first_action"])
    a = O.OracleMPC(dyn, H, K, proposal="policy_gaussian", policy=pol, learned_reward=True).get_action(state, g["eps"])
    assert np.array_equal(a, g[f"{tag}_selfexp_first_action"])
    a = O.OracleMCTS(dyn, pol, H, N, R, "uniform").get_action(state, g["first_uni"])
    assert np.array_equal(a, g[f"{tag}_tree_uniform_action"])
    a = O.OracleMCTS(dyn, pol, H, N, R, "policy_gaussian").get_action(state, g["first_eps"])
    assert np.array_equal(a, g[f"{tag}_tree_selfexp_action"])


def test_trained_weights_plan(golden_dir):
    g = np.load(os.path.join(golden_dir, "vars_ppo_mpc.npz"))
    dyn = O.OracleDynamics([(g[f"dyn_w{i}"], g[f"dyn_b{i}"]) for i in range(3)], O.synthetic_normalization())
    H, K = g["plan_actions"].shape[:2]
    act, dbg = O.OracleMPC(dyn, H, K).get_action(g["plan_state"], g["plan_actions"], return_debug=True)
    assert np.array_equal(act, g["plan_first_action"])
    assert np.array_equal(dbg["costs"], g["plan_costs"])
    mean, std = O.OraclePolicy.rms_from_sums(g["ob_sum"], g["ob_sumsq"], float(g["ob_count"]))
    assert std.min() >= 0.1 - 1e-7 and abs(mean[17] - 10.3) < 0.05  # min sumsq/count = 1e-2


def test_zero_std_dim_is_stable():
    rng, norm, weights, state = O.synthetic_problem(5, (32, 32))
    assert norm["std_obs"][18] == 0 and norm["std_deltas"][18] == 0
    state[18] = 0.0
    dyn = O.OracleDynamics(weights, norm)
    s = np.tile(state, [4, 1])
    for _ in range(3):
        s = dyn.predict(s, rng.uniform(-1, 1, (4, 6)))
    assert np.all(s[:, 18] == 0.0) and np.all(np.isfinite(s))
