"""Oracle restatement vs the reference's own numpy modules, executed verbatim.
Runs only where /root/reference is mounted (this container, not the GPU box)."""
import os
import sys

import numpy as np
import pytest

from oracle import mpc_oracle as O

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "controllers.py")),
                                reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, REF)
    try:
        import controllers
        import cost_functions
    finally:
        sys.path.remove(REF)
    return controllers, cost_functions


class _Space:
    def __init__(self, n):
        self.low, self.high, self.shape = -np.ones(n), np.ones(n), (n,)


class _Env:
    action_space, observation_space = _Space(6), _Space(20)


def test_cost_matches_reference_random(ref):
    _, cf = ref
    rng = np.random.default_rng(0)
    for _ in range(20):
        S = rng.standard_normal((257, 20)) * 0.3
        S[rng.integers(0, 257, 10), rng.integers(5, 8, 10)] = rng.choice([0.0, -0.0, 0.2], 10)
        N = S + rng.standard_normal((257, 20)) * 0.1
        assert np.array_equal(cf.cheetah_cost_fn(S, None, N), O.cheetah_cost_fn(S, None, N))
        assert cf.cheetah_cost_fn(S[3], None, N[3]) == O.cheetah_cost_fn(S[3], None, N[3])


@pytest.mark.parametrize("seed,K,H", [(0, 1000, 15), (1, 129, 7), (2, 1, 1)])
def test_get_action_matches_reference(ref, seed, K, H):
    ctl, cf = ref
    rng, norm, weights, state = O.synthetic_problem(seed, (48, 48))
    dyn = O.OracleDynamics(weights, norm)
    actions = rng.uniform(-1, 1, (H, K, 6))
    c = ctl.MPCcontroller(_Env(), dyn, horizon=H, cost_fn=cf.cheetah_cost_fn, num_simulated_paths=K)
    c.sample_random_actions = lambda: actions
    want = c.get_action(state)
    got = O.OracleMPC(dyn, H, K).get_action(state, actions)
    assert np.array_equal(want, got)


def test_policy_controller_matches_reference(ref):
    ctl, cf = ref
    rng, norm, weights, state = O.synthetic_problem(4, (48, 48))
    dyn = O.OracleDynamics(weights, norm)
    pol = O.OraclePolicy(*O.synthetic_policy(rng, hidden=(16, 16), norm=norm))
    H, K = 5, 77
    uni = rng.uniform(-1, 1, (H, K, 6))

    class P:
        def act(self, ob, stochastic=True):
            return pol.act(ob, False)

    c = ctl.MPCcontrollerPolicyNet(_Env(), dyn, P(), explore=0.3, self_exp=False, horizon=H,
                                   cost_fn=cf.cheetah_cost_fn, num_simulated_paths=K)
    c.sample_random_actions = lambda: uni
    want = c.get_action(state)
    got = O.OracleMPC(dyn, H, K, proposal="policy_mix", policy=pol, explore=0.3).get_action(state, uni)
    assert np.array_equal(want, got)


@pytest.mark.parametrize("first_stage", ["uniform", "policy_gaussian", "policy_mean"])
def test_tree_search_controller_matches_reference(ref, first_stage):
    """MCTScontrollerPolicyNetReward.get_action (controllers.py:397-457): N first-stage actions, deterministic-policy
    rollouts, reward_1 + mean rollout reward, argmax."""
    ctl, _ = ref
    rng, norm, weights, rweights, state = O.synthetic_reward_problem(8, 40)
    dyn = O.OracleRewardDynamics(weights, rweights, norm)
    pol = O.OraclePolicy(*O.synthetic_policy(rng, hidden=(16, 16), norm=norm))
    H, N, R = 4, 23, 3
    inj = rng.uniform(-1, 1, (N, 6)) if first_stage == "uniform" else rng.standard_normal((N, 6))

    class _Pol:
        calls = 0

        def act(self, ob, stochastic=True):
            if stochastic:          # the N first-stage draws come one row at a time (:414)
                out = pol.act(ob, True, noise=inj[self.calls:self.calls + 1])
                self.calls += 1
                return out
            return pol.act(ob, False)

    class _Space(_Env.action_space.__class__):
        n = 0

        def sample(self):
            _Space.n += 1
            return inj[_Space.n - 1]

    env = _Env()
    env.action_space = _Space(6)
    c = ctl.MCTScontrollerPolicyNetReward(env, dyn, _Pol(), self_exp=first_stage == "policy_gaussian", horizon=H,
                                          num_first_stage_actions=N, random_path_per_action=R,
                                          random_first_stage_action=first_stage == "uniform")
    want = c.get_action(state)
    got = O.OracleMCTS(dyn, pol, H, N, R, first_stage).get_action(state, inj)
    assert want.shape == (1, 6) and np.array_equal(want, got)


def test_sample_random_actions_layout(ref):
    ctl, cf = ref
    c = ctl.MPCcontroller(_Env(), None, horizon=3, cost_fn=None, num_simulated_paths=5)
    a = c.sample_random_actions()
    assert a.shape == (3, 5, 6) and a.dtype == np.float64 and (a >= -1).all() and (a < 1).all()


@pytest.mark.parametrize("gamma", [1.0, 0.9])
def test_learned_reward_controller_matches_reference(ref, gamma):
    """MPCcontrollerReward (controllers.py:90-158): sum of gamma**i * predicted reward, argmax, first action."""
    ctl, _ = ref
    rng, norm, weights, rweights, state = O.synthetic_reward_problem(6, 40)
    dyn = O.OracleRewardDynamics(weights, rweights, norm)
    H, K = 6, 300
    actions = rng.uniform(-1, 1, (H, K, 6))
    c = ctl.MPCcontrollerReward(_Env(), dyn, horizon=H, num_simulated_paths=K, gamma=gamma)
    c.sample_random_actions = lambda: actions
    want = c.get_action(state)
    got, dbg = O.OracleMPC(dyn, H, K, learned_reward=True, gamma=gamma).get_action(state, actions, return_debug=True)
    assert np.array_equal(want, got)
    assert dbg["costs"].shape == (K,)


def test_learned_reward_policy_controller_matches_reference(ref, capsys):
    """MPCcontrollerPolicyNetReward (controllers.py:289-363): policy proposals, plain reward sum, argmax."""
    ctl, _ = ref
    rng, norm, weights, rweights, state = O.synthetic_reward_problem(7, 40)
    dyn = O.OracleRewardDynamics(weights, rweights, norm)
    pol = O.OraclePolicy(*O.synthetic_policy(rng, hidden=(16, 16), norm=norm))
    H, K = 5, 200
    uni = rng.uniform(-1, 1, (H, K, 6))

    class _Det:      # deterministic branch: act(stochastic=False) is the mean
        def act(self, ob, stochastic=True):
            return pol.act(ob, stochastic=False)

    c = ctl.MPCcontrollerPolicyNetReward(_Env(), dyn, _Det(), explore=0.3, self_exp=False, horizon=H, num_simulated_paths=K, gamma=0.5)
    c.sample_random_actions = lambda: uni
    want = c.get_action(state)
    got = O.OracleMPC(dyn, H, K, proposal="policy_mix", policy=pol, explore=0.3, learned_reward=True, gamma=1.0).get_action(state, uni)
    assert np.allclose(want, got, rtol=0, atol=0)
