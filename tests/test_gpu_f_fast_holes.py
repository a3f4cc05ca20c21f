"""bf16 fast-mode coverage added in round 2 (VERDICT r1 "missing" 2-5): planners and shapes that used to run in
verification mode only."""
import numpy as np
import pytest

from oracle import mpc_oracle as O
from tests._util import f32_uniform

pytestmark = pytest.mark.gpu


class _Space:
    def __init__(self, n):
        self.low, self.high, self.shape = -np.ones(n), np.ones(n), (n,)


class _Env:
    action_space, observation_space = _Space(6), _Space(20)


@pytest.mark.parametrize("first_stage", ["uniform", "policy_gaussian", "policy_mean"])
@pytest.mark.parametrize("S", [500, 256])
def test_tree_search_controller_fast_mode(first_stage, S):
    """MCTScontrollerPolicyNetReward.get_action (controllers.py:397-457) on the tcgen05 kernel: the step-0 proposal mode
    (uniform / stochastic policy / deterministic policy) differs from the continuation's (deterministic policy)."""
    from bc_mpc_b200 import controllers, dynamics, policy
    H, N, R = 5, 640, 4
    rng, norm, weights, rweights, state = O.synthetic_reward_problem(37, S)
    n10 = (norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_reward"],
           norm["std_reward"], None, None, norm["mean_deltas"], norm["std_deltas"])
    dm = dynamics.NNDynamicsRewardModel(_Env(), n10, size=S)
    dm.set_weights(weights, rweights)
    od = O.OracleRewardDynamics(weights, rweights, norm)
    pw, logstd, rms_mean, rms_std = O.synthetic_policy(rng, hidden=(128, 128), norm=norm)
    pol = policy.MlpPolicy(None, _Env(), 128, 2)
    pol.set_weights(pw, logstd, rms_mean, rms_std)
    op = O.OraclePolicy(pw, logstd, rms_mean, rms_std)
    z = f32_uniform(rng, (N, 6)) if first_stage == "uniform" else rng.standard_normal((N, 6)).astype(np.float32).astype(np.float64)
    kw = dict(self_exp=first_stage == "policy_gaussian", horizon=H, num_first_stage_actions=N, random_path_per_action=R,
              random_first_stage_action=first_stage == "uniform")
    c = controllers.MCTScontrollerPolicyNetReward(_Env(), dm, pol, precision="bf16", **kw)
    c.sample_first_stage = lambda: z
    got = c.get_action(state)
    want, dbg = O.OracleMCTS(od, op, H, N, R, first_stage).get_action(state, z, return_debug=True)
    assert got.shape == (1, 6)
    inj = np.zeros((H + 1, N, 6))
    inj[0] = z
    _, ao, co = c._planner.rollout_debug(state, inj)
    # first-stage actions: the injected draw (uniform), or the proposal network's output in bf16 (+ noise)
    assert np.abs(ao[0] - dbg["actions"]).max() < (1e-6 if first_stage == "uniform" else 0.03)
    score = -co
    spread = dbg["total"].std() if first_stage != "policy_mean" else 1.0
    assert np.abs(score - dbg["total"]).max() < 0.35 * max(spread, 1e-3) + 0.05 * np.abs(dbg["total"]).max()
    if first_stage != "policy_mean":        # with the deterministic first stage all N candidates tie
        assert np.corrcoef(score, dbg["total"])[0, 1] > 0.97
        regret = (dbg["total"].max() - dbg["total"][c.last_record["index"]]) / spread
        assert regret < 0.5, regret
    else:
        assert c.last_record["index"] == 0 and np.abs(got - want).max() < 0.03
    # production flavour (device-drawn first stage) == replay of the drawn tensor through the debug flavour
    np.random.seed(5)
    got2 = c.__class__(_Env(), dm, pol, precision="bf16", **kw).get_action(state)
    np.random.seed(5)
    seed = int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 31)
    zd = c._planner.sample_injection(seed)[0, 0]
    c.sample_first_stage = lambda: zd
    assert np.array_equal(c.get_action(state), got2)
