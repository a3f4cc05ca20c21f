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


def _norm_err(got, ref, norm):
    sd = np.where(norm["std_deltas"] > 0, norm["std_deltas"], 1.0)
    return (got - ref) / sd


@pytest.mark.parametrize("S,proposal", [(500, "uniform"), (256, "uniform"), (100, "uniform"), (320, "policy_mix"), (500, "policy_gaussian"), (64, "uniform")])
def test_plain_tanh_dynamics_fast_mode(S, proposal):
    """NNDynamicsModel with tanh hidden layers - the reference builder's default activation (dynamics.py:60) - on the
    tcgen05 kernel: explicit biases in the epilogues, tanh.approx.bf16x2 on the rounded pair."""
    from bc_mpc_b200 import Planner
    K, H = 1000, 8
    rng, norm, weights, state = O.synthetic_problem(41, (S, S))
    dyn = O.OracleDynamics(weights, norm, activation="tanh")
    pl = Planner(20, 6, 2, S, "tanh", num_paths=K, horizon=H, precision="bf16", proposal=proposal, explore=0.4)
    pl.set_dynamics(weights)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
    pl.set_action_bounds(-np.ones(6), np.ones(6))
    pol = None
    if proposal != "uniform":
        pw, logstd, rms_mean, rms_std = O.synthetic_policy(rng, hidden=(128, 128), norm=norm)
        pol = O.OraclePolicy(pw, logstd, rms_mean, rms_std)
        pl.set_policy(pw, logstd, rms_mean, rms_std, clip=5.0)
    mpc = O.OracleMPC(dyn, H, K, proposal=proposal, policy=pol, explore=0.4)
    inj = rng.standard_normal((H, K, 6)).astype(np.float32).astype(np.float64) if proposal == "policy_gaussian" else f32_uniform(rng, (H, K, 6))
    act, dbg = mpc.get_action(state, inj, return_debug=True)
    so, ao, co = pl.rollout_debug(state, inj)
    e = _norm_err(so, dbg["states"], norm)
    assert np.abs(_norm_err(so[1], dbg["states"][1], norm)).max() < 0.1 and np.abs(e).max() < 0.6, (np.abs(e[1]).max(), np.abs(e).max())
    assert np.corrcoef(co, dbg["costs"])[0, 1] > 0.98
    rec = pl.plan(state, inj=inj)[0]
    assert rec["index"] == int(np.argmin(co))
    assert (dbg["costs"][rec["index"]] - dbg["costs"].min()) / dbg["costs"].std() < 0.5
    prod = pl.plan(state, seed=11)[0]
    again = pl.plan(state, inj=pl.sample_injection(11)[0])[0]
    assert prod["index"] == again["index"] and prod["cost"] == again["cost"]
    # the same net with LayerNorm after every tanh (the builder default activation + FLAGS.LAYER_NORM): variants 26 / 27
    lns = [(1.0 + 0.1 * rng.standard_normal(S).astype(np.float32), 0.1 * rng.standard_normal(S).astype(np.float32)) for _ in range(2)]
    pl.set_layer_norm(lns)
    mpc_ln = O.OracleMPC(O.OracleDynamics(weights, norm, activation="tanh", layer_norm=lns), H, K, proposal=proposal, policy=pol, explore=0.4)
    act, dbg = mpc_ln.get_action(state, inj, return_debug=True)
    so_ln, ao, co = pl.rollout_debug(state, inj)
    e = _norm_err(so_ln, dbg["states"], norm)
    assert np.abs(e[1]).max() < 0.15 and np.abs(e).max() < 1.0, (np.abs(e[1]).max(), np.abs(e).max())
    assert np.corrcoef(co, dbg["costs"])[0, 1] > 0.98 and np.abs(so_ln - so).max() > 1e-3
    rec = pl.plan(state, inj=inj)[0]
    assert rec["index"] == int(np.argmin(co))
    assert (dbg["costs"][rec["index"]] - dbg["costs"].min()) / dbg["costs"].std() < 0.5
    prod = pl.plan(state, seed=11)[0]
    again = pl.plan(state, inj=pl.sample_injection(11)[0])[0]
    assert prod["index"] == again["index"] and prod["cost"] == again["cost"]


@pytest.mark.parametrize("S,proposal,ln", [(100, "uniform", False), (320, "policy_mix", False), (400, "uniform", True), (448, "policy_gaussian", True),
                                           (130, "policy_mix", True), (384, "uniform", False)])
def test_learned_reward_model_at_every_width(S, proposal, ln):
    """The learned-reward planners (and their LayerNorm / proposal-network combinations) at widths other than the
    reference's 500 / 256: every number of 64-unit chunks has a tensor-core instantiation in CTA-pair mode."""
    from bc_mpc_b200 import Planner
    K, H, gamma = 1000, 6, 0.95
    rng, norm, weights, rweights, state = O.synthetic_reward_problem(33, S)
    dyn = O.OracleRewardDynamics(weights, rweights, norm)
    pl = Planner(20, 6, 2, S, "tanh", num_paths=K, horizon=H, precision="bf16", proposal=proposal, explore=0.4)
    pl.set_dynamics(weights)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
    lns = None
    if ln:
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
    e = _norm_err(so, dbg["states"], norm)
    assert np.abs(_norm_err(so[1], dbg["states"][1], norm)).max() < (0.1 if ln else 0.06) and np.abs(e).max() < (0.6 if ln else 0.3)
    assert np.corrcoef(co, dbg["costs"])[0, 1] > 0.98
    rec = pl.plan(state, inj=inj)[0]
    assert rec["index"] == int(np.argmin(co))
    assert (dbg["costs"][rec["index"]] - dbg["costs"].min()) / dbg["costs"].std() < 0.5


@pytest.mark.parametrize("S", [330, 400, 448])
def test_layer_norm_with_policy_at_every_width(S):
    """FLAGS.LAYER_NORM + MPCcontrollerPolicyNet (the reference's default run) at 5, 6 and 7 hidden chunks."""
    from bc_mpc_b200 import Planner
    K, H = 1000, 6
    rng, norm, weights, state = O.synthetic_problem(23, (S, S))
    lns = [(1.0 + 0.1 * rng.standard_normal(S).astype(np.float32), 0.1 * rng.standard_normal(S).astype(np.float32)) for _ in range(2)]
    dyn = O.OracleDynamics(weights, norm, layer_norm=lns)
    pw, logstd, rms_mean, rms_std = O.synthetic_policy(rng, hidden=(128, 128), norm=norm)
    pol = O.OraclePolicy(pw, logstd, rms_mean, rms_std)
    pl = Planner(20, 6, 2, S, "relu", num_paths=K, horizon=H, precision="bf16", proposal="policy_mix", explore=0.5)
    pl.set_dynamics(weights)
    pl.set_layer_norm(lns)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
    pl.set_action_bounds(-np.ones(6), np.ones(6))
    pl.set_policy(pw, logstd, rms_mean, rms_std, clip=5.0)
    mpc = O.OracleMPC(dyn, H, K, proposal="policy_mix", policy=pol, explore=0.5)
    inj = f32_uniform(rng, (H, K, 6))
    act, dbg = mpc.get_action(state, inj, return_debug=True)
    so, ao, co = pl.rollout_debug(state, inj)
    e = _norm_err(so, dbg["states"], norm)
    assert np.abs(e).max() < 0.6 and np.corrcoef(co, dbg["costs"])[0, 1] > 0.98
    rec = pl.plan(state, inj=inj)[0]
    assert rec["index"] == int(np.argmin(co))
    assert (dbg["costs"][rec["index"]] - dbg["costs"].min()) / dbg["costs"].std() < 0.5
