"""Proposal networks of up to 256 units per hidden layer on the tcgen05 path (kVarPolWide kernel variants; VERDICT r1
"missing" 5): MlpPolicy(hid_size=256) is what various_train/train_mpc_ppo_r.py:142 builds, next to the learned-reward
model.  Checked against the float32 / float64 oracle on injected noise, like the 128-unit networks in test_gpu_b_fast.py."""
import numpy as np
import pytest

from oracle import mpc_oracle as O
from tests._util import f32_uniform

pytestmark = pytest.mark.gpu


def _norm_err(got, ref, norm):
    sd = np.where(norm["std_deltas"] > 1e-6, norm["std_deltas"], 1.0)
    return (got - ref) / sd


def _build(seed, S, activation, proposal, hid, ln, reward, K, H, explore=0.4, gamma=0.95, policy_first=False):
    """planner + oracle for one (dynamics flavour, proposal network) combination"""
    from bc_mpc_b200 import Planner
    if reward:
        rng, norm, weights, rweights, state = O.synthetic_reward_problem(seed, S)
    else:
        rng, norm, weights, state = O.synthetic_problem(seed, (S, S))
        rweights = None
    n_ln = 3 if reward else 2
    lns = [(1.0 + 0.1 * rng.standard_normal(S).astype(np.float32), 0.1 * rng.standard_normal(S).astype(np.float32)) for _ in range(n_ln)] if ln else None
    if reward:
        dyn = O.OracleRewardDynamics(weights, rweights, norm, layer_norm=lns[:2], reward_layer_norm=lns[2]) if ln else O.OracleRewardDynamics(weights, rweights, norm)
    else:
        dyn = O.OracleDynamics(weights, norm, activation=activation, layer_norm=lns) if ln else O.OracleDynamics(weights, norm, activation=activation)
    pw, logstd, rms_mean, rms_std = O.synthetic_policy(rng, hidden=hid, norm=norm)
    pol = O.OraclePolicy(pw, logstd, rms_mean, rms_std)
    pl = Planner(20, 6, 2, S, activation, num_paths=K, horizon=H, precision="bf16", proposal=proposal, explore=explore)
    if policy_first:
        pl.set_policy(pw, logstd, rms_mean, rms_std, clip=5.0)
    pl.set_dynamics(weights)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
    pl.set_action_bounds(-np.ones(6), np.ones(6))
    if ln:
        pl.set_layer_norm(lns[:2])
    if reward:
        pl.set_reward_head(rweights, norm["mean_reward"], norm["std_reward"], gamma, layer_norm=lns[2] if ln else None)
    if not policy_first:
        pl.set_policy(pw, logstd, rms_mean, rms_std, clip=5.0)
    mpc = O.OracleMPC(dyn, H, K, proposal=proposal, policy=pol, explore=explore, learned_reward=reward, gamma=gamma if reward else 1.0)
    return rng, norm, state, pl, mpc, (pw, logstd, rms_mean, rms_std)


def _check(rng, norm, state, pl, mpc, proposal, K, H, tol_first, tol_all):
    inj = rng.standard_normal((H, K, 6)).astype(np.float32).astype(np.float64) if proposal == "policy_gaussian" else f32_uniform(rng, (H, K, 6))
    act, dbg = mpc.get_action(state, inj, return_debug=True)
    so, ao, co = pl.rollout_debug(state, inj)
    assert np.abs(ao[0] - dbg["actions"][0]).max() < 0.03, np.abs(ao[0] - dbg["actions"][0]).max()     # first proposals: same state, bf16 policy net
    e = _norm_err(so, dbg["states"], norm)
    assert np.abs(e[1]).max() < tol_first and np.abs(e).max() < tol_all, (np.abs(e[1]).max(), np.abs(e).max())
    assert np.corrcoef(co, dbg["costs"])[0, 1] > 0.98
    rec = pl.plan(state, inj=inj)[0]
    assert rec["index"] == int(np.argmin(co))
    assert np.allclose(rec["action"][:6], ao[0, rec["index"]])
    regret = (dbg["costs"][rec["index"]] - dbg["costs"].min()) / dbg["costs"].std()
    assert regret < 0.5, regret
    # production flavour (device-drawn randomness) == debug flavour replaying the drawn tensor, bit for bit
    prod = pl.plan(state, seed=11)[0]
    again = pl.plan(state, inj=pl.sample_injection(11)[0])[0]
    assert prod["index"] == again["index"] and prod["cost"] == again["cost"] and np.array_equal(prod["action"], again["action"])
    assert not np.array_equal(pl.plan(state, seed=12)[0]["action"], prod["action"])
    return float(regret)


@pytest.mark.parametrize("S,activation,proposal,hid,ln,reward", [
    (500, "relu", "policy_gaussian", (256, 256), False, False),      # BASELINE configs[3] shape with MlpPolicy(hid_size=256)
    (256, "relu", "policy_mix", (192, 256), False, False),           # the reference's default dynamics size; ragged first layer
    (330, "relu", "policy_mix", (256, 130), True, False),            # LayerNorm dynamics (FLAGS.LAYER_NORM) + three chunks in layer 2
    (500, "tanh", "policy_mix", (256, 256), False, True),            # train_mpc_ppo_r.py:142: learned-reward model + hid_size 256
    (256, "tanh", "policy_gaussian", (256, 256), True, True),        # ... with its three LayerNorms
    (320, "tanh", "policy_mix", (256, 256), False, False),           # plain tanh NNDynamicsModel
    (64, "relu", "policy_gaussian", (129, 64), False, False),        # one unit over the 128 limit of the regular variants
])
def test_wide_proposal_network_fast_mode(S, activation, proposal, hid, ln, reward):
    K, H = 1000, 6
    rng, norm, state, pl, mpc, _ = _build(71, S, activation, proposal, hid, ln, reward, K, H)
    regret = _check(rng, norm, state, pl, mpc, proposal, K, H, 0.1 if (ln or activation == "tanh") else 0.06, 0.6 if (ln or activation == "tanh") else 0.3)
    print(dict(S=S, activation=activation, proposal=proposal, policy_hidden=hid, layer_norm=ln, reward=reward, regret_in_cost_std=regret))
    pl.close()


def test_wide_proposal_network_configs3_shape():
    """BASELINE configs[3] (K=8192 x H=15, 2x500 ReLU dynamics) with a 256-unit proposal network."""
    K, H = 8192, 15
    rng, norm, state, pl, mpc, _ = _build(21, 500, "relu", "policy_gaussian", (256, 256), False, False, K, H)
    _check(rng, norm, state, pl, mpc, "policy_gaussian", K, H, 0.06, 0.3)
    pl.close()


@pytest.mark.parametrize("S", [500, 256])
def test_switching_between_narrow_and_wide_proposal_networks(S):
    """A wider network selects another kernel variant whose weight ring has another stage size: set_policy re-packs the
    dynamics images, in whichever order the uploads arrive, and the plans equal those of a fresh planner."""
    K, H = 700, 5
    rng, norm, state, pl, mpc, wide = _build(5, S, "relu", "policy_mix", (256, 256), False, False, K, H, policy_first=True)
    inj = f32_uniform(rng, (H, K, 6))
    r_wide = pl.plan(state, inj=inj)[0]
    _, _, _, pl2, _, _ = _build(5, S, "relu", "policy_mix", (256, 256), False, False, K, H, policy_first=False)
    r_wide2 = pl2.plan(state, inj=inj)[0]
    assert r_wide["index"] == r_wide2["index"] and r_wide["cost"] == r_wide2["cost"]
    # down to a 128-unit network (the regular variants; at S = 256 the narrow-net kernel family) ...
    pw, logstd, rms_mean, rms_std = O.synthetic_policy(np.random.default_rng(9), hidden=(128, 128), norm=norm)
    pl.set_policy(pw, logstd, rms_mean, rms_std, clip=5.0)
    r_narrow = pl.plan(state, inj=inj)[0]
    _, _, _, pl3, _, _ = _build(5, S, "relu", "policy_mix", (128, 128), False, False, K, H)
    pl3.set_policy(pw, logstd, rms_mean, rms_std, clip=5.0)
    r_narrow3 = pl3.plan(state, inj=inj)[0]
    assert r_narrow["index"] == r_narrow3["index"] and r_narrow["cost"] == r_narrow3["cost"]
    # ... and back up
    pl.set_policy(*wide, clip=5.0)
    r_back = pl.plan(state, inj=inj)[0]
    assert r_back["index"] == r_wide["index"] and r_back["cost"] == r_wide["cost"] and np.array_equal(r_back["action"], r_wide["action"])
    for p in (pl, pl2, pl3):
        p.close()


def test_proposal_network_wider_than_256_is_rejected_in_fast_mode():
    from bc_mpc_b200 import Planner
    rng, norm, weights, state = O.synthetic_problem(3, (500, 500))
    pw, logstd, rms_mean, rms_std = O.synthetic_policy(rng, hidden=(320, 64), norm=norm)
    pl = Planner(20, 6, 2, 500, "relu", num_paths=256, horizon=3, precision="bf16", proposal="policy_mix")
    with pytest.raises(RuntimeError, match="<= 256"):
        pl.set_policy(pw, logstd, rms_mean, rms_std, clip=5.0)
    pl.close()


def test_wide_proposal_network_device_handoff_and_class_switch_on_a_side_stream():
    """bcmpc_set_policy_device with a 256-unit network: the hand-off (torch CUDA tensors, nn.Linear layout) that switches the
    planner from a 128-unit to a 256-unit network re-packs the dynamics images on the CALLER's stream, behind other work,
    and plans bit-identically to the host upload."""
    import torch
    from tests._util import make_pair
    K, H = 512, 5
    rng, norm, state, host, mpc = make_pair(71, (500, 500), K, H, precision="bf16", proposal="policy_gaussian")
    wide = O.synthetic_policy(rng, hidden=(256, 200), norm=norm)
    narrow = O.synthetic_policy(rng, hidden=(128, 96), norm=norm)
    host.set_policy(*wide, clip=5.0)
    want_wide = host.rollout_debug(state, seed=4)
    host.set_policy(*narrow, clip=5.0)
    want_narrow = host.rollout_debug(state, seed=4)
    assert not np.array_equal(want_wide[1], want_narrow[1])
    dev = torch.device("cuda:0")
    side = torch.cuda.Stream()

    def tensors(ws):
        return [(torch.from_numpy(np.ascontiguousarray(w.T)).to(dev), torch.from_numpy(b).to(dev)) for w, b in ws]
    tw, tn = tensors(wide[0]), tensors(narrow[0])
    torch.cuda.synchronize()
    _, _, _, pl, _ = make_pair(71, (500, 500), K, H, precision="bf16", proposal="policy_gaussian")      # starts with a 128-unit network
    with torch.cuda.stream(side):
        junk = torch.randn(4096, 4096, device=dev) @ torch.randn(4096, 4096, device=dev)
        pl.set_policy_device(tw, wide[1], wide[2], wide[3], clip=5.0, layout="out_in", stream_ptr=side.cuda_stream)
    got_wide = pl.rollout_debug(state, seed=4)
    with torch.cuda.stream(side):
        junk = junk @ junk
        pl.set_policy_device(tn, narrow[1], narrow[2], narrow[3], clip=5.0, layout="out_in", stream_ptr=side.cuda_stream)
    got_narrow = pl.rollout_debug(state, seed=4)
    del junk
    for a, b in zip(want_wide + want_narrow, got_wide + got_narrow):
        assert np.array_equal(a, b)
    pl.close(); host.close()
