"""BF16 fast mode (tcgen05 kernel): accuracy against the float64/float32 oracle, argmin agreement,
and bit-exact self-consistency properties (batching, sharding, sampling replay) at full sizes."""
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


@pytest.mark.parametrize("hidden", [(500, 500), (256, 256), (64, 64)])
def test_one_step_accuracy(hidden):
    K = 1000
    rng, norm, state, pl, mpc = make_pair(3, hidden, K, 1, precision="bf16")
    actions = f32_uniform(rng, (1, K, 6))
    so, ao, co = pl.rollout_debug(state, actions)
    ref_s, ref_a, ref_c = mpc.rollout(state, actions)
    assert np.array_equal(ao, actions)
    e = _norm_err(so[1], ref_s[1], norm)            # error of the network output in normalised-delta units
    assert np.abs(e).max() < 0.1, np.abs(e).max()
    assert np.sqrt((e ** 2).mean()) < 0.02
    assert np.abs(so[0] - ref_s[0]).max() < 1e-5     # root state only rounds float64 -> float32


@pytest.mark.parametrize("K,H", [(1000, 15), (4096, 15)])
def test_rollout_tracks_oracle_and_argmin_agreement(K, H):
    agree, regrets, worst, margins = 0, [], 0.0, []
    seeds = range(24)
    for seed in seeds:
        rng, norm, state, pl, mpc = make_pair(seed, (500, 500), K, H, precision="bf16")
        actions = f32_uniform(rng, (H, K, 6))
        act, dbg = mpc.get_action(state, actions, return_debug=True)
        so, ao, co = pl.rollout_debug(state, actions)
        e = _norm_err(so, dbg["states"], norm)
        worst = max(worst, float(np.abs(e).max()))
        rec = pl.plan(state, inj=actions)[0]
        assert rec["index"] == int(np.argmin(co))                      # plan == argmin of its own costs
        assert np.array_equal(rec["action"][:6], actions[0, rec["index"]])
        agree += int(rec["index"] == dbg["argmin"])
        spread = dbg["costs"].std()
        margins.append(float(np.diff(np.sort(dbg["costs"])[:2])[0] / spread))      # how close the runner-up is
        regrets.append(float((dbg["costs"][rec["index"]] - dbg["costs"][dbg["argmin"]]) / spread))
        assert np.corrcoef(co, dbg["costs"])[0, 1] > 0.98
    report = dict(K=K, H=H, seeds=len(seeds), argmin_agreement=agree / len(seeds), max_regret_in_cost_std=max(regrets),
                  worst_state_err_in_delta_std=worst, median_runner_up_margin_in_cost_std=float(np.median(margins)),
                  min_runner_up_margin_in_cost_std=float(min(margins)))
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/bf16_agreement_K{K}.json", "w") as f:
        json.dump(report, f)
    print(report)
    assert max(regrets) < 0.5       # the bf16 winner is within half a cost-std of the exact winner


@pytest.mark.parametrize("proposal,bc,K", [("policy_gaussian", False, 8192), ("policy_mix", False, 1000), ("bc_uniform", True, 1000)])
def test_policy_guided_fast_mode_tracks_oracle(proposal, bc, K):
    """BASELINE.json configs[3] (BC/policy-guided MPC, K=8192 x H=15) on the tcgen05 path: proposal network, noise
    mixing, rollout and cost fused; compared with the float32/float64 oracle on injected noise."""
    H = 15
    rng, norm, state, pl, mpc = make_pair(21, (500, 500), K, H, precision="bf16", proposal=proposal, bc=bc)
    inj = rng.standard_normal((H, K, 6)).astype(np.float32).astype(np.float64) if proposal == "policy_gaussian" \
        else f32_uniform(rng, (H, K, 6))
    act, dbg = mpc.get_action(state, inj, return_debug=True)
    so, ao, co = pl.rollout_debug(state, inj)
    assert np.abs(ao[0] - dbg["actions"][0]).max() < 0.03           # first proposals: same state, bf16 policy net
    assert np.abs(ao - dbg["actions"]).max() < 0.25                 # later ones drift with the bf16 rollout
    e = _norm_err(so, dbg["states"], norm)
    assert np.abs(e).max() < 0.25, np.abs(e).max()
    assert np.corrcoef(co, dbg["costs"])[0, 1] > 0.98
    rec = pl.plan(state, inj=inj)[0]
    assert rec["index"] == int(np.argmin(co))
    assert np.allclose(rec["action"][:6], ao[0, rec["index"]])
    regret = (dbg["costs"][rec["index"]] - dbg["costs"].min()) / dbg["costs"].std()
    print(dict(proposal=proposal, K=K, argmin_match=bool(rec["index"] == dbg["argmin"]), regret_in_cost_std=float(regret),
               max_action_err=float(np.abs(ao - dbg["actions"]).max()), max_state_err_in_delta_std=float(np.abs(e).max())))
    assert regret < 0.5
    # device-side sampling: runs, is reproducible, and differs across seeds
    r1, r2, r3 = pl.plan(state, seed=5)[0], pl.plan(state, seed=5)[0], pl.plan(state, seed=6)[0]
    assert r1["index"] == r2["index"] and r1["cost"] == r2["cost"]
    assert not np.array_equal(r1["action"], r3["action"])


@pytest.mark.parametrize("S,proposal,K,gamma,ln", [(500, "uniform", 4096, 0.95, False), (500, "policy_mix", 1000, 1.0, False), (256, "uniform", 1000, 1.0, False),
                                                   (500, "uniform", 1000, 0.95, True), (500, "policy_gaussian", 1000, 1.0, True), (256, "policy_mix", 1000, 1.0, True)])
def test_learned_reward_fast_mode_tracks_oracle(S, proposal, K, gamma, ln):
    """NNDynamicsRewardModel + MPCcontrollerReward / MPCcontrollerPolicyNetReward (dynamics.py:121-238,
    controllers.py:90-158,289-363) on the tcgen05 path: shared tanh layer, both 500-wide tanh heads as ONE 1024-wide
    second layer, block-structured output layer (20 deltas + reward), discounted-reward objective in the epilogue."""
    from bc_mpc_b200 import Planner
    H = 10
    rng, norm, weights, rweights, state = O.synthetic_reward_problem(33, S)
    dyn = O.OracleRewardDynamics(weights, rweights, norm)
    pl = Planner(20, 6, 2, S, "tanh", num_paths=K, horizon=H, precision="bf16", proposal=proposal, explore=0.4)
    pl.set_dynamics(weights)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
    if proposal == "uniform":
        pl.plan(state, seed=1)          # (a plain tanh NNDynamicsModel has its own tensor-core variant: test_gpu_f_fast_holes.py)
    if os.environ.get("BCMPC_TC_CG") == "1" and proposal != "uniform" and S > 256:
        # single-CTA debug mode: one CTA cannot keep the two 512-wide heads AND a proposal network resident
        with pytest.raises(RuntimeError, match="weight ring"):
            pl.set_reward_head(rweights, norm["mean_reward"], norm["std_reward"], gamma)
        return
    lns = None
    if ln:      # what run.sh launches: LAYER_NORM=True (default), LEARN_REWARD=True, SELFEXP=True
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
    one = np.abs(_norm_err(so[1], dbg["states"][1], norm)).max()
    cstd = dbg["costs"].std()
    rec = pl.plan(state, inj=inj)[0]
    regret = (dbg["costs"][rec["index"]] - dbg["costs"].min()) / cstd
    print(dict(S=S, proposal=proposal, ln=ln, one_step_err_in_delta_std=float(one), max_state_err_in_delta_std=float(np.abs(e).max()),
               cost_corr=float(np.corrcoef(co, dbg["costs"])[0, 1]), max_cost_err_in_std=float(np.abs(co - dbg["costs"]).max() / cstd),
               argmin_match=bool(rec["index"] == dbg["argmin"]), regret_in_cost_std=float(regret)))
    assert one < (0.08 if ln else 0.05) and np.abs(e).max() < (0.6 if ln else 0.3)
    assert np.corrcoef(co, dbg["costs"])[0, 1] > 0.98
    assert rec["index"] == int(np.argmin(co)) and np.allclose(rec["action"][:6], ao[0, rec["index"]])
    assert regret < 0.5
    r1, r2 = pl.plan(state, seed=5)[0], pl.plan(state, seed=5)[0]
    assert r1["index"] == r2["index"] and r1["cost"] == r2["cost"]
    pl.set_dynamics(weights)            # a weight refresh keeps the reward head (and resets the dynamics LayerNorm)
    if ln:
        with pytest.raises(RuntimeError, match="all three LayerNorms"):
            pl.plan(state, seed=5)
        pl.set_layer_norm(lns[:2])
    assert pl.plan(state, seed=5)[0]["cost"] == r1["cost"]


@pytest.mark.parametrize("S", [500, 256, 64])
def test_layer_norm_fast_mode_tracks_oracle(S):
    """FLAGS.LAYER_NORM dynamics (dynamics.py:68-69) on the tcgen05 path: un-normalised activations through the MMAs,
    the per-row mean / variance correction applied in the next layer's epilogue."""
    K, H = 1000, 15
    rng, norm, state, pl, mpc = make_pair(23, (S, S), K, H, precision="bf16")
    ln = [(1.0 + 0.1 * rng.standard_normal(S).astype(np.float32), 0.1 * rng.standard_normal(S).astype(np.float32)) for _ in range(2)]
    pl.set_layer_norm(ln)
    mpc.dyn_model.layer_norm = ln
    actions = f32_uniform(rng, (H, K, 6))
    act, dbg = mpc.get_action(state, actions, return_debug=True)
    so, ao, co = pl.rollout_debug(state, actions)
    e = _norm_err(so, dbg["states"], norm)
    print(dict(S=S, max_state_err_in_delta_std=float(np.abs(e).max()), one_step=float(np.abs(e[1]).max())))
    assert np.abs(e[1]).max() < 0.1
    assert np.abs(e).max() < 0.5
    assert np.corrcoef(co, dbg["costs"])[0, 1] > 0.98
    rec = pl.plan(state, inj=actions)[0]
    assert rec["index"] == int(np.argmin(co))
    regret = (dbg["costs"][rec["index"]] - dbg["costs"].min()) / dbg["costs"].std()
    assert regret < 0.5
    pl.set_layer_norm(None)                          # back to the plain network
    mpc.dyn_model.layer_norm = None
    so2, _, _ = pl.rollout_debug(state, actions)
    assert np.abs(_norm_err(so2, mpc.rollout(state, actions)[0], norm)).max() < 0.25


@pytest.mark.parametrize("S,proposal", [(500, "policy_gaussian"), (500, "policy_mix"), (256, "policy_gaussian")])
def test_layer_norm_with_policy_fast_mode_tracks_oracle(S, proposal):
    """The reference's default run (train_mpc_ppo.py:52 LAYER_NORM=True with MPCcontrollerPolicyNet) on the tcgen05 path."""
    K, H = 1000, 15
    rng, norm, state, pl, mpc = make_pair(27, (S, S), K, H, precision="bf16", proposal=proposal)
    ln = [(1.0 + 0.1 * rng.standard_normal(S).astype(np.float32), 0.1 * rng.standard_normal(S).astype(np.float32)) for _ in range(2)]
    pl.set_layer_norm(ln)
    mpc.dyn_model.layer_norm = ln
    inj = rng.standard_normal((H, K, 6)).astype(np.float32).astype(np.float64) if proposal == "policy_gaussian" else f32_uniform(rng, (H, K, 6))
    act, dbg = mpc.get_action(state, inj, return_debug=True)
    so, ao, co = pl.rollout_debug(state, inj)
    e = _norm_err(so, dbg["states"], norm)
    rec = pl.plan(state, inj=inj)[0]
    regret = (dbg["costs"][rec["index"]] - dbg["costs"].min()) / dbg["costs"].std()
    print(dict(S=S, proposal=proposal, max_state_err_in_delta_std=float(np.abs(e).max()), max_action_err=float(np.abs(ao - dbg["actions"]).max()),
               cost_corr=float(np.corrcoef(co, dbg["costs"])[0, 1]), argmin_match=bool(rec["index"] == dbg["argmin"]), regret_in_cost_std=float(regret)))
    assert np.abs(ao[0] - dbg["actions"][0]).max() < 0.03
    assert np.abs(e).max() < 0.6 and np.corrcoef(co, dbg["costs"])[0, 1] > 0.98
    assert rec["index"] == int(np.argmin(co)) and regret < 0.5
    pl.set_layer_norm(None)             # back to the plain policy-guided kernel
    so2, _, _ = pl.rollout_debug(state, inj)
    assert np.abs(so2 - so).max() > 1e-3


@pytest.mark.parametrize("variant", ["policy_gaussian", "policy_mix", "bc_uniform", "layer_norm", "layer_norm+policy", "reward", "reward+policy",
                                     "reward+layer_norm", "reward+layer_norm+policy"])
def test_production_flavour_equals_debug_flavour(variant):
    """Every kernel variant has two flavours of the same source: production (device-drawn randomness, records only) and
    debug (injected tensor, per-path outputs).  A device-sampled plan must equal, bit for bit, the plan replayed from the
    tensor bcmpc_sample_injection reports for the same seed."""
    from bc_mpc_b200 import Planner
    K, H, S, B = 700, 5, 500, 2
    rew, ln = "reward" in variant, "layer_norm" in variant
    proposal = variant if variant in ("policy_gaussian", "policy_mix", "bc_uniform") else ("policy_gaussian" if "policy" in variant else "uniform")
    rng, norm, weights, rweights, state = O.synthetic_reward_problem(61, S)
    if os.environ.get("BCMPC_TC_CG") == "1" and rew and proposal != "uniform":
        pytest.skip("single-CTA debug mode cannot hold the reward heads and a proposal network")
    pl = Planner(20, 6, 2, S, "tanh" if rew else "relu", num_paths=K, horizon=H, max_states=B, precision="bf16", proposal=proposal, explore=0.3)
    pl.set_dynamics(weights)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
    lns = [(1.0 + 0.1 * rng.standard_normal(S).astype(np.float32), 0.1 * rng.standard_normal(S).astype(np.float32)) for _ in range(3)]
    if ln:
        pl.set_layer_norm(lns[:2])
    if rew:
        pl.set_reward_head(rweights, norm["mean_reward"], norm["std_reward"], 0.97, layer_norm=lns[2] if ln else None)
    if proposal == "bc_uniform":
        pl.set_policy(O.glorot_mlp(rng, (20, 128, 64, 6), 0.05))
    elif proposal != "uniform":
        pw, logstd, rms_mean, rms_std = O.synthetic_policy(rng, hidden=(128, 128), norm=norm)
        pl.set_policy(pw, logstd, rms_mean, rms_std, clip=5.0)
    states = norm["mean_obs"] + norm["std_obs"] * rng.standard_normal((B, 20))
    prod = pl.plan(states, seed=123)
    dbg = pl.plan(states, inj=pl.sample_injection(123, B))
    assert np.array_equal(prod["index"], dbg["index"]) and np.array_equal(prod["cost"], dbg["cost"])
    assert np.array_equal(prod["action"], dbg["action"])
    assert not np.array_equal(pl.plan(states, seed=124)["action"], prod["action"])


def test_fast_mode_sampling_batching_sharding_are_bit_consistent():
    import torch
    from bc_mpc_b200 import _lib as L
    K, H, B = 4096, 20, 3                      # BASELINE.json configs[2] shape per state (fewer states)
    rng, norm, state, pl, mpc = make_pair(9, (500, 500), K, H, precision="bf16", max_states=B)
    states = norm["mean_obs"] + norm["std_obs"] * rng.standard_normal((B, 20))
    recs = pl.plan(states, seed=77)
    inj = pl.sample_injection(77, B)
    for b in range(B):
        _, _, _, p1, _ = make_pair(9, (500, 500), K, H, precision="bf16")
        r = p1.plan(states[b], inj=inj[b])[0]
        assert r["index"] == recs[b]["index"] and r["cost"] == recs[b]["cost"]
        assert np.array_equal(r["action"], recs[b]["action"])
    # shard the K paths of state 0 over 3 handles: same winner, bit-identical cost
    bounds = [0, 1000, 2500, K]
    gathered = torch.zeros(3 * 80, dtype=torch.uint8, device="cuda")
    out = torch.zeros(80, dtype=torch.uint8, device="cuda")
    d_state = torch.tensor(states[0], dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    # state index is part of the Philox counter: shard handles plan state 0 of a 1-state batch
    want = make_pair(9, (500, 500), K, H, precision="bf16")[3].plan(states[0], seed=5)[0]
    keep = []
    for g in range(3):
        sp = make_pair(9, (500, 500), bounds[g + 1] - bounds[g], H, precision="bf16", path_offset=bounds[g], total_paths=K)[3]
        sp.plan_device(d_state.data_ptr(), 1, gathered.data_ptr() + g * 80, seed=5)
        keep.append(sp)
    torch.cuda.synchronize()
    pl.reduce_records_device(gathered.data_ptr(), 3, 1, out.data_ptr())
    torch.cuda.synchronize()
    got = np.frombuffer(out.cpu().numpy().tobytes(), L.RECORD_DTYPE)[0]
    assert got["index"] == want["index"] and got["cost"] == want["cost"] and np.array_equal(got["action"], want["action"])


def test_fast_mode_large_k_properties():
    """Full-size property checks (oracle too slow here): winner's recorded cost equals the cost of
    re-rolling that single path, and duplicated paths tie to the lowest index."""
    K, H = 131072, 20            # one GPU's share of BASELINE.json configs[2]
    rng, norm, state, pl, mpc = make_pair(13, (500, 500), K, H, precision="bf16")
    rec = pl.plan(state, seed=42)[0]
    assert 0 <= rec["index"] < K
    # replay only the winning path - its device-drawn action sequence - through a one-path bf16 planner (bit-exact cost) and
    # through the ORACLE (trajectory within one delta-sigma, cost difference explained by the drift of x17 / threshold flips)
    _, _, _, one, mpc1 = make_pair(13, (500, 500), 1, H, precision="bf16", path_offset=int(rec["index"]), total_paths=K)
    inj1 = one.sample_injection(42)[0]
    assert np.array_equal(inj1[0, 0], rec["action"][:6])
    print("winner replay (worst delta-sigma, |cost diff|, threshold flips):", check_winner_replay(rec, one, mpc1, state, inj1, norm))
    # five hidden layers: beyond every fast kernel (the cluster-split kernel takes 2..4), verification mode only
    from bc_mpc_b200 import Planner as P
    with pytest.raises(RuntimeError, match="n_hidden|hidden layers"):
        P(20, 6, 5, 64, "relu", num_paths=128, horizon=2, precision="bf16")


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_device_weight_handoff_matches_host_upload(precision):
    """bcmpc_set_dynamics_device (refresh after dyn_model.fit without a host round trip): torch CUDA
    tensors in TF [in,out] or nn.Linear [out,in] layout give bit-identical plans to the host upload; the hand-off is
    asynchronous on the trainer's stream and later plans are ordered after it."""
    import torch
    from bc_mpc_b200 import Planner
    K, H, S = 1024, 6, 500
    rng, norm, weights, state = O.synthetic_problem(51, (S, S))
    _, _, weights2, _ = O.synthetic_problem(52, (S, S))

    def planner():
        pl = Planner(20, 6, 2, S, "relu", num_paths=K, horizon=H, precision=precision)
        pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
        return pl

    host = planner()
    host.set_dynamics(weights)
    want1 = host.rollout_debug(state, seed=3)
    host.set_dynamics(weights2)
    want2 = host.rollout_debug(state, seed=3)
    assert not np.array_equal(want1[2], want2[2])

    dev = torch.device("cuda:0")
    side = torch.cuda.Stream()
    for layout in ("in_out", "out_in"):
        def tensors(ws):
            return [(torch.from_numpy(w if layout == "in_out" else np.ascontiguousarray(w.T)).to(dev), torch.from_numpy(b).to(dev)) for w, b in ws]
        t1, t2 = tensors(weights), tensors(weights2)
        torch.cuda.synchronize()
        pl = planner()
        pl.set_dynamics_device(t1, layout=layout)
        got1 = pl.rollout_debug(state, seed=3)
        with torch.cuda.stream(side):                      # refresh enqueued on the trainer's stream, behind some work
            junk = torch.randn(4096, 4096, device=dev) @ torch.randn(4096, 4096, device=dev)
            pl.set_dynamics_device(t2, layout=layout, stream_ptr=side.cuda_stream)
        got2 = pl.rollout_debug(state, seed=3)             # on the handle's own stream: must wait for the hand-off
        del junk
        for a, b in zip(want1, got1):
            assert np.array_equal(a, b)
        for a, b in zip(want2, got2):
            assert np.array_equal(a, b)
        r_host, r_dev = host.plan(state, seed=5)[0], pl.plan(state, seed=5)[0]
        assert r_host["index"] == r_dev["index"] and r_host["cost"] == r_dev["cost"]
    with pytest.raises(ValueError):
        pl.set_dynamics_device([(w.cpu(), b) for w, b in t1])
    # through the reference-shaped classes: NNDynamicsModel.set_weights_device -> MPCcontroller
    from bc_mpc_b200 import controllers, cost_functions, dynamics

    class Space:
        def __init__(self, n):
            self.low, self.high, self.shape = -np.ones(n), np.ones(n), (n,)

    class Env:
        action_space, observation_space = Space(6), Space(20)

    n10 = (norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], None, None, None, None,
           norm["mean_deltas"], norm["std_deltas"])
    dm = dynamics.NNDynamicsModel(Env(), 2, S, "relu", None, n10)
    c = controllers.MPCcontroller(Env(), dm, horizon=H, cost_fn=cost_functions.cheetah_cost_fn, num_simulated_paths=K, precision=precision)
    for ws, ts in ((weights, t1), (weights2, t2)):       # t1, t2 are in nn.Linear layout after the loop above
        dm.set_weights_device(ts, layout="out_in")
        np.random.seed(3)
        a_dev = c.get_action(state)
        dm.set_weights(ws)
        np.random.seed(3)
        assert np.array_equal(a_dev, c.get_action(state))


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_device_policy_handoff_matches_host_upload(precision):
    """bcmpc_set_policy_device: the proposal network refreshed from torch CUDA tensors (TF or nn.Linear layout) plans
    bit-identically to the host upload, also when enqueued on another stream behind other work."""
    import torch
    K, H = 512, 5
    rng, norm, state, host, mpc = make_pair(71, (500, 500), K, H, precision=precision, proposal="policy_gaussian")
    pw, logstd, rms_mean, rms_std = O.synthetic_policy(rng, hidden=(128, 96), norm=norm)
    pw2, logstd2, _, _ = O.synthetic_policy(rng, hidden=(128, 96), norm=norm)
    host.set_policy(pw, logstd, rms_mean, rms_std, clip=5.0)
    want1 = host.rollout_debug(state, seed=4)
    host.set_policy(pw2, logstd2, rms_mean, rms_std, clip=5.0)
    want2 = host.rollout_debug(state, seed=4)
    assert not np.array_equal(want1[1], want2[1])
    dev = torch.device("cuda:0")
    side = torch.cuda.Stream()
    for layout in ("in_out", "out_in"):
        def tensors(ws):
            return [(torch.from_numpy(w if layout == "in_out" else np.ascontiguousarray(w.T)).to(dev), torch.from_numpy(b).to(dev)) for w, b in ws]
        t1, t2 = tensors(pw), tensors(pw2)
        torch.cuda.synchronize()
        _, _, _, pl, _ = make_pair(71, (500, 500), K, H, precision=precision, proposal="policy_gaussian")
        pl.set_policy_device(t1, logstd, rms_mean, rms_std, clip=5.0, layout=layout)
        got1 = pl.rollout_debug(state, seed=4)
        with torch.cuda.stream(side):
            junk = torch.randn(4096, 4096, device=dev) @ torch.randn(4096, 4096, device=dev)
            pl.set_policy_device(t2, logstd2, rms_mean, rms_std, clip=5.0, layout=layout, stream_ptr=side.cuda_stream)
        got2 = pl.rollout_debug(state, seed=4)
        del junk
        for a, b in zip(want1 + want2, got1 + got2):
            assert np.array_equal(a, b)
        assert pl.plan(state, seed=8)[0]["cost"] == host.plan(state, seed=8)[0]["cost"]


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_device_handoff_of_the_whole_reward_model(precision):
    """dynamics + LayerNorms + reward head + reward LayerNorm + proposal network, all from torch CUDA tensors in nn.Linear
    layout through the reference-shaped classes, against the host upload of the same parameters (bit-identical plans)."""
    import torch
    from bc_mpc_b200 import controllers, dynamics, policy

    class Space:
        def __init__(self, n):
            self.low, self.high, self.shape = -np.ones(n), np.ones(n), (n,)

    class Env:
        action_space, observation_space = Space(6), Space(20)

    K, H, S = 384, 5, 500
    rng, norm, weights, rweights, state = O.synthetic_reward_problem(81, S)
    lns = [(1.0 + 0.1 * rng.standard_normal(S).astype(np.float32), 0.1 * rng.standard_normal(S).astype(np.float32)) for _ in range(3)]
    pw, logstd, rms_mean, rms_std = O.synthetic_policy(rng, hidden=(128, 128), norm=norm)
    n10 = (norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_reward"], norm["std_reward"], None, None,
           norm["mean_deltas"], norm["std_deltas"])
    dev = torch.device("cuda:0")

    def lin(ws):       # nn.Linear layout
        return [(torch.from_numpy(np.ascontiguousarray(w.T)).to(dev), torch.from_numpy(b).to(dev)) for w, b in ws]

    def vec(ps):
        return [(torch.from_numpy(g).to(dev), torch.from_numpy(b).to(dev)) for g, b in ps]

    def controller(device_side):
        dm = dynamics.NNDynamicsRewardModel(Env(), n10, size=S, layer_norm=True)
        pol = policy.MlpPolicy(None, Env(), 128, 2)
        if device_side:
            dm.set_weights_device(lin(weights), lin(rweights), layer_norm_tensors=vec(lns[:2]), reward_layer_norm_tensors=vec(lns[2:])[0])
            pol.set_weights_device(lin(pw), logstd, rms_mean, rms_std)
        else:
            dm.set_weights(weights, rweights, lns[:2], lns[2])
            pol.set_weights(pw, logstd, rms_mean, rms_std)
        return controllers.MPCcontrollerPolicyNetReward(Env(), dm, pol, self_exp=True, horizon=H, num_simulated_paths=K, precision=precision), dm

    (ch, dmh), (cd, dmd) = controller(False), controller(True)
    if os.environ.get("BCMPC_TC_CG") == "1" and precision == "bf16":
        pytest.skip("single-CTA debug mode cannot hold the reward heads and a proposal network")
    for seed in (1, 2):
        np.random.seed(seed)
        a_host = ch.get_action(state)
        np.random.seed(seed)
        a_dev = cd.get_action(state)
        assert np.array_equal(a_host, a_dev) and ch.last_record["cost"] == cd.last_record["cost"]
    nh, rh = dmh.predict(state[None], np.zeros((1, 6)))
    nd, rd = dmd.predict(state[None], np.zeros((1, 6)))
    assert np.array_equal(nh, nd) and np.array_equal(rh, rd)


def test_fast_mode_nan_and_all_equal_costs():
    """np.argmin semantics in the fast kernel's reductions: the first NaN wins; identical costs -> lowest index."""
    K, H = 700, 3
    rng, norm, state, pl, mpc = make_pair(91, (500, 500), K, H, precision="bf16")
    bad = state.copy()
    bad[3] = np.nan                                     # every path sees NaN from step 0 on
    rec = pl.plan(bad, seed=1)[0]
    assert np.isnan(rec["cost"]) and rec["index"] == 0
    inj = np.tile(f32_uniform(rng, (H, 1, 6)), (1, K, 1))     # K identical action sequences: a K-way tie
    rec = pl.plan(state, inj=inj)[0]
    _, _, co = pl.rollout_debug(state, inj)
    assert np.all(co == co[0]) and rec["index"] == 0 and rec["cost"] == co[0]


def test_two_devices_in_one_process():
    """One process, planners on two GPUs (per-device kernel attributes, streams and buffers): identical plans."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from bc_mpc_b200 import Planner
    rng, norm, weights, state = O.synthetic_problem(13, (500, 500))
    recs = []
    for dev in (0, 1, 0):
        for precision in ("bf16", "fp32"):
            pl = Planner(20, 6, 2, 500, "relu", num_paths=600, horizon=4, precision=precision, device=dev)
            pl.set_dynamics(weights)
            pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
            recs.append((precision, pl.plan(state, seed=3)[0]))
            pl.close()
    for precision in ("bf16", "fp32"):
        same = [r for p, r in recs if p == precision]
        assert all(r["index"] == same[0]["index"] and r["cost"] == same[0]["cost"] for r in same)


def _narrow_pair(monkeypatch, mode, seed, S, K, H, **kw):
    # BCMPC_NARROW: 0 = wide kernel only, 1 = per launch (narrow when the wide kernel would need more than one wave), 2 = narrow
    monkeypatch.setenv("BCMPC_NARROW", mode)
    return make_pair(seed, (S, S), K, H, precision="bf16", **kw)


@pytest.mark.parametrize("S", [256, 200, 128, 64, 7])
def test_narrow_kernel_tracks_oracle_and_matches_the_wide_kernel(monkeypatch, S):
    """rollout_narrow_kernel (hidden <= 256: two 128-path tiles per CTA, all weights resident) against the oracle, and -
    where the wide kernel uses the same explicit-bias weight images (hidden 64 / 128 / 256) - bit for bit against it."""
    if os.environ.get("BCMPC_TC_CG") == "1":
        pytest.skip("the narrow kernel is a CTA-pair kernel")
    K, H = 1000, 12
    rng, norm, state, pl, mpc = _narrow_pair(monkeypatch, "2", 15, S, K, H)
    actions = f32_uniform(rng, (H, K, 6))
    act, dbg = mpc.get_action(state, actions, return_debug=True)
    so, ao, co = pl.rollout_debug(state, actions)
    e = _norm_err(so, dbg["states"], norm)
    rec = pl.plan(state, inj=actions)[0]
    regret = (dbg["costs"][rec["index"]] - dbg["costs"].min()) / dbg["costs"].std()
    print(dict(S=S, max_state_err_in_delta_std=float(np.abs(e).max()), cost_corr=float(np.corrcoef(co, dbg["costs"])[0, 1]),
               argmin_match=bool(rec["index"] == dbg["argmin"]), regret=float(regret)))
    assert np.array_equal(ao, actions) and np.abs(e).max() < 0.3 and np.corrcoef(co, dbg["costs"])[0, 1] > 0.98
    assert rec["index"] == int(np.argmin(co)) and np.array_equal(rec["action"][:6], actions[0, rec["index"]]) and regret < 0.5
    # production flavour == debug flavour; batching; ragged K (tile count padded to a multiple of 4)
    B = 3
    _, _, _, plb, _ = _narrow_pair(monkeypatch, "2", 15, S, 700, 5, max_states=B)
    states = norm["mean_obs"] + norm["std_obs"] * rng.standard_normal((B, 20))
    prod = plb.plan(states, seed=31)
    inj = plb.sample_injection(31, B)
    dbgr = plb.plan(states, inj=inj)
    assert np.array_equal(prod["index"], dbgr["index"]) and np.array_equal(prod["cost"], dbgr["cost"])
    for K1 in (1, 129):
        _, _, _, p1, _ = _narrow_pair(monkeypatch, "2", 15, S, K1, 3)
        r = p1.plan(state, seed=2)[0]
        assert 0 <= r["index"] < K1
    # K sharded over three handles (path_offset / total_paths, as on three GPUs): the best shard record is the single plan
    want = pl.plan(state, seed=5)[0]
    best = None
    for lo, hi in ((0, 300), (300, 812), (812, K)):
        _, _, _, sh, _ = _narrow_pair(monkeypatch, "2", 15, S, hi - lo, H, path_offset=lo, total_paths=K)
        r = sh.plan(state, seed=5)[0]
        if best is None or r["cost"] < best["cost"] or (r["cost"] == best["cost"] and r["index"] < best["index"]):
            best = r
    assert best["index"] == want["index"] and best["cost"] == want["cost"] and np.array_equal(best["action"], want["action"])
    # against the wide kernel
    _, _, _, wide, _ = _narrow_pair(monkeypatch, "0", 15, S, K, H)
    sw, _, cw = wide.rollout_debug(state, actions)
    if S in (64, 128, 256):
        assert np.array_equal(sw, so) and np.array_equal(cw, co)
    else:
        assert np.abs(_norm_err(sw, so, norm)).max() < 0.3


@pytest.mark.parametrize("S", [256, 150, 64])
def test_narrow_kernel_layer_norm(monkeypatch, S):
    """LayerNorm dynamics on the narrow-net kernel (per-path statistics stay in the registers of the row's thread):
    against the oracle, against the wide kernel, production flavour == debug flavour."""
    if os.environ.get("BCMPC_TC_CG") == "1":
        pytest.skip("the narrow kernel is a CTA-pair kernel")
    K, H = 1000, 10
    ln = None

    def build(mode):
        nonlocal ln
        rng, norm, state, pl, mpc = _narrow_pair(monkeypatch, mode, 19, S, K, H)
        ln = [(1.0 + 0.1 * rng.standard_normal(S).astype(np.float32), 0.1 * rng.standard_normal(S).astype(np.float32)) for _ in range(2)]
        pl.set_layer_norm(ln)
        mpc.dyn_model.layer_norm = ln
        return rng, norm, state, pl, mpc

    rng, norm, state, pl, mpc = build("2")
    actions = f32_uniform(rng, (H, K, 6))
    act, dbg = mpc.get_action(state, actions, return_debug=True)
    so, ao, co = pl.rollout_debug(state, actions)
    e = _norm_err(so, dbg["states"], norm)
    rec = pl.plan(state, inj=actions)[0]
    regret = (dbg["costs"][rec["index"]] - dbg["costs"].min()) / dbg["costs"].std()
    print(dict(S=S, max_state_err_in_delta_std=float(np.abs(e).max()), cost_corr=float(np.corrcoef(co, dbg["costs"])[0, 1]), regret=float(regret)))
    assert np.abs(e).max() < 0.6 and np.corrcoef(co, dbg["costs"])[0, 1] > 0.98 and regret < 0.5
    prod = pl.plan(state, seed=77)[0]
    dbgr = pl.plan(state, inj=pl.sample_injection(77)[0])[0]
    assert prod["index"] == dbgr["index"] and prod["cost"] == dbgr["cost"]
    _, _, _, wide, _ = build("0")
    sw, _, cw = wide.rollout_debug(state, actions)
    assert np.abs(_norm_err(sw, so, norm)).max() < 0.3          # (the two kernels sum the statistics in different orders)


@pytest.mark.parametrize("S,proposal,ln", [(256, "policy_gaussian", False), (256, "policy_mix", True), (128, "bc_uniform", False),
                                          (200, "policy_gaussian", True)])
def test_narrow_kernel_policy_guided(monkeypatch, S, proposal, ln):
    """Policy-guided shooting (and LayerNorm + policy, the reference's default run at size 256) on the narrow-net kernel:
    against the oracle, against the wide kernel, production flavour == debug flavour."""
    if os.environ.get("BCMPC_TC_CG") == "1":
        pytest.skip("the narrow kernel is a CTA-pair kernel")
    K, H = 1000, 10
    lns = None

    def build(mode):
        nonlocal lns
        monkeypatch.setenv("BCMPC_NARROW", mode)
        rng, norm, state, pl, mpc = make_pair(23, (S, S), K, H, precision="bf16", proposal=proposal, bc=proposal == "bc_uniform")
        if ln:
            lns = [(1.0 + 0.1 * rng.standard_normal(S).astype(np.float32), 0.1 * rng.standard_normal(S).astype(np.float32)) for _ in range(2)]
            pl.set_layer_norm(lns)
            mpc.dyn_model.layer_norm = lns
        return rng, norm, state, pl, mpc

    rng, norm, state, pl, mpc = build("2")
    inj = rng.standard_normal((H, K, 6)).astype(np.float32).astype(np.float64) if proposal == "policy_gaussian" else f32_uniform(rng, (H, K, 6))
    act, dbg = mpc.get_action(state, inj, return_debug=True)
    so, ao, co = pl.rollout_debug(state, inj)
    e = _norm_err(so, dbg["states"], norm)
    rec = pl.plan(state, inj=inj)[0]
    regret = (dbg["costs"][rec["index"]] - dbg["costs"].min()) / dbg["costs"].std()
    print(dict(S=S, proposal=proposal, ln=ln, first_action_err=float(np.abs(ao[0] - dbg["actions"][0]).max()),
               max_state_err_in_delta_std=float(np.abs(e).max()), cost_corr=float(np.corrcoef(co, dbg["costs"])[0, 1]), regret=float(regret)))
    assert np.abs(ao[0] - dbg["actions"][0]).max() < 0.03
    assert np.abs(e).max() < 0.6 and np.corrcoef(co, dbg["costs"])[0, 1] > 0.98 and regret < 0.5
    assert rec["index"] == int(np.argmin(co)) and np.allclose(rec["action"][:6], ao[0, rec["index"]])
    prod = pl.plan(state, seed=77)[0]
    dbgr = pl.plan(state, inj=pl.sample_injection(77)[0])[0]
    assert prod["index"] == dbgr["index"] and prod["cost"] == dbgr["cost"]
    _, _, _, wide, _ = build("0")
    sw, aw, cw = wide.rollout_debug(state, inj)
    assert np.array_equal(aw[0], ao[0])                            # same proposal network images: identical first proposals
    if S in (128, 256) and not ln:
        assert np.array_equal(sw, so) and np.array_equal(cw, co)   # ... and identical dynamics images: bit-identical rollouts
    else:
        assert np.abs(_norm_err(sw, so, norm)).max() < 0.3
