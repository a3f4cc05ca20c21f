"""Round-2 GPU tests of the host/boundary fixes: stream ordering of device hand-offs, current-device hygiene, the value
head of MlpPolicy, injected sampling through the batched entry point."""
import numpy as np
import pytest

from oracle import mpc_oracle as O
from tests._util import f32_uniform

pytestmark = pytest.mark.gpu


class _Space:
    def __init__(self, n):
        self.low, self.high, self.shape = -np.ones(n), np.ones(n), (n,)

    def sample(self):
        return np.random.uniform(self.low, self.high)


class _Env:
    action_space, observation_space = _Space(6), _Space(20)


def _planner(norm, S, K, H, precision, **kw):
    from bc_mpc_b200 import Planner
    pl = Planner(20, 6, 2, S, "relu", num_paths=K, horizon=H, precision=precision, **kw)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
    return pl


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_handoff_without_synchronize_is_ordered_after_the_producer(precision):
    """set_*_device with the default stream argument must be ordered after the torch kernels that produced the
    weights (no torch.cuda.synchronize() anywhere), and a dynamics hand-off on stream A followed by a policy hand-off on
    stream B must both be visible to a later plan / policy_act on the handle's own stream."""
    import torch
    K, H, S = 512, 5, 256
    rng, norm, w_old, state = O.synthetic_problem(61, (S, S))
    _, _, w_new, _ = O.synthetic_problem(62, (S, S))
    pw_old, logstd, rms_mean, rms_std = O.synthetic_policy(np.random.default_rng(7), hidden=(64, 64), norm=norm)
    pw_new, _, _, _ = O.synthetic_policy(np.random.default_rng(8), hidden=(64, 64), norm=norm)
    host = _planner(norm, S, K, H, precision, proposal="policy_mix", explore=0.5)
    host.set_dynamics(w_new)
    host.set_policy(pw_new, logstd, rms_mean, rms_std, clip=5.0)
    want = host.rollout_debug(state, seed=3)
    ob = norm["mean_obs"] + norm["std_obs"] * np.random.default_rng(3).standard_normal((33, 20))
    want_act = host.policy_act(ob)

    dev = torch.device("cuda:0")
    to_dev = lambda ws: [(torch.from_numpy(np.ascontiguousarray(w.T)).to(dev), torch.from_numpy(b).to(dev)) for w, b in ws]
    d_w, d_wnew = to_dev(w_old), to_dev(w_new)
    d_p, d_pnew = to_dev(pw_old), to_dev(pw_new)
    pl = _planner(norm, S, K, H, precision, proposal="policy_mix", explore=0.5)
    pl.set_dynamics_device(d_w, layout="out_in")
    pl.set_policy_device(d_p, logstd, rms_mean, rms_std, clip=5.0, layout="out_in")
    torch.cuda.synchronize()
    assert not np.array_equal(pl.rollout_debug(state, seed=3)[2], want[2])
    # "optimizer step": the parameters are overwritten by kernels queued behind ~50 ms of other work on torch's CURRENT
    # streams; the hand-offs follow immediately, with no synchronize and no explicit stream argument
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    junk = torch.randn(8192, 8192, device=dev)
    with torch.cuda.stream(sa):
        for _ in range(6):
            junk2 = junk @ junk
        for (w, b), (wn, bn) in zip(d_w, d_wnew):
            w.copy_(wn); b.copy_(bn)
        pl.set_dynamics_device(d_w, layout="out_in")          # default stream = torch's current stream = sa
    with torch.cuda.stream(sb):
        for _ in range(3):
            junk3 = junk @ junk
        for (w, b), (wn, bn) in zip(d_p, d_pnew):
            w.copy_(wn); b.copy_(bn)
        pl.set_policy_device(d_p, logstd, rms_mean, rms_std, clip=5.0, layout="out_in")   # on sb, chained after the hand-off on sa
    got_act = pl.policy_act(ob)                                # the handle's own stream waits for both hand-offs
    got = pl.rollout_debug(state, seed=3)
    del junk2, junk3
    assert np.array_equal(got_act, want_act)
    for a, b in zip(want, got):
        assert np.array_equal(a, b)
    r0, r1 = host.plan(state, seed=9)[0], pl.plan(state, seed=9)[0]
    assert r0["index"] == r1["index"] and r0["cost"] == r1["cost"]


def test_entry_points_restore_the_current_device():
    """no entry point may leave the thread's current CUDA device changed, and the helper planners behind
    predict() / act() / cheetah_cost_fn live on the CURRENT device, not on device 0."""
    import torch
    from bc_mpc_b200 import Planner, _lib, cost_functions, dynamics, policy
    n_dev = torch.cuda.device_count()
    rng, norm, weights, state = O.synthetic_problem(13, (64, 64))
    other = 1 if n_dev > 1 else 0
    torch.cuda.set_device(other)
    try:
        pl = Planner(20, 6, 2, 64, "relu", num_paths=256, horizon=3, precision="bf16", device=0)
        pl.set_dynamics(weights)
        pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
        pl.plan(state, seed=1)
        pl.predict(state[None], np.zeros((1, 6)))
        assert _lib.current_device() == other == torch.cuda.current_device()
        pl.close()
        assert _lib.current_device() == other
        # helper planners follow the current device
        n10 = (norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], None, None, None, None,
               norm["mean_deltas"], norm["std_deltas"])
        dm = dynamics.NNDynamicsModel(_Env(), 2, 64, "relu", None, n10)
        dm.set_weights(weights)
        dm.predict(state[None], np.zeros((1, 6)))
        assert dm._predictor.cfg.device == other
        cost_functions.cheetah_cost_fn(state[None], None, state[None])
        assert other in cost_functions._cost_planners
        pw, logstd, rms_mean, rms_std = O.synthetic_policy(rng, hidden=(32, 32), norm=norm)
        pi = policy.MlpPolicy(env=_Env(), hid_size=32)
        pi.set_weights(pw, logstd, rms_mean, rms_std)
        pi.act(state, stochastic=False)
        assert pi._fwd.cfg.device == other and _lib.current_device() == other
    finally:
        torch.cuda.set_device(0)


def test_mlp_policy_value_head_and_the_segment_generator():
    """MlpPolicy.act returns (ac, vpred) like ppo_bc_policy.py:174-185 once the 'vf' branch is loaded, and
    traj_segment_generator runs with the package's own policy (it raised TypeError on vpred = None)."""
    from bc_mpc_b200 import controllers, cost_functions, dynamics, policy, sampling
    rng, norm, weights, state = O.synthetic_problem(21, (64, 64))
    pw, logstd, rms_mean, rms_std = O.synthetic_policy(rng, hidden=(32, 32), norm=norm)
    vw = O.glorot_mlp(rng, (20, 32, 32, 1), 0.05)
    pi = policy.MlpPolicy(env=_Env(), hid_size=32)
    pi.set_weights(pw, logstd, rms_mean, rms_std)
    ob = norm["mean_obs"] + norm["std_obs"] * rng.standard_normal((17, 20))
    assert pi.act(ob, stochastic=False)[1] is None
    pi.set_value_weights(vw)
    ac, vpred = pi.act(ob, stochastic=False)
    # oracle: the same filtered / clipped observation through the vf MLP (ppo_bc_policy.py:59-64)
    want_v = O.OraclePolicy(vw, np.zeros(1), rms_mean, rms_std).mean(ob)[:, 0]
    want_a = O.OraclePolicy(pw, logstd, rms_mean, rms_std).mean(ob)
    assert vpred.shape == (17,) and np.allclose(vpred, want_v, rtol=1e-5, atol=1e-6)
    assert np.allclose(ac, want_a, rtol=1e-5, atol=1e-6)
    # load_vars picks the vf branch up by its TF names
    vd = {}
    for i, n in enumerate(("fc1", "fc2", "final")):
        vd[f"pi/pi/pol/{n}/kernel:0"], vd[f"pi/pi/pol/{n}/bias:0"] = pw[i]
        vd[f"pi/pi/vf/{n}/kernel:0"], vd[f"pi/pi/vf/{n}/bias:0"] = vw[i]
    vd["pi/pi/pol/logstd:0"] = logstd
    vd["pi/pi/obfilter/runningsum:0"], vd["pi/pi/obfilter/runningsumsq:0"], vd["pi/pi/obfilter/count:0"] = np.zeros(20), np.ones(20), 1.0
    pi2 = policy.MlpPolicy(env=_Env(), hid_size=32)
    pi2.load_vars(vd)
    assert pi2.act(ob, stochastic=False)[1].shape == (17,)

    class Env(_Env):
        def __init__(self):
            self.s = state.copy()

        def reset(self):
            return self.s.copy()

        def step(self, a):
            self.s = self.s + 0.01 * np.resize(a, 20)
            return self.s.copy(), float(a[0]), False, {}

    n10 = (norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], None, None, None, None,
           norm["mean_deltas"], norm["std_deltas"])
    dm = dynamics.NNDynamicsModel(Env(), 2, 64, "relu", None, n10)
    dm.set_weights(weights)
    c = controllers.MPCcontrollerPolicyNet(Env(), dm, pi, explore=0.5, self_exp=False, horizon=3,
                                           cost_fn=cost_functions.cheetah_cost_fn, num_simulated_paths=128)
    seg = sampling.traj_segment_generator(pi, c, c, None, Env(), True, True, 4)
    assert seg["vpred"].shape == (4,) and np.all(np.isfinite(seg["vpred"])) and np.any(seg["vpred"] != 0)
    assert seg["ac"].shape == (4, 6) and seg["mpcac"].shape == (4, 6)
    pi_nov = policy.MlpPolicy(env=_Env(), hid_size=32)
    pi_nov.set_weights(pw, logstd, rms_mean, rms_std)
    with pytest.raises(RuntimeError, match="value head"):
        sampling.traj_segment_generator(pi_nov, c, c, None, Env(), True, True, 2)


def test_get_actions_honours_an_overridden_sampler_and_plan_rejects_small_obs():
    """get_actions injects sample_random_actions() per root state when it is overridden, like get_action;
    planning on cheetah_cost_fn needs obs[17]."""
    from bc_mpc_b200 import Planner, controllers, cost_functions, dynamics
    K, H, B = 256, 4, 3
    rng, norm, weights, state = O.synthetic_problem(31, (64, 64))
    n10 = (norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], None, None, None, None,
           norm["mean_deltas"], norm["std_deltas"])
    dm = dynamics.NNDynamicsModel(_Env(), 2, 64, "relu", None, n10)
    dm.set_weights(weights)
    draws = [f32_uniform(rng, (H, K, 6)) for _ in range(B)]

    class Fixed(controllers.MPCcontroller):
        n = 0

        def sample_random_actions(self):
            Fixed.n += 1
            return draws[(Fixed.n - 1) % B]

    c = Fixed(_Env(), dm, horizon=H, cost_fn=cost_functions.cheetah_cost_fn, num_simulated_paths=K)
    states = norm["mean_obs"] + norm["std_obs"] * rng.standard_normal((B, 20))
    batched = c.get_actions(states)
    Fixed.n = 0
    single = np.stack([c.get_action(s) for s in states])
    assert np.array_equal(batched, single)
    mpc = O.OracleMPC(O.OracleDynamics(weights, norm), H, K)
    for b in range(B):
        assert np.array_equal(batched[b], mpc.get_action(states[b], draws[b]))
    small = Planner(10, 2, 1, 8, "relu", num_paths=32, horizon=2)
    small.set_dynamics([(np.zeros((12, 8), np.float32), np.zeros(8, np.float32)), (np.zeros((8, 10), np.float32), np.zeros(10, np.float32))])
    small.set_normalization(np.zeros(10), np.ones(10), np.zeros(2), np.ones(2), np.zeros(10), np.ones(10))
    with pytest.raises(RuntimeError, match="obs\\[17\\]"):
        small.plan(np.zeros(10))


def test_peer_exchange_matches_nccl_all_gather():
    """The NVLink peer-store record exchange (bcmpc_set_peer_exchange, one kernel after the rollout) against ncclAllGather +
    min-loc and against a single handle over all paths: bit-identical winners on every rank (2 GPUs, torchrun)."""
    import json
    import os
    import subprocess
    import sys

    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29547", os.path.join(root, "tools", "multi_gpu_check.py")], cwd=root, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1]
    assert json.loads(line)["identical_on_all_ranks"] is True
