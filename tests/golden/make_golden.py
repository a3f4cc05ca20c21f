"""Generates tests/golden/*.npz by running the REFERENCE's own controllers.py and
cost_functions.py (imported verbatim from /root/reference, numpy-only modules) around
the oracle's restated predict / policy.  Run here (the reference does not travel to the
GPU box):   python tests/golden/make_golden.py

Files:
  cost_kat.npz        known-answer vectors of cheetah_cost_fn / trajectory_cost_fn
  mpc_uniform.npz     MPCcontroller.get_action (controllers.py:57-88), injected actions
  mpc_policy.npz      MPCcontrollerPolicyNet.get_action (controllers.py:189-237),
                      self_exp True (injected N(0,1)) and False (explore mix)
  vars_ppo_mpc.npz    the trained dynamics + policy weights of
                      saved_data/ppo-mpc/vars.pkl (data fixture) and a reference-run
                      plan on them
  mpc_reward.npz      the learned-reward and tree-search planners: MPCcontrollerReward
                      (controllers.py:90-158), MPCcontrollerPolicyNetReward (:289-363),
                      MCTScontrollerPolicyNetReward (:365-457) run by the reference on an
                      oracle NNDynamicsRewardModel (with its three LayerNorms in one case)
                      (python tests/golden/make_golden.py reward  regenerates only this file)
"""
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, "/root/reference")

import controllers as ref_controllers  # noqa: E402  (reference, verbatim)
import cost_functions as ref_cost  # noqa: E402  (reference, verbatim)

from oracle import mpc_oracle as O  # noqa: E402


class _Space:
    def __init__(self, low, high):
        self.low, self.high, self.shape = np.asarray(low, np.float64), np.asarray(high, np.float64), (len(low),)

    def sample(self):
        return np.random.uniform(self.low, self.high)


class FakeEnv:
    def __init__(self, obs_dim=20, act_dim=6):
        self.action_space = _Space([-1.0] * act_dim, [1.0] * act_dim)
        self.observation_space = _Space([-np.inf] * obs_dim, [np.inf] * obs_dim)


class _InjectedPolicy:
    """Adapts OraclePolicy to policy_net.act(ob, stochastic) with a queue of injected eps."""

    def __init__(self, pol, eps):
        self.pol, self.eps, self.i = pol, eps, 0

    def act(self, ob, stochastic=True):
        if stochastic:
            out = self.pol.act(ob, True, noise=self.eps[self.i])
            self.i += 1
            return out
        return self.pol.act(ob, False)


def ref_uniform(dyn, state, actions, H, K):
    c = ref_controllers.MPCcontroller(FakeEnv(), dyn, horizon=H, cost_fn=ref_cost.cheetah_cost_fn,
                                      num_simulated_paths=K)
    c.sample_random_actions = lambda: actions
    return c.get_action(state)


def ref_rollout_costs(dyn, state, actions):
    """controllers.py:63-80 re-executed with the reference cost functions to expose costs."""
    H, K = actions.shape[:2]
    states = np.tile(state, [K, 1])
    allst = [states]
    for i in range(H):
        states = dyn.predict(states, actions[i])
        allst.append(states)
    allst = np.asarray(allst)
    costs = ref_cost.trajectory_cost_fn(ref_cost.cheetah_cost_fn, allst[:-1], actions, allst[1:])
    return allst, costs


def reward_and_tree():
    out = {}
    seed, S = 11, 64
    rng, norm, weights, rweights, state = O.synthetic_reward_problem(seed, S)
    lns = [(1.0 + 0.1 * rng.standard_normal(S).astype(np.float32), 0.1 * rng.standard_normal(S).astype(np.float32)) for _ in range(3)]
    pw, logstd, rms_mean, rms_std = O.synthetic_policy(rng, hidden=(32, 32), norm=norm)
    pol = O.OraclePolicy(pw, logstd, rms_mean, rms_std)
    H, K, N, R = 5, 120, 40, 3
    actions = rng.uniform(-1, 1, (H, K, 6)).astype(np.float32).astype(np.float64)
    eps = rng.standard_normal((H, K, 6)).astype(np.float32)
    first_uni = rng.uniform(-1, 1, (N, 6)).astype(np.float32).astype(np.float64)
    first_eps = rng.standard_normal((N, 6)).astype(np.float32)
    out.update(seed=seed, hidden=S, state=state, actions=actions, eps=eps, first_uni=first_uni, first_eps=first_eps,
               ln_gamma=np.stack([g for g, _ in lns]), ln_beta=np.stack([b for _, b in lns]), N=N, R=R)
    for tag, dyn in (("plain", O.OracleRewardDynamics(weights, rweights, norm)),
                     ("ln", O.OracleRewardDynamics(weights, rweights, norm, layer_norm=lns[:2], reward_layer_norm=lns[2]))):
        for gamma in (1.0, 0.9):
            c = ref_controllers.MPCcontrollerReward(FakeEnv(), dyn, horizon=H, num_simulated_paths=K, gamma=gamma)
            c.sample_random_actions = lambda: actions
            out[f"{tag}_uniform_g{gamma}_first_action"] = c.get_action(state)
        # the per-path objective of controllers.py:131-153 re-executed for gamma = 0.9
        st, tot = np.tile(state, [K, 1]), 0.0
        for i in range(H):
            st, r = dyn.predict(st, actions[i])
            tot = tot + r * 0.9 ** i
        out[f"{tag}_uniform_g0.9_rewards"] = np.reshape(tot, [-1])
        c = ref_controllers.MPCcontrollerPolicyNetReward(FakeEnv(), dyn, _InjectedPolicy(pol, eps), explore=0.4, self_exp=False,
                                                         horizon=H, num_simulated_paths=K, gamma=0.5)
        c.sample_random_actions = lambda: actions
        out[f"{tag}_mix_first_action"] = c.get_action(state)
        c = ref_controllers.MPCcontrollerPolicyNetReward(FakeEnv(), dyn, _InjectedPolicy(pol, eps), explore=0.4, self_exp=True,
                                                         horizon=H, num_simulated_paths=K)
        out[f"{tag}_selfexp_first_action"] = c.get_action(state)
        # tree search: first stage uniform / stochastic policy
        env = FakeEnv()
        draws = iter(first_uni)
        env.action_space.sample = lambda: next(draws)
        c = ref_controllers.MCTScontrollerPolicyNetReward(env, dyn, _InjectedPolicy(pol, None), horizon=H, num_first_stage_actions=N,
                                                          random_path_per_action=R, random_first_stage_action=True)
        out[f"{tag}_tree_uniform_action"] = c.get_action(state)
        c = ref_controllers.MCTScontrollerPolicyNetReward(FakeEnv(), dyn, _InjectedPolicy(pol, first_eps[:, None, :]), self_exp=True, horizon=H,
                                                          num_first_stage_actions=N, random_path_per_action=R)
        out[f"{tag}_tree_selfexp_action"] = c.get_action(state)
    np.savez_compressed(os.path.join(HERE, "mpc_reward.npz"), **out)
    print("mpc_reward.npz", os.path.getsize(os.path.join(HERE, "mpc_reward.npz")))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "reward":
        return reward_and_tree()
    # ---- cost KATs -------------------------------------------------
    rng = np.random.default_rng(1234)
    S = rng.standard_normal((64, 20)) * 0.3
    S[0] = 0.0
    S[1, 5] = 0.2
    S[2, 5] = 0.19999
    S[3, 6] = -0.0
    S[4, 7] = 1e-300
    N = S + rng.standard_normal((64, 20)) * 0.05
    N[0] = 0.0
    A = rng.uniform(-1, 1, (64, 6))
    batched = ref_cost.cheetah_cost_fn(S, A, N)
    scalar = np.array([ref_cost.cheetah_cost_fn(S[i], A[i], N[i]) for i in range(64)], np.float64)
    St = rng.standard_normal((4, 8, 20)) * 0.3
    Nt = St + rng.standard_normal((4, 8, 20)) * 0.05
    At = rng.uniform(-1, 1, (4, 8, 6))
    traj = ref_cost.trajectory_cost_fn(ref_cost.cheetah_cost_fn, St, At, Nt)
    np.savez(os.path.join(HERE, "cost_kat.npz"), S=S, A=A, N=N, batched=batched, scalar=scalar,
             St=St, At=At, Nt=Nt, traj=traj)

    # ---- uniform shooting ------------------------------------------------------------------
    out = {}
    for tag, (seed, hidden, K, H) in {"a": (0, (64, 64), 96, 5), "b": (1, (500, 500), 200, 4),
                                      "c": (2, (40,), 33, 3)}.items():
        rng, norm, weights, state = O.synthetic_problem(seed, hidden)
        dyn = O.OracleDynamics(weights, norm)
        actions = rng.uniform(-1, 1, (H, K, 6)).astype(np.float32).astype(np.float64)
        if tag == "a":
            actions[:, 7] = actions[:, 3]  # tie: identical sequences -> lowest index wins
        act = ref_uniform(dyn, state, actions, H, K)
        allst, costs = ref_rollout_costs(dyn, state, actions)
        out.update({f"{tag}_seed": seed, f"{tag}_hidden": np.array(hidden), f"{tag}_state": state,
                    f"{tag}_actions": actions, f"{tag}_first_action": act, f"{tag}_costs": costs,
                    f"{tag}_states": allst, f"{tag}_argmin": int(np.argmin(costs))})
    np.savez_compressed(os.path.join(HERE, "mpc_uniform.npz"), **out)

    # ---- policy proposals ------------------------------------------------------------------
    out = {}
    seed, hidden, K, H = 3, (64, 64), 80, 4
    rng, norm, weights, state = O.synthetic_problem(seed, hidden)
    dyn = O.OracleDynamics(weights, norm)
    pw, logstd, rms_mean, rms_std = O.synthetic_policy(rng, hidden=(32, 32), norm=norm)
    pol = O.OraclePolicy(pw, logstd, rms_mean, rms_std)
    eps = rng.standard_normal((H, K, 6)).astype(np.float32)
    uni = rng.uniform(-1, 1, (H, K, 6)).astype(np.float32).astype(np.float64)
    c = ref_controllers.MPCcontrollerPolicyNet(FakeEnv(), dyn, _InjectedPolicy(pol, eps), explore=0.5,
                                               self_exp=True, horizon=H,
                                               cost_fn=ref_cost.cheetah_cost_fn, num_simulated_paths=K)
    out["selfexp_first_action"] = c.get_action(state)
    c = ref_controllers.MPCcontrollerPolicyNet(FakeEnv(), dyn, _InjectedPolicy(pol, eps), explore=0.5,
                                               self_exp=False, horizon=H,
                                               cost_fn=ref_cost.cheetah_cost_fn, num_simulated_paths=K)
    c.sample_random_actions = lambda: uni
    out["mix_first_action"] = c.get_action(state)
    out.update(seed=seed, hidden=np.array(hidden), state=state, eps=eps, uni=uni, explore=0.5)
    np.savez_compressed(os.path.join(HERE, "mpc_policy.npz"), **out)

    # ---- trained weights fixture (saved_data/ppo-mpc/vars.pkl) ------------------------------
    d = pickle.load(open("/root/reference/saved_data/ppo-mpc/vars.pkl", "rb"))
    fx = {}
    for i, nm in enumerate(["dense", "dense_1", "dense_2"]):
        fx[f"dyn_w{i}"] = d[f"NNDynamicsModel/{nm}/kernel:0"]
        fx[f"dyn_b{i}"] = d[f"NNDynamicsModel/{nm}/bias:0"]
    for i, nm in enumerate(["fc1", "fc2", "final"]):
        fx[f"pol_w{i}"] = d[f"pi/pi/pol/{nm}/kernel:0"]
        fx[f"pol_b{i}"] = d[f"pi/pi/pol/{nm}/bias:0"]
    fx["pol_logstd"] = d["pi/pi/pol/logstd:0"]
    fx["ob_sum"] = d["pi/pi/obfilter/runningsum:0"]
    fx["ob_sumsq"] = d["pi/pi/obfilter/runningsumsq:0"]
    fx["ob_count"] = d["pi/pi/obfilter/count:0"]
    # a reference-run plan on the trained dynamics net
    rng = np.random.default_rng(7)
    norm = O.synthetic_normalization()
    dyn = O.OracleDynamics([(fx[f"dyn_w{i}"], fx[f"dyn_b{i}"]) for i in range(3)], norm)
    state = O.synthetic_state(rng, norm)
    H, K = 6, 128
    actions = rng.uniform(-1, 1, (H, K, 6)).astype(np.float32).astype(np.float64)
    fx["plan_state"] = state
    fx["plan_actions"] = actions
    fx["plan_first_action"] = ref_uniform(dyn, state, actions, H, K)
    allst, costs = ref_rollout_costs(dyn, state, actions)
    fx["plan_costs"] = costs
    fx["plan_states"] = allst
    np.savez_compressed(os.path.join(HERE, "vars_ppo_mpc.npz"), **fx)
    reward_and_tree()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
