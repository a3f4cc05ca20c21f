"""CPU oracle for the bc-mpc random-shooting planner.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  Nothing under
``bc_mpc_b200/`` imports it: the product path is the CUDA library and fails
loudly when that library is missing.

What is restated and what it follows (citations into the reference tree):

* ``OracleDynamics.predict``      - ``dynamics.py:106-119`` (precision map:
  float64 normalise ``(x-mu)/(sigma+1e-10)`` -> float32 MLP
  ``dynamics.py:54-71`` -> float64 denormalise ``y*sigma+mu`` (no epsilon,
  ``dynamics.py:77-79``) -> float64 ``s + delta``).  Optional LayerNorm after
  the activation (``dynamics.py:68-69``).
* ``OraclePolicy.act``            - ``ppo_bc_policy.py:60,70-72,79-86,174-185``
  (obs filter + clip +-5, tanh MLP, mean; ``sample = mean + exp(logstd)*eps``
  per ``distributions.py:79-80``; ``mode = mean`` ``distributions.py:68-69``).
* ``OracleBC.predict``            - ``behavioral_cloning.py:21-26,44-47``.
* ``cheetah_cost_fn`` / ``trajectory_cost_fn`` - ``cost_functions.py:9-63``.
* ``OracleRewardDynamics.predict`` - ``dynamics.py:121-238`` (shared tanh layer, delta head, reward head).
* ``OracleMPC.get_action``        - ``controllers.py:57-88`` (uniform shooting),
  ``controllers.py:189-237`` (policy proposals) and the legacy
  ``MPCcontroller_BC`` recovered from ``controllers.pyc``.

Pinning status
--------------
* cost / trajectory cost / argmin / first-action gather / tensor layout: PINNED
  against the reference's own ``controllers.py`` and ``cost_functions.py``
  executed here (``tests/test_oracle_vs_reference.py`` when ``/root/reference``
  is mounted) and against golden vectors those files produced
  (``tests/golden/*.npz``, generator ``tests/golden/make_golden.py``).
* the MLP forward itself (``tf.layers.dense``, ``tf.nn.relu``, baselines
  ``RunningMeanStd``, tflearn ``fully_connected``) lives in TensorFlow 1.x /
  baselines / tflearn, which are un-vendored, unpinned third-party packages
  that are not installed in this image: **parity unpinned** for those
  three ops.  They are restated from their published semantics
  (``y = x @ kernel + bias`` with kernel ``[in,out]``; ``relu = max(x,0)``;
  ``rms.std = sqrt(max(sumsq/count - mean^2, 1e-2))``).
"""
from __future__ import annotations

import copy

import numpy as np

# ----------------------------------------------------------------------------
# cost functions (cost_functions.py:9-63)
# ----------------------------------------------------------------------------


def cheetah_cost_fn(state, action, next_state):
    """Restates cost_functions.py:9-52 (batched branch :10-30, scalar :32-52)."""
    state = np.asarray(state)
    next_state = np.asarray(next_state)
    if state.ndim > 1:
        scores = np.zeros((state.shape[0],))
        scores[state[:, 5] >= 0.2] += 10  # cost_functions.py:16-18
        scores[state[:, 6] >= 0] += 10  # :20-22
        scores[state[:, 7] >= 0] += 10  # :24-26
        scores -= (next_state[:, 17] - state[:, 17]) / 0.01  # :28
        return scores
    score = 0
    if state[5] >= 0.2:
        score += 10
    if state[6] >= 0:
        score += 10
    if state[7] >= 0:
        score += 10
    score -= (next_state[17] - state[17]) / 0.01
    return score


def trajectory_cost_fn(cost_fn, states, actions, next_states):
    """Restates cost_functions.py:59-63 (sum in step order, start value int 0)."""
    trajectory_cost = 0
    for i in range(len(actions)):
        trajectory_cost += cost_fn(states[i], actions[i], next_states[i])
    return trajectory_cost


# ----------------------------------------------------------------------------
# models
# ----------------------------------------------------------------------------

_ACT = {
    "relu": lambda x: np.maximum(x, np.float32(0)),
    "tanh": np.tanh,
    None: lambda x: x,
}


def _dense(x, w, b):
    # tf.layers.dense: x @ kernel + bias, kernel [in, out], all float32.
    return (x @ w + b).astype(np.float32, copy=False)


def _layer_norm(x, gamma, beta, eps=1e-12):
    # tf.contrib.layers.layer_norm: moments over the last axis, variance_epsilon
    # 1e-12, center + scale.  [from memory - not in tree; parity unpinned]
    mean = x.mean(axis=-1, keepdims=True, dtype=np.float32)
    var = ((x - mean) ** 2).mean(axis=-1, keepdims=True, dtype=np.float32)
    return ((x - mean) / np.sqrt(var + np.float32(eps)) * gamma + beta).astype(np.float32)


class OracleDynamics:
    """NNDynamicsModel.predict (dynamics.py:106-119) without TensorFlow.

    weights: list of (kernel[in,out] float32, bias[out] float32), hidden layers
             first, output layer last (names NNDynamicsModel/dense{,_1,_2}).
    normalization: the 10-tuple of utils.py:158 / dynamics.py:41, or a dict with
             mean_obs,std_obs,mean_action,std_action,mean_deltas,std_deltas.
    """

    def __init__(self, weights, normalization, activation="relu", output_activation=None,
                 layer_norm=None):
        self.weights = [(np.asarray(w, np.float32), np.asarray(b, np.float32)) for w, b in weights]
        if isinstance(normalization, dict):
            n = normalization
            self.mean_obs, self.std_obs = n["mean_obs"], n["std_obs"]
            self.mean_action, self.std_action = n["mean_action"], n["std_action"]
            self.mean_deltas, self.std_deltas = n["mean_deltas"], n["std_deltas"]
        else:
            (self.mean_obs, self.std_obs, self.mean_action, self.std_action, _mr, _sr, _mn, _sn,
             self.mean_deltas, self.std_deltas) = normalization
        for k in ("mean_obs", "std_obs", "mean_action", "std_action", "mean_deltas", "std_deltas"):
            setattr(self, k, np.asarray(getattr(self, k), np.float64))
        self.activation = activation
        self.output_activation = output_activation
        self.layer_norm = layer_norm  # list of (gamma, beta) per hidden layer or None

    def mlp(self, x):
        """dynamics.py:54-71, float32 throughout."""
        out = np.asarray(x, np.float32)
        n_hidden = len(self.weights) - 1
        for i in range(n_hidden):
            w, b = self.weights[i]
            out = _ACT[self.activation](_dense(out, w, b))
            if self.layer_norm is not None:
                g, be = self.layer_norm[i]
                out = _layer_norm(out, np.asarray(g, np.float32), np.asarray(be, np.float32))
        w, b = self.weights[-1]
        return _ACT[self.output_activation](_dense(out, w, b))

    def predict(self, unnormalized_state, unnormalized_action):
        s = np.asarray(unnormalized_state)
        a = np.asarray(unnormalized_action)
        ns = (s - self.mean_obs) / (self.std_obs + 1e-10)  # dynamics.py:109
        na = (a - self.mean_action) / (self.std_action + 1e-10)  # :110
        x = np.concatenate([ns.astype(np.float32), na.astype(np.float32)], axis=1)  # :26, feed cast
        y = self.mlp(x)  # :112
        delta = (y * self.std_deltas) + self.mean_deltas  # :113 / :77-79, float64
        return s + delta  # :116


class OracleRewardDynamics(OracleDynamics):
    """NNDynamicsRewardModel (dynamics.py:121-238) without TensorFlow: a shared tanh layer, a tanh + linear head that
    predicts the normalised state delta and a tanh + linear head that predicts the normalised reward (:162-193).

    weights: [(kernel, bias)] x 3 for share / delta-hidden / delta-out (names dense, dense_1, dense_2 in creation order)
    reward_weights: [(kernel, bias)] x 2 for reward-hidden / reward-out (dense_3, dense_4).
    normalization must carry mean_reward / std_reward (utils.py:158 order when it is the 10-tuple).
    layer_norm: [(gamma, beta)] for share and delta-hidden; reward_layer_norm: (gamma, beta) for reward-hidden.
    predict returns (next_state float64 [K,D], reward float64 [K,1]) like dynamics.py:221-238."""

    def __init__(self, weights, reward_weights, normalization, layer_norm=None, reward_layer_norm=None):
        super().__init__(weights, normalization, activation="tanh", layer_norm=layer_norm)
        self.reward_weights = [(np.asarray(w, np.float32), np.asarray(b, np.float32)) for w, b in reward_weights]
        if isinstance(normalization, dict):
            self.mean_reward, self.std_reward = normalization["mean_reward"], normalization["std_reward"]
        else:
            self.mean_reward, self.std_reward = normalization[4], normalization[5]
        self.mean_reward = np.float64(self.mean_reward)
        self.std_reward = np.float64(self.std_reward)
        self.reward_layer_norm = reward_layer_norm

    def heads(self, x):
        x = np.asarray(x, np.float32)
        (w0, b0), (w1, b1), (w2, b2) = self.weights
        (r1, c1), (r2, c2) = self.reward_weights
        share = np.tanh(_dense(x, w0, b0)).astype(np.float32)                       # dynamics.py:165 / :181
        if self.layer_norm is not None:
            share = _layer_norm(share, *[np.asarray(t, np.float32) for t in self.layer_norm[0]])
        d = np.tanh(_dense(share, w1, b1)).astype(np.float32)                       # :169 / :183
        if self.layer_norm is not None:
            d = _layer_norm(d, *[np.asarray(t, np.float32) for t in self.layer_norm[1]])
        delta = _dense(d, w2, b2)                                                   # :172 / :184
        r = np.tanh(_dense(share, r1, c1)).astype(np.float32)                       # :175 / :187
        if self.reward_layer_norm is not None:
            r = _layer_norm(r, *[np.asarray(t, np.float32) for t in self.reward_layer_norm])
        reward = _dense(r, r2, c2)                                                  # :178 / :188
        return delta, reward

    def predict(self, unnormalized_state, unnormalized_action):
        s = np.asarray(unnormalized_state)
        a = np.asarray(unnormalized_action)
        ns = (s - self.mean_obs) / (self.std_obs + 1e-10)                           # :224
        na = (a - self.mean_action) / (self.std_action + 1e-10)                     # :225
        delta, reward = self.heads(np.concatenate([ns.astype(np.float32), na.astype(np.float32)], axis=1))
        nxt = s + (delta * self.std_deltas + self.mean_deltas)                      # :233-234
        return nxt, reward * self.std_reward + self.mean_reward                    # :236


class OraclePolicy:
    """MlpPolicy 'pol' branch + act (ppo_bc_policy.py:56-60,66-88,174-185)."""

    def __init__(self, weights, logstd, rms_mean, rms_std, clip=5.0):
        self.weights = [(np.asarray(w, np.float32), np.asarray(b, np.float32)) for w, b in weights]
        self.logstd = np.asarray(logstd, np.float32).reshape(-1)
        self.rms_mean = np.asarray(rms_mean, np.float32)
        self.rms_std = np.asarray(rms_std, np.float32)
        self.clip = clip

    @staticmethod
    def rms_from_sums(runningsum, runningsumsq, count):
        """baselines RunningMeanStd [from memory - parity unpinned]: mean and std are
        float32 casts; std = sqrt(max(sumsq/count - mean^2, 1e-2))."""
        mean = (np.asarray(runningsum, np.float64) / count).astype(np.float32)
        msq = (np.asarray(runningsumsq, np.float64) / count).astype(np.float32)
        std = np.sqrt(np.maximum(msq - mean * mean, np.float32(1e-2))).astype(np.float32)
        return mean, std

    def mean(self, ob):
        ob = np.asarray(ob, np.float32)  # feed cast to the tf.float32 placeholder
        if ob.ndim == 1:
            ob = ob[None]  # ppo_bc_policy.py:176-178
        out = (ob - self.rms_mean) / self.rms_std
        if self.clip is not None:
            out = np.clip(out, np.float32(-self.clip), np.float32(self.clip))  # :60
        for w, b in self.weights[:-1]:
            out = np.tanh(_dense(out, w, b)).astype(np.float32)  # :70-71
        w, b = self.weights[-1]
        return _dense(out, w, b)  # :72

    def act(self, ob, stochastic=True, noise=None):
        """Returns (ac, vpred).  vpred is discarded by the planner (controllers.py:203);
        the oracle returns None for it.  ``noise`` is the injected N(0,1) draw that
        tf.random_normal would have produced (distributions.py:79-80)."""
        mean = self.mean(ob)
        if stochastic:
            if noise is None:
                noise = np.random.standard_normal(mean.shape)
            eps = np.asarray(noise, np.float32)
            std = np.exp(self.logstd).astype(np.float32)
            return (mean + (std * eps).astype(np.float32)).astype(np.float32), None
        return mean, None


class OracleBC:
    """BCnetwork.predict (behavioral_cloning.py:21-26,44-47): tanh,tanh,linear."""

    def __init__(self, weights):
        self.weights = [(np.asarray(w, np.float32), np.asarray(b, np.float32)) for w, b in weights]

    def predict(self, obs):
        out = np.asarray(obs, np.float32)
        for w, b in self.weights[:-1]:
            out = np.tanh(_dense(out, w, b)).astype(np.float32)
        w, b = self.weights[-1]
        return _dense(out, w, b)


# ----------------------------------------------------------------------------
# planner
# ----------------------------------------------------------------------------


class OracleMPC:
    """Random-shooting planner with injectable randomness.

    proposal:
      "uniform" (P0) controllers.py:57-88   actions = inj            (inj ~ U[low,high))
      "policy_gaussian" (P1) :203           actions = act(s, True)   (inj = N(0,1) eps)
      "policy_mix" (P2) :205-208            actions = (1-e)*act(s, False) + e*inj  (inj ~ U[low,high))
      "bc_uniform" (P3) controllers.pyc 120-163   actions = net(s) + inj  (inj ~ U[-1,1))
    ``inj`` has the reference's action-path layout [H, K, A] (controllers.py:53).
    """

    def __init__(self, dyn_model, horizon, num_simulated_paths, cost_fn=cheetah_cost_fn,
                 proposal="uniform", policy=None, explore=1.0, learned_reward=False, gamma=1.0):
        self.dyn_model = dyn_model
        self.horizon = horizon
        self.num_simulated_paths = num_simulated_paths
        self.cost_fn = cost_fn
        self.proposal = proposal
        self.policy = policy
        self.explore = explore
        # learned-reward planners: MPCcontrollerReward (controllers.py:121-158, sum of gamma**i * reward, argmax) and
        # MPCcontrollerPolicyNetReward (:319-363, plain sum, argmax).  ``costs`` are then MINUS the reward sums, so
        # that argmin(costs) == argmax(rewards) with the same first-index tie rule.
        self.learned_reward = learned_reward
        self.gamma = gamma

    def rollout(self, state, inj):
        K, H = self.num_simulated_paths, self.horizon
        inj = np.asarray(inj)
        assert inj.shape[:2] == (H, K), (inj.shape, H, K)
        states = np.tile(state, [K, 1])  # controllers.py:63
        states_all = [states]
        actions_all = []
        rewards_all = []
        for i in range(H):
            if self.proposal == "uniform":
                actions = inj[i, :, :]  # :70
            elif self.proposal == "policy_gaussian":
                actions, _ = self.policy.act(states, stochastic=True, noise=inj[i])  # :203
            elif self.proposal == "policy_mix":
                actions, _ = self.policy.act(states, stochastic=False)  # :205
                actions = (1 - self.explore) * actions + self.explore * inj[i, :, :]  # :208
            elif self.proposal == "bc_uniform":
                actions = self.policy.predict(states)  # controllers.pyc src 120-163
                actions += inj[i, :, :]  # in place: float32 array += float64 draw stays float32
            else:
                raise ValueError(self.proposal)
            if self.learned_reward:
                states, reward = self.dyn_model.predict(states, actions)  # :137 / :340
                rewards_all.append(reward * self.gamma ** i)  # :139 (gamma = 1 reproduces :345)
            else:
                states = self.dyn_model.predict(states, actions)  # :70 / :210
            states_all.append(states)
            actions_all.append(actions)
        states_all = np.asarray(states_all)  # :74
        actions_all = np.asarray(actions_all)
        if self.learned_reward:
            total = np.reshape(np.sum(np.asarray(rewards_all), axis=0), [-1])  # :151-153
            return states_all, actions_all, -np.asarray(total, np.float64)
        costs = trajectory_cost_fn(self.cost_fn, states_all[:-1], actions_all, states_all[1:])  # :80
        return states_all, actions_all, np.asarray(costs, np.float64)

    def get_action(self, state, inj, return_debug=False):
        states_all, actions_all, costs = self.rollout(state, inj)
        j = int(np.argmin(costs))  # :82 first index on ties, first NaN wins
        action = copy.copy(actions_all[:, j, :][0])  # :84-85
        if return_debug:
            return action, dict(states=states_all, actions=actions_all, costs=costs, argmin=j)
        return action


class OracleMCTS:
    """MCTScontrollerPolicyNetReward.get_action (controllers.py:397-457) with injectable randomness.

    N first-stage actions, each followed by R rollouts of H steps under the DETERMINISTIC policy (:429); the score of a
    first-stage action is reward_1 + mean over its R rollouts of the summed rewards (:445-450), argmax (:453).
    first_stage: "uniform" (random_first_stage_action, :409-411; inj = the N sampled actions [N,A]),
                 "policy_gaussian" (self_exp, :413-414; inj = N(0,1) noise [N,A]),
                 "policy_mean" (:415-416; no randomness, all N actions are equal).
    Returns the [1,A] first-stage action like the reference does (:455-457)."""

    def __init__(self, dyn_model, policy, horizon, num_first_stage_actions, random_path_per_action, first_stage):
        self.dyn_model, self.policy, self.horizon = dyn_model, policy, horizon
        self.N, self.R, self.first_stage = num_first_stage_actions, random_path_per_action, first_stage

    def get_action(self, state, inj=None, return_debug=False):
        state_init = np.expand_dims(state, axis=0)  # :402
        action_1s, reward_1s, states_all_actions = [], [], []
        for n in range(self.N):  # :405
            if self.first_stage == "uniform":
                action_1 = np.expand_dims(inj[n], axis=0)  # :410-411
            elif self.first_stage == "policy_gaussian":
                action_1, _ = self.policy.act(state_init, stochastic=True, noise=inj[n:n + 1])  # :414
            else:
                action_1, _ = self.policy.act(state_init, stochastic=False)  # :416
            state_1, reward_1 = self.dyn_model.predict(state_init, action_1)  # :418
            reward_1s.append(reward_1[0][0])
            action_1s.append(action_1)
            states_all_actions.append(np.tile(state_1, [self.R, 1]))  # :424-425
        states = np.asarray(states_all_actions).reshape((-1, state.shape[0]))  # :427-429
        rewards_all = []
        for _ in range(self.horizon):  # :431
            actions, _ = self.policy.act(states, stochastic=False)  # :433
            states, reward = self.dyn_model.predict(states, actions)  # :440
            rewards_all.append(reward)
        rewards_all = np.sum(np.asarray(rewards_all), axis=0).reshape((self.N, -1))  # :444-447
        total = np.asarray(reward_1s) + np.mean(rewards_all, axis=1)  # :449-451
        best = int(np.argmax(total))  # :454
        if return_debug:
            return action_1s[best], dict(total=np.asarray(total, np.float64), best=best,
                                         actions=np.concatenate(action_1s, axis=0))
        return action_1s[best]


# ----------------------------------------------------------------------------
# synthetic inputs - shared by tests, smoke and bench
# ----------------------------------------------------------------------------

# obs-filter-derived HalfCheetah statistics; index 18 has zero variance.
CHEETAH_MEAN_OBS = np.array([-0.10, 0.06, -0.05, 0.02, -0.01, -0.09, -0.15, -0.18, 2.18, -0.01,
                             0.01, 0.00, -0.01, 0.00, -0.01, -0.01, -0.01, 10.30, 0.0, 0.46])
CHEETAH_STD_OBS = np.array([0.05, 0.08, 0.21, 0.16, 0.22, 0.20, 0.22, 0.13, 1.16, 0.53, 1.17,
                            5.01, 3.68, 6.65, 4.59, 5.72, 2.69, 7.64, 0.0, 0.04])


def glorot_mlp(rng, sizes, bias_scale=0.0):
    """tf.layers.dense default init (glorot-uniform kernels, zero biases) [from memory];
    ``bias_scale`` > 0 draws N(0, bias_scale) biases so bias-add bugs are visible."""
    ws = []
    for fan_in, fan_out in zip(sizes[:-1], sizes[1:]):
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        w = rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(np.float32)
        b = (rng.standard_normal(fan_out) * bias_scale).astype(np.float32)
        ws.append((w, b))
    return ws


def synthetic_normalization(obs_dim=20, act_dim=6):
    mean_obs = np.resize(CHEETAH_MEAN_OBS, obs_dim).astype(np.float64)
    std_obs = np.resize(CHEETAH_STD_OBS, obs_dim).astype(np.float64)
    mean_deltas = np.zeros(obs_dim)
    if obs_dim > 17:
        mean_deltas[17] = 0.1
    return dict(mean_obs=mean_obs, std_obs=std_obs,
                mean_action=np.zeros(act_dim), std_action=np.full(act_dim, 0.58),
                mean_deltas=mean_deltas, std_deltas=0.1 * std_obs)


def synthetic_state(rng, norm):
    s = norm["mean_obs"] + norm["std_obs"] * rng.standard_normal(norm["mean_obs"].shape[0])
    return s


def synthetic_problem(seed, hidden=(500, 500), obs_dim=20, act_dim=6, bias_scale=0.1):
    rng = np.random.default_rng(seed)
    norm = synthetic_normalization(obs_dim, act_dim)
    weights = glorot_mlp(rng, (obs_dim + act_dim, *hidden, obs_dim), bias_scale)
    state = synthetic_state(rng, norm)
    return rng, norm, weights, state


def synthetic_reward_problem(seed, hidden=500, obs_dim=20, act_dim=6, bias_scale=0.1):
    """NNDynamicsRewardModel inputs: shared + delta-head weights (3 layers), reward-head weights (2 layers), and a
    normalisation dict that also carries mean_reward / std_reward."""
    rng = np.random.default_rng(seed)
    norm = synthetic_normalization(obs_dim, act_dim)
    norm["mean_reward"], norm["std_reward"] = 0.3, 1.7
    weights = glorot_mlp(rng, (obs_dim + act_dim, hidden, hidden, obs_dim), bias_scale)
    reward_weights = glorot_mlp(rng, (hidden, hidden, 1), bias_scale)
    state = synthetic_state(rng, norm)
    return rng, norm, weights, reward_weights, state


def synthetic_policy(rng, obs_dim=20, act_dim=6, hidden=(128, 128), norm=None):
    ws = glorot_mlp(rng, (obs_dim, *hidden, act_dim), 0.05)
    logstd = np.full((1, act_dim), -1.7, np.float32) + 0.1 * rng.standard_normal((1, act_dim)).astype(np.float32)
    norm = norm or synthetic_normalization(obs_dim, act_dim)
    rms_mean = norm["mean_obs"].astype(np.float32)
    rms_std = np.sqrt(np.maximum(norm["std_obs"] ** 2, 1e-2)).astype(np.float32)
    return ws, logstd, rms_mean, rms_std
