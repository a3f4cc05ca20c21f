"""Drop-in for the reference controllers.py planners, backed by the fused GPU rollout.

Same constructors and get_action / sample_random_actions signatures as controllers.py:26-88
(MPCcontroller), :160-237 (MPCcontrollerPolicyNet) and the legacy MPCcontroller_BC recovered from
controllers.pyc.  Differences that follow from fusing the loop:
  * randomness is drawn on the device (Philox, seeded from np.random so np.random.seed still makes runs
    reproducible) unless sample_random_actions is overridden, in which case the returned [H,K,A] tensor
    is injected exactly like controllers.py:60 does;
  * cost_fn must be cheetah_cost_fn (the cost is fused into the kernel epilogue).
"""
import copy

import numpy as np

from . import cost_functions
from .planner import Planner


class Controller:
    def __init__(self):
        pass

    def get_action(self, state):
        pass


class RandomController(Controller):
    def __init__(self, env):
        self.env = env

    def get_action(self, state):
        return self.env.action_space.sample()  # controllers.py:24


def _check_cost(cost_fn):
    if cost_fn is None or getattr(cost_fn, "__name__", "") != "cheetah_cost_fn":
        raise ValueError("the fused planner implements cheetah_cost_fn only (cost_functions.py:9-30); "
                         f"got {cost_fn!r}")


class _GpuShootingController(Controller):
    proposal = "uniform"
    first_proposal = None       # tree-search planner only: proposal mode of step 0

    learned_reward = False      # True: maximise the model's predicted reward instead of minimising cheetah_cost_fn
    gamma = 1.0

    def _init_common(self, env, dyn_model, horizon, cost_fn, num_simulated_paths, precision, device):
        if not self.learned_reward:
            _check_cost(cost_fn)
        self.env, self.dyn_model, self.horizon, self.cost_fn = env, dyn_model, horizon, cost_fn
        self.num_simulated_paths = num_simulated_paths
        self.precision, self.device = precision, device
        self._planner = None
        self._dyn_version = None
        self._pol_version = None
        self.last_record = None

    def sample_random_actions(self):
        """controllers.py:43-55 / :181-187: float64 [H,K,A], U[low,high).  Kept for API parity; the
        default planning path samples on the device instead of calling this."""
        a = self.env.action_space
        return np.random.uniform(low=a.low, high=a.high, size=[self.horizon, self.num_simulated_paths, len(a.high)])

    def _sampling_overridden(self):
        return ("sample_random_actions" in self.__dict__ or
                type(self).sample_random_actions is not _GpuShootingController.sample_random_actions)

    def _policy(self):
        return None

    def _explore(self):
        return 1.0

    def _injects(self):
        """True when this controller's get_action would inject sample_random_actions() instead of drawing on the device."""
        return self._sampling_overridden() and self.proposal in ("uniform", "policy_mix")

    def _ensure_planner(self, n_states=1):
        dm = self.dyn_model
        if self._planner is not None and n_states > self._planner.max_states:
            self._planner.close()        # grow: one handle sized for the largest batch of root states seen so far
            self._planner = None
        if self._planner is None:
            self._planner = Planner(dm.obs_dim, dm.act_dim, dm.n_layers, dm.size, dm.activation,
                                    num_paths=self.num_simulated_paths, horizon=self._planner_horizon(),
                                    max_states=1 if n_states == 1 else 1 << (n_states - 1).bit_length(),
                                    precision=self.precision, proposal=self.proposal, explore=self._explore(),
                                    device=self.device, first_proposal=self.first_proposal)
            self._planner.set_action_bounds(self.env.action_space.low, self.env.action_space.high)
            self._dyn_version = self._pol_version = None
        if self._dyn_version != dm.version:     # weights change after every dyn_model.fit (train_mpc_ppo.py:279)
            dm.upload(self._planner)
            self._planner.set_normalization(*dm.normalization6())
            if self.learned_reward:
                dm.upload_reward_head(self._planner, self.gamma)
            self._dyn_version = dm.version
        pol = self._policy()
        if pol is not None and self._pol_version != pol.version:
            pol.upload(self._planner)
            self._pol_version = pol.version
        return self._planner

    def _planner_horizon(self):
        return self.horizon

    def _plan(self, state, inj):
        p = self._ensure_planner()
        seed = int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 31)
        rec = p.plan(np.asarray(state, np.float64), inj=inj, seed=seed)[0]
        self.last_record = rec
        return copy.copy(rec["action"][: self.dyn_model.act_dim])   # float64 [A] (controllers.py:85)

    def get_actions(self, states):
        """Batched get_action: states [B,D] -> actions [B,A], all B plans in ONE launch (B x K paths in flight).  Not
        in the reference, whose samplers call get_action once per env step (utils.py:115,200); this is what a
        vectorised sampler over many envs calls (sampling.sample_vectorized).  Randomness: device Philox, counter
        (path, root state, step), so every root state gets its own K action sequences."""
        states = np.ascontiguousarray(states, np.float64)
        if states.ndim != 2:
            raise ValueError(f"states must be [B, obs_dim], got {states.shape}")
        p = self._ensure_planner(states.shape[0])
        seed = int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 31)
        # an overridden sample_random_actions is honoured exactly like in get_action: one [H,K,A] draw per root state
        inj = np.stack([self.sample_random_actions() for _ in range(states.shape[0])]) if self._injects() else None
        recs = p.plan(states, inj=inj, seed=seed)
        self.last_record = recs
        return recs["action"][:, : self.dyn_model.act_dim].copy()


class MPCcontroller(_GpuShootingController):
    """Random-shooting MPC (controllers.py:26-88)."""

    def __init__(self, env, dyn_model, horizon=5, cost_fn=None, num_simulated_paths=10, gamma=1.,
                 precision="fp32", device=0):
        self._init_common(env, dyn_model, horizon, cost_fn, num_simulated_paths, precision, device)
        self.gamma = gamma

    def get_action(self, state):
        inj = self.sample_random_actions() if self._sampling_overridden() else None
        return self._plan(state, inj)


class MPCcontrollerPolicyNet(_GpuShootingController):
    """Policy-proposal shooting (controllers.py:160-237)."""

    def __init__(self, env, dyn_model, policy_net, explore=1., self_exp=True, horizon=5, cost_fn=None,
                 num_simulated_paths=10, precision="fp32", device=0):
        self.policy_net, self.self_exp, self.explore = policy_net, self_exp, explore
        self.proposal = "policy_gaussian" if self_exp else "policy_mix"
        self._init_common(env, dyn_model, horizon, cost_fn, num_simulated_paths, precision, device)

    def _policy(self):
        return self.policy_net

    def _explore(self):
        return self.explore

    def get_action(self, state):
        # controllers.py:192 always draws `exploration`, but uses it only when not self_exp (:208)
        inj = self.sample_random_actions() if (not self.self_exp and self._sampling_overridden()) else None
        return self._plan(state, inj)


class MPCcontroller_BC(_GpuShootingController):
    """Legacy BC-guided shooting (controllers.pyc src 90-163): a = bc_network.predict(s) + U[-1,1)."""
    proposal = "bc_uniform"

    def __init__(self, env, dyn_model, bc_network, horizon=5, cost_fn=None, num_simulated_paths=10,
                 precision="fp32", device=0):
        self.bc_network = bc_network
        self._init_common(env, dyn_model, horizon, cost_fn, num_simulated_paths, precision, device)

    def _policy(self):
        return self.bc_network

    def get_action(self, state):
        return self._plan(state, None)


class MPCcontrollerReward(_GpuShootingController):
    """Random shooting on the learned reward (controllers.py:90-158): argmax over paths of sum_i gamma**i * reward_i
    predicted by an NNDynamicsRewardModel.  cost_fn is accepted and ignored, like in the reference."""
    learned_reward = True

    def __init__(self, env, dyn_model, horizon=5, cost_fn=None, num_simulated_paths=10, gamma=1., precision="fp32", device=0):
        self.gamma = gamma
        self._init_common(env, dyn_model, horizon, cost_fn, num_simulated_paths, precision, device)

    def get_action(self, state):
        inj = self.sample_random_actions() if self._sampling_overridden() else None
        return self._plan(state, inj)


class MPCcontrollerPolicyNetReward(MPCcontrollerPolicyNet):
    """Policy-proposal shooting on the learned reward (controllers.py:289-363): plain sum of predicted rewards (the
    reference stores gamma but does not apply it, :345-352), argmax."""
    learned_reward = True

    def __init__(self, env, dyn_model, policy_net, explore=1., self_exp=True, horizon=5, cost_fn=None,
                 num_simulated_paths=10, gamma=1., precision="fp32", device=0):
        super().__init__(env, dyn_model, policy_net, explore=explore, self_exp=self_exp, horizon=horizon, cost_fn=cost_fn,
                         num_simulated_paths=num_simulated_paths, precision=precision, device=device)
        self.gamma = 1.0      # not applied by the reference
        self.stored_gamma = gamma


class MCTScontrollerPolicyNetReward(_GpuShootingController):
    """Two-stage search on the learned reward (controllers.py:365-457): num_first_stage_actions candidate first actions
    (uniform, stochastic policy or deterministic policy), each continued for `horizon` steps under the deterministic
    policy; score = reward_1 + mean over the rollouts of the summed rewards, argmax; returns the [1,A] action.

    The reference runs random_path_per_action copies of every continuation, but act(stochastic=False) (:433) makes
    the copies identical, so their mean is the single value; one fused launch of N paths x (horizon + 1) steps with
    the step-0 proposal switched computes the same scores.  random_path_per_action is kept for the signature."""
    learned_reward = True
    proposal = "policy_mean"

    def __init__(self, env, dyn_model, policy_net, explore=1., self_exp=True, horizon=5, cost_fn=None,
                 num_first_stage_actions=10, random_path_per_action=10, random_first_stage_action=False,
                 precision="fp32", device=0):
        self.policy_net, self.self_exp, self.explore = policy_net, self_exp, explore
        self.num_first_stage_actions, self.random_path_per_action = num_first_stage_actions, random_path_per_action
        self.random_first_stage_action = random_first_stage_action
        self.first_proposal = ("uniform" if random_first_stage_action else "policy_gaussian" if self_exp else "policy_mean")
        self._init_common(env, dyn_model, horizon, cost_fn, num_first_stage_actions, precision, device)

    def _planner_horizon(self):
        return self.horizon + 1          # the first-stage step plus `horizon` continuation steps (:431)

    def _policy(self):
        return self.policy_net

    def sample_first_stage(self):
        """Override to inject the first-stage randomness: [N,A] actions (random_first_stage_action) or N(0,1) noise
        (self_exp).  None = drawn on the device."""
        return None

    def get_action(self, state):
        z = self.sample_first_stage()
        inj = None
        if z is not None:
            inj = np.zeros((self.horizon + 1, self.num_first_stage_actions, self.dyn_model.act_dim))
            inj[0] = z
        return self._plan(state, inj)[None, :]        # the reference returns action_1 with its batch axis (:455)
