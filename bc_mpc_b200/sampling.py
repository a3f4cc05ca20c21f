"""Callers of the planner: the reference's env samplers (utils.py:91-126 `sample`, :128-130 `path_cost`) and a
vectorised variant that feeds many concurrent envs to one batched plan per time step (SURVEY.md section 8f rank 3 -
the caller that can actually present 256 root states at once, BASELINE.json configs[2])."""
import numpy as np

from .cost_functions import trajectory_cost_fn


def _reset(env):
    # utils.py:108-111: HalfCheetah-v1 resets through reset_model(); anything else through reset()
    return env.reset_model() if hasattr(env, "reset_model") else env.reset()


def sample(env, controller, num_paths=10, horizon=1000, render=False, verbose=False):
    """utils.py:91-126, one env: num_paths rollouts of `horizon` steps, one get_action per step."""
    paths = []
    for _ in range(num_paths):
        st = _reset(env)
        path = {"observations": [], "actions": [], "rewards": [], "next_observations": []}
        for _t in range(horizon):
            at = controller.get_action(st)
            st_next, r_t, _, _ = env.step(at)
            path["observations"].append(st)
            path["actions"].append(at)
            path["rewards"].append(r_t)
            path["next_observations"].append(st_next)
            st = st_next
        paths.append(path)
    return paths


def sample_vectorized(envs, controller, horizon=1000):
    """len(envs) rollouts in lock step: per time step ONE controller.get_actions(states[B,D]) call plans for every env
    (B x K paths in one launch), then every env steps.  Returns the same list of path dicts as `sample` with
    num_paths = len(envs)."""
    states = [_reset(e) for e in envs]
    paths = [{"observations": [], "actions": [], "rewards": [], "next_observations": []} for _ in envs]
    for _t in range(horizon):
        actions = controller.get_actions(np.asarray(states, np.float64))
        for i, env in enumerate(envs):
            st_next, r_t, _, _ = env.step(actions[i])
            paths[i]["observations"].append(states[i])
            paths[i]["actions"].append(actions[i])
            paths[i]["rewards"].append(r_t)
            paths[i]["next_observations"].append(st_next)
            states[i] = st_next
    return paths


def traj_segment_generator(pi, mpc_controller, mpc_ppo_controller, bc_data_buffer, env, mpc, ppo_mpc, horizon):
    """utils.py:163-242: one PPO segment of `horizon` env steps.  Every step evaluates the policy (action + value) and,
    when `mpc`, asks the planner for the action that is actually executed (mpc_ppo_controller if ppo_mpc)."""
    return traj_segments_vectorized(pi, mpc_ppo_controller if ppo_mpc else mpc_controller, [env], mpc, horizon,
                                    batched=False)[0]


def traj_segments_vectorized(pi, controller, envs, mpc, horizon, batched=True):
    """len(envs) PPO segments in lock step; with `batched` the planner is called once per time step for all envs
    (controller.get_actions), otherwise once per env (controller.get_action), exactly like utils.py:200-202."""
    n = len(envs)
    ac0 = [e.action_space.sample() for e in envs]           # :165 "not used, just so we have the datatype"
    obs0 = [e.reset() for e in envs]
    seg = [{"ob": np.array([obs0[i] for _ in range(horizon)]), "nxt_ob": np.array([obs0[i] for _ in range(horizon)]),
            "rew": np.zeros(horizon, "float32"), "vpred": np.zeros(horizon, "float32"), "new": np.zeros(horizon, "int32"),
            "ac": np.array([ac0[i] for _ in range(horizon)])} for i in range(n)]
    for sg in seg:
        sg["prevac"], sg["mpcac"] = sg["ac"].copy(), sg["ac"].copy()
    ob, ac, new = list(obs0), list(ac0), [True] * n
    ret = [0.0] * n
    vpred = [0.0] * n
    for t in range(horizon):
        prevac = list(ac)
        for i in range(n):
            a_i, v_i = pi.act(ob[i], stochastic=True)                             # :196
            if v_i is None:
                raise RuntimeError("traj_segment_generator needs pi.act to return vpred (utils.py:196,238): give the policy its "
                                   "value head (MlpPolicy.set_value_weights, or load_vars with the pi/vf/* variables)")
            # the reference's act returns [1,A] / [1] for a 1-D observation and stores them by broadcasting (:208-211)
            ac[i], vpred[i] = np.asarray(a_i).reshape(-1), float(np.asarray(v_i).reshape(-1)[0])
        if mpc:
            if batched:
                mpc_ac = list(controller.get_actions(np.asarray(ob, np.float64)))
            else:
                mpc_ac = [controller.get_action(ob[i]) for i in range(n)]          # :200-202
        else:
            mpc_ac = [np.array(a, copy=True) for a in ac]                          # :204
        for i, env in enumerate(envs):
            sg = seg[i]
            sg["ob"][t], sg["vpred"][t], sg["new"][t] = ob[i], vpred[i], new[i]
            sg["ac"][t], sg["prevac"][t], sg["mpcac"][t] = ac[i], prevac[i], mpc_ac[i]
            ob[i], rew, _done, _ = env.step(mpc_ac[i])                             # :213 (episodes are not cut at `done`)
            new[i] = False
            sg["nxt_ob"][t], sg["rew"][t] = ob[i], rew
            ret[i] += rew
    for i, sg in enumerate(seg):
        sg["nextvpred"] = vpred[i] * (1 - new[i])                               # :238
        sg["ep_rets"], sg["ep_lens"] = [ret[i]], [horizon]                         # :226-227
    return seg


def path_cost(cost_fn, path):
    """utils.py:128-130."""
    return trajectory_cost_fn(cost_fn, path["observations"], path["actions"], path["next_observations"])
