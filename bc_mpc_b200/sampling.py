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


def path_cost(cost_fn, path):
    """utils.py:128-130."""
    return trajectory_cost_fn(cost_fn, path["observations"], path["actions"], path["next_observations"])
