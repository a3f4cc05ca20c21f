"""Drop-in for the reference cost_functions.py.  Inside the planner the cost is fused into the
rollout kernel's epilogue; these stand-alone entry points (used by the reference on REAL rollouts,
utils.py:128-130 path_cost) run the same arithmetic on the device through bcmpc_cheetah_cost."""
import numpy as np

from . import _lib
from .planner import Planner

_cost_planners = {}


def _planner():
    """One helper handle per GPU, on the calling thread's CURRENT device (rank r of a multi-GPU job stays on GPU r)."""
    dev = _lib.current_device()
    if dev not in _cost_planners:
        _cost_planners[dev] = Planner(num_paths=1, horizon=1, hidden_size=4, n_hidden=1, device=dev)
    return _cost_planners[dev]


def cheetah_cost_fn(state, action, next_state):
    """cost_functions.py:9-52: batched [K,D] -> float64 [K]; 1-D inputs -> python float."""
    state = np.asarray(state, np.float64)
    next_state = np.asarray(next_state, np.float64)
    if state.ndim > 1:
        return _planner().cheetah_cost(state, next_state)
    return float(_planner().cheetah_cost(state[None], next_state[None])[0])


def trajectory_cost_fn(cost_fn, states, actions, next_states):
    """cost_functions.py:59-63."""
    trajectory_cost = 0
    for i in range(len(actions)):
        trajectory_cost += cost_fn(states[i], actions[i], next_states[i])
    return trajectory_cost
