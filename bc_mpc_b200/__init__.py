"""bc_mpc_b200: B200-native drop-in for the planning hot path of Bonnieccc/bc-mpc.

Python surface mirrors the reference modules for this path:
  controllers     Controller, RandomController, MPCcontroller, MPCcontrollerPolicyNet, MPCcontroller_BC
  dynamics        NNDynamicsModel (predict + weight hand-off; training stays in the host framework)
  cost_functions  cheetah_cost_fn, trajectory_cost_fn
  policy          MlpPolicy (act), BCnetwork (predict)
underneath: planner.Planner -> ctypes -> libbcmpc.so (include/bcmpc.h) -> sm_100a kernels.
There is no CPU fallback anywhere in this package.
"""
from .planner import Planner  # noqa: F401

__all__ = ["Planner"]
