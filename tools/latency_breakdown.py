"""Single-plan latency of the host-buffer C-ABI call (configs[0] shape) split into kernel time and everything else."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bc_mpc_b200 import Planner  # noqa: E402
from oracle import mpc_oracle as O  # noqa: E402

rng, norm, weights, state = O.synthetic_problem(0, (500, 500))
print("BCMPC_DEEP =", os.environ.get("BCMPC_DEEP"))
for K, H in ((1000, 15), (1000, 1), (4096, 15)):
    pl = Planner(20, 6, 2, 500, "relu", num_paths=K, horizon=H, precision="bf16")
    pl.set_dynamics(weights)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
    pl.enable_timing(True)
    for i in range(20):
        pl.plan(state, seed=i)
    n = 200
    t0 = time.perf_counter()
    for i in range(n):
        pl.plan(state, seed=100 + i)
    wall = (time.perf_counter() - t0) / n
    k_ms = pl.last_kernel_ms()
    print(f"K={K} H={H}: wall {wall * 1e6:.1f} us per plan, rollout+reduce kernels {k_ms * 1e3:.1f} us, rest {wall * 1e6 - k_ms * 1e3:.1f} us")
    pl.close()

# how much of "rest" is the Python wrapper: the same C call with preallocated arrays
import ctypes as C
from bc_mpc_b200 import _lib as L
pl = Planner(20, 6, 2, 500, "relu", num_paths=1000, horizon=15, precision="bf16")
pl.set_dynamics(weights)
pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
st = np.ascontiguousarray(state, np.float64)
out = np.zeros(1, L.RECORD_DTYPE)
ps, po = L.ptr(st), L.ptr(out)
f = pl._lib.bcmpc_plan
for i in range(20):
    f(pl._h, ps, 1, None, C.c_uint64(i), po)
n = 500
t0 = time.perf_counter()
for i in range(n):
    f(pl._h, ps, 1, None, C.c_uint64(100 + i), po)
raw = (time.perf_counter() - t0) / n
t0 = time.perf_counter()
for i in range(n):
    pl.plan(state, seed=100 + i)
wrapped = (time.perf_counter() - t0) / n
print(f"K=1000 H=15: raw C-ABI call {raw * 1e6:.1f} us, Planner.plan {wrapped * 1e6:.1f} us")
