"""One stand-alone cost-stage launch (2^21 rows of [n,20] float64 states / next states) for an ncu capture."""
import numpy as np
from bc_mpc_b200 import Planner

pl = Planner(20, 6, 2, 500, "relu", num_paths=1024, horizon=4, max_states=1, precision="bf16")
pl.enable_timing(True)
n = 1 << 21
S0 = np.random.default_rng(4).standard_normal((n, 20))
for _ in range(2):
    pl.cheetah_cost(S0, S0 + 0.01)
ms = pl.last_kernel_ms()
print("kernel_ms", ms, "array GB/s", (2 * S0.nbytes + 8 * n) / ms / 1e6)
