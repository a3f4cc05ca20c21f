"""One stand-alone sampling-stage launch (K = 131072, H = 20, 4 root states) for an ncu capture."""
from bc_mpc_b200 import Planner
pl = Planner(20, 6, 2, 500, "relu", num_paths=1 << 17, horizon=20, max_states=4, precision="bf16")
pl.enable_timing(True)
for i in range(2):
    inj = pl.sample_injection(7 + i, n_states=4)
print("kernel_ms", pl.last_kernel_ms(), "GB/s", inj.nbytes / pl.last_kernel_ms() / 1e6)
