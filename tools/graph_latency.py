"""Single-plan latency (configs[0]: K=1000 x H=15, 2x500 net): Planner.plan (host buffers, eager launches) against one CUDA graph
[pinned states -> device, rollout + min-loc, records -> pinned host] replayed with a device-resident seed."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bc_mpc_b200 import Planner, _lib as L  # noqa: E402
from oracle import mpc_oracle as O  # noqa: E402

rng, norm, weights, state = O.synthetic_problem(0, (500, 500))
for K, H, B in ((1000, 15, 1), (400, 7, 1), (400, 7, 256)):
    pl = Planner(20, 6, 2, 500, "relu", num_paths=K, horizon=H, max_states=B, precision="bf16")
    pl.set_dynamics(weights)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
    host_states = norm["mean_obs"] + norm["std_obs"] * rng.standard_normal((B, 20))
    for i in range(20):
        pl.plan(host_states, seed=i)
    n = 300
    t0 = time.perf_counter()
    for i in range(n):
        pl.plan(host_states, seed=100 + i)
    eager = (time.perf_counter() - t0) / n
    pin_s = torch.from_numpy(host_states).pin_memory()
    pin_o = torch.zeros(B * 80, dtype=torch.uint8).pin_memory()
    d_s = torch.zeros(B, 20, dtype=torch.float64, device="cuda")
    d_o = torch.zeros(B * 80, dtype=torch.uint8, device="cuda")
    seed = torch.tensor([100], dtype=torch.int64, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        d_s.copy_(pin_s, non_blocking=True)
        pl.plan_device(d_s.data_ptr(), B, d_o.data_ptr(), stream_ptr=s.cuda_stream, d_seed_ptr=seed.data_ptr())
        pin_o.copy_(d_o, non_blocking=True)
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        d_s.copy_(pin_s, non_blocking=True)
        pl.plan_device(d_s.data_ptr(), B, d_o.data_ptr(), stream_ptr=torch.cuda.current_stream().cuda_stream, d_seed_ptr=seed.data_ptr())
        seed.add_(1)
        pin_o.copy_(d_o, non_blocking=True)
    for i in range(20):
        g.replay(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        pin_s.numpy()[:] = host_states          # the env's new observation
        g.replay()
        torch.cuda.synchronize()
        rec = np.frombuffer(pin_o.numpy().tobytes(), L.RECORD_DTYPE)
    graph = (time.perf_counter() - t0) / n
    want = pl.plan(host_states, seed=int(seed.item()) - 1)
    assert np.array_equal(rec["index"], want["index"])
    print(f"K={K} H={H} B={B}: Planner.plan {eager * 1e6:.1f} us, graph replay {graph * 1e6:.1f} us per plan")
    pl.close()
