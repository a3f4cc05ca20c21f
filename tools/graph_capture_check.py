"""Can a plan be captured into a CUDA graph (torch.cuda.graph) and replayed?  Checks the three fast kernels and the
verification kernel: replayed records equal the eager ones (the seed is baked into the captured launch)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bc_mpc_b200 import Planner, _lib as L  # noqa: E402
from oracle import mpc_oracle as O  # noqa: E402

for hidden, prec, B, K in (((500, 500), "bf16", 4, 1000), ((256, 256), "bf16", 40, 4096), ((1024, 1024, 1024), "bf16", 2, 1000), ((500, 500), "fp32", 2, 500)):
    rng, norm, weights, state = O.synthetic_problem(0, hidden)
    pl = Planner(20, 6, len(hidden), hidden[0], "relu", num_paths=K, horizon=5, max_states=B, precision=prec)
    pl.set_dynamics(weights)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
    states = torch.from_numpy(norm["mean_obs"] + norm["std_obs"] * rng.standard_normal((B, 20))).cuda()
    out = torch.zeros(B * 80, dtype=torch.uint8, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(2):      # warm-up outside the capture (lazy attribute set-up, allocations)
            pl.plan_device(states.data_ptr(), B, out.data_ptr(), seed=7, stream_ptr=s.cuda_stream)
    s.synchronize()
    eager = np.frombuffer(out.cpu().numpy().tobytes(), L.RECORD_DTYPE).copy()
    g = torch.cuda.CUDAGraph()
    out.zero_()
    with torch.cuda.graph(g, stream=s):
        pl.plan_device(states.data_ptr(), B, out.data_ptr(), seed=7, stream_ptr=torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        out.zero_()
        g.replay()
        torch.cuda.synchronize()
        rep = np.frombuffer(out.cpu().numpy().tobytes(), L.RECORD_DTYPE)
        assert np.array_equal(rep["index"], eager["index"]) and np.array_equal(rep["cost"], eager["cost"]), (hidden, prec)
    # new root states through the same graph
    states.copy_(torch.from_numpy(norm["mean_obs"] + norm["std_obs"] * rng.standard_normal((B, 20))).cuda())
    g.replay()
    torch.cuda.synchronize()
    rep = np.frombuffer(out.cpu().numpy().tobytes(), L.RECORD_DTYPE).copy()
    pl.plan_device(states.data_ptr(), B, out.data_ptr(), seed=7, stream_ptr=0)
    torch.cuda.synchronize()
    again = np.frombuffer(out.cpu().numpy().tobytes(), L.RECORD_DTYPE)
    assert np.array_equal(rep["index"], again["index"]) and np.array_equal(rep["cost"], again["cost"])
    print("graph capture ok:", hidden, prec, "B", B, "K", K)
    pl.close()
