"""CUDA-graph capture of the planning call (bcmpc_plan_device_seeded): the vectorised env loop's plan becomes a graph node,
and every replay draws new action sequences from a device-resident seed.  Replays must equal eager plans bit for bit."""
import numpy as np
import pytest

from oracle import mpc_oracle as O

pytestmark = pytest.mark.gpu


def _records(t):
    from bc_mpc_b200 import _lib as L
    return np.frombuffer(t.cpu().numpy().tobytes(), L.RECORD_DTYPE).copy()


@pytest.mark.parametrize("hidden,precision,proposal,B,K", [((500, 500), "bf16", "uniform", 4, 1000), ((500, 500), "bf16", "policy_gaussian", 3, 700),
                                                          ((256, 256), "bf16", "uniform", 40, 4096), ((1024, 1024, 1024), "bf16", "uniform", 2, 1000),
                                                          ((500, 500), "fp32", "policy_mix", 2, 300)])
def test_plan_captured_in_a_cuda_graph_replays_with_a_device_seed(hidden, precision, proposal, B, K):
    import torch
    from bc_mpc_b200 import Planner
    H = 5
    rng, norm, weights, state = O.synthetic_problem(0, hidden)
    pl = Planner(20, 6, len(hidden), hidden[0], "relu", num_paths=K, horizon=H, max_states=B, precision=precision, proposal=proposal, explore=0.4)
    pl.set_dynamics(weights)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
    if proposal != "uniform":
        pw, logstd, rms_mean, rms_std = O.synthetic_policy(rng, hidden=(128, 128), norm=norm)
        pl.set_policy(pw, logstd, rms_mean, rms_std, clip=5.0)
    host_states = norm["mean_obs"] + norm["std_obs"] * rng.standard_normal((B, 20))
    states = torch.from_numpy(host_states).cuda()
    out = torch.zeros(B * 80, dtype=torch.uint8, device="cuda")
    seed = torch.tensor([7], dtype=torch.int64, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):          # one un-captured call: kernel attributes, scratch allocations
        pl.plan_device(states.data_ptr(), B, out.data_ptr(), stream_ptr=s.cuda_stream, d_seed_ptr=seed.data_ptr())
    s.synchronize()
    eager7 = _records(out)
    ref7 = pl.plan(host_states, seed=7)                  # the host-buffer call with the seed as an argument
    assert np.array_equal(eager7["index"], ref7["index"]) and np.array_equal(eager7["cost"], ref7["cost"])
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        pl.plan_device(states.data_ptr(), B, out.data_ptr(), stream_ptr=torch.cuda.current_stream().cuda_stream, d_seed_ptr=seed.data_ptr())
        seed.add_(1)                                      # the graph advances its own seed
    for step in range(3):
        out.zero_()
        g.replay()
        torch.cuda.synchronize()
        got = _records(out)
        want = pl.plan(host_states, seed=7 + step)
        assert np.array_equal(got["index"], want["index"]) and np.array_equal(got["cost"], want["cost"]), step
        assert np.array_equal(got["action"], want["action"])
    assert int(seed.item()) == 10
    # new root states through the same graph
    host_states = norm["mean_obs"] + norm["std_obs"] * rng.standard_normal((B, 20))
    states.copy_(torch.from_numpy(host_states).cuda())
    g.replay()
    torch.cuda.synchronize()
    want = pl.plan(host_states, seed=10)
    got = _records(out)
    assert np.array_equal(got["index"], want["index"]) and np.array_equal(got["cost"], want["cost"])
    pl.close()
