#!/usr/bin/env python
"""Multi-GPU A/B of the two record-exchange transports of ShardedPlanner (run under torchrun, one rank per GPU):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_check.py
Checks that the NVLink peer-store exchange (bcmpc_set_peer_exchange) gives bit-identical winners to ncclAllGather + min-loc
and to a single-GPU plan over all K paths, on every rank, and times both (CUDA events, max over ranks)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from bc_mpc_b200 import Planner, _lib as L
    from bc_mpc_b200.distributed import ShardedPlanner
    from oracle import mpc_oracle as O      # synthetic inputs only
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K, H, B = 4096, 20, 32 * world
    rng, norm, weights, state = O.synthetic_problem(0, (500, 500))
    states = norm["mean_obs"] + norm["std_obs"] * np.random.default_rng(1).standard_normal((B, 20))
    d_states = torch.from_numpy(states).cuda()

    def make(exchange):
        sp = ShardedPlanner(K, H, rank=rank, world_size=world, device=local, exchange=exchange, obs_dim=20, act_dim=6, n_hidden=2,
                            hidden_size=500, activation="relu", max_states=B, precision="bf16", proposal="uniform")
        sp.planner.set_dynamics(weights)
        sp.planner.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
        return sp

    peer, nccl = make("peer"), make("nccl")
    assert peer.exchange == "peer" and nccl.exchange == "nccl"
    ok = True
    for seed in range(6):
        a = np.frombuffer(peer.plan_device(d_states, B, seed=seed).cpu().numpy().tobytes(), L.RECORD_DTYPE)
        b = np.frombuffer(nccl.plan_device(d_states, B, seed=seed).cpu().numpy().tobytes(), L.RECORD_DTYPE)
        ok &= bool(np.array_equal(a["index"], b["index"]) and np.array_equal(a["cost"], b["cost"]) and np.array_equal(a["action"], b["action"]))
        if seed == 0 and rank == 0:      # against ONE handle over all K paths of the first 4 states
            one = Planner(20, 6, 2, 500, "relu", num_paths=K, horizon=H, max_states=4, precision="bf16", device=local)
            one.set_dynamics(weights)
            one.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
            w = one.plan(states[:4], seed=0)
            ok &= bool(np.array_equal(w["index"], a["index"][:4]) and np.array_equal(w["cost"], a["cost"][:4]))
            one.close()
    # every rank holds the same winners
    t = torch.from_numpy(a["index"].copy()).cuda()
    ref = t.clone()
    dist.broadcast(ref, 0)
    ok &= bool(torch.equal(t, ref))
    stream = torch.cuda.current_stream()
    times = {}
    for name, sp in (("peer", peer), ("nccl", nccl)):
        for i in range(5):
            sp.plan_device(d_states, B, seed=100 + i)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 200
        e0.record(stream)
        for i in range(n):
            sp.plan_device(d_states, B, seed=200 + i)
        e1.record(stream)
        dist.barrier(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        times[name] = float(ms.item())
    flag = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "states": B, "K": K, "H": H, "identical_on_all_ranks": bool(flag.item()),
                          "ms_per_plan_peer_exchange": times["peer"], "ms_per_plan_nccl_allgather": times["nccl"],
                          "transitions_per_s_peer": B * K * H / times["peer"] * 1e3, "transitions_per_s_nccl": B * K * H / times["nccl"] * 1e3}))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
