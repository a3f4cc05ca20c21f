"""Multi-GPU planning: paths shard over the ranks of one box (one process per GPU), weights are
replicated, and the only exchange per plan is one 80-byte record per root state and rank, min-loc'ed
with ties broken by GLOBAL path index, so the result does not depend on the shard count (np.argmin,
controllers.py:82).  Two transports with identical results:
  "peer" (default on NVLink boxes): every rank's exchange kernel stores its records straight into every peer's
         buffer (torch symmetric memory = peer-mapped device pointers), flags them, waits for the peers' flags and
         reduces - one kernel after the rollout (bcmpc_set_peer_exchange);
  "nccl": ncclAllGather of the records + a local min-loc kernel (NCCL has no MINLOC) - also the path the gloo CPU
         tests exercise.  BCMPC_EXCHANGE=nccl / peer forces one."""
import numpy as np

from . import _lib as L
from .planner import Planner


def shard_bounds(total_paths, world_size):
    """Contiguous shards; rank r owns global path indices [bounds[r], bounds[r+1])."""
    base, rem = divmod(int(total_paths), int(world_size))
    sizes = [base + (1 if r < rem else 0) for r in range(world_size)]
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


def gather_records(local_records, group=None):
    """local_records: torch uint8 tensor [B*80] (or [B,80]) on this rank's device.  Returns the
    rank-major gathered tensor [world, B, 80].  Works with nccl (GPU) and gloo (CPU tests)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    flat = local_records.reshape(-1).contiguous()
    out = torch.empty(world * flat.numel(), dtype=flat.dtype, device=flat.device)
    dist.all_gather_into_tensor(out, flat, group=group)
    return out.reshape(world, -1, L.RECORD_DTYPE.itemsize)


class ShardedPlanner:
    """Rank-local planner over this rank's share of the K paths + the cross-rank min-loc."""

    def __init__(self, total_paths, horizon, rank=None, world_size=None, group=None, device=None, exchange=None, **planner_kw):
        import os

        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world_size is None else world_size
        self.bounds = shard_bounds(total_paths, self.world)
        lo, hi = int(self.bounds[self.rank]), int(self.bounds[self.rank + 1])
        self.device = torch.cuda.current_device() if device is None else device
        self.planner = Planner(num_paths=hi - lo, horizon=horizon, path_offset=lo, total_paths=total_paths,
                               device=self.device, **planner_kw)
        B = self.planner.max_states
        dev = torch.device("cuda", self.device)
        self.local = torch.zeros(B * 80, dtype=torch.uint8, device=dev)
        self.result = torch.zeros(B * 80, dtype=torch.uint8, device=dev)
        self.exchange = "nccl"
        want = exchange or os.environ.get("BCMPC_EXCHANGE", "auto")
        if self.world > 1 and want in ("auto", "peer"):
            try:
                self._setup_peer_exchange(dev)
                self.exchange = "peer"
            except Exception as e:           # no symmetric memory on this box / backend: NCCL carries the records
                if want == "peer":
                    raise
                self.peer_error = repr(e)

    def _setup_peer_exchange(self, dev):
        """One peer-mapped buffer per rank (torch symmetric memory over the group), zero-filled; every rank hands the
        pointers of all ranks' buffers to its handle."""
        torch, dist = self.torch, self.dist
        import torch.distributed._symmetric_memory as symm_mem
        nbytes = self.planner.peer_exchange_bytes(self.world)
        group = self.group if self.group is not None else dist.group.WORLD
        self._symm = symm_mem.empty(nbytes, dtype=torch.uint8, device=dev)
        self._symm.zero_()
        self._symm_hdl = symm_mem.rendezvous(self._symm, group.group_name)
        torch.cuda.synchronize(dev)
        dist.barrier(group)                                   # every rank's buffer is zero before anyone writes a flag
        self.planner.set_peer_exchange(self.world, self.rank, list(self._symm_hdl.buffer_ptrs))

    def plan_device(self, d_states, n_states, seed=0):
        """d_states: torch float64 [B,D] on this rank's GPU.  Enqueues rollout + all-gather + min-loc on
        the current torch stream; returns the device tensor of [B] winning records (uint8 view)."""
        torch = self.torch
        stream = torch.cuda.current_stream().cuda_stream
        self.planner.plan_device(d_states.data_ptr(), n_states, self.local.data_ptr(), seed=seed, stream_ptr=stream)
        if self.world == 1 or self.exchange == "peer":       # (peer exchange: the plan itself was the collective)
            return self.local[: n_states * 80]
        gathered = gather_records(self.local[: n_states * 80], self.group)
        self.planner.reduce_records_device(gathered.data_ptr(), self.world, n_states, self.result.data_ptr(), stream)
        return self.result[: n_states * 80]

    def plan(self, states, seed=0):
        torch = self.torch
        states = np.ascontiguousarray(states, np.float64).reshape(-1, self.planner.D)
        d = torch.from_numpy(states).to(torch.device("cuda", self.device))
        out = self.plan_device(d, states.shape[0], seed)
        return np.frombuffer(out.cpu().numpy().tobytes(), L.RECORD_DTYPE)
