"""CPU-side checks: the C-ABI library builds/loads and exports every symbol include/bcmpc.h declares
(no compute calls without a GPU), struct layouts match, and the Python host logic behaves."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from bc_mpc_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "bcmpc.h")).read()
    declared = set(re.findall(r"\b(bcmpc_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 19
    lib = _lib.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.bcmpc_abi_version() == _lib.ABI_VERSION == 4


def test_struct_layouts_match_header():
    from bc_mpc_b200 import _lib
    assert ctypes.sizeof(_lib.Record) == 80 and _lib.RECORD_DTYPE.itemsize == 80
    assert _lib.Config.path_offset.offset == 48 and _lib.Config.explore.offset == 64
    assert ctypes.sizeof(_lib.Config) == 72


def test_no_gpu_fails_loudly_not_silently():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from bc_mpc_b200 import Planner
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Planner()


def test_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "bc_mpc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+\S*oracle", txt, re.M), f"{f} imports the oracle"
                assert "mpc_oracle" not in txt, f"{f} references the oracle module"


class _Space:
    def __init__(self, n):
        self.low, self.high, self.shape = -np.ones(n), np.ones(n), (n,)

    def sample(self):
        return np.random.uniform(self.low, self.high)


class _Env:
    action_space, observation_space = _Space(6), _Space(20)


def test_controller_host_logic(golden_dir):
    from bc_mpc_b200 import controllers, cost_functions, dynamics, policy
    env = _Env()
    g = np.load(os.path.join(golden_dir, "vars_ppo_mpc.npz"))
    var = {}
    for i, nm in enumerate(["dense", "dense_1", "dense_2"]):
        var[f"NNDynamicsModel/{nm}/kernel:0"], var[f"NNDynamicsModel/{nm}/bias:0"] = g[f"dyn_w{i}"], g[f"dyn_b{i}"]
    norm = tuple(np.zeros(20) for _ in range(10))
    dm = dynamics.NNDynamicsModel(env, 2, 256, "relu", None, norm, 512, 60, 1e-3, None)
    assert dm.version == 0
    dm.load_vars(var)
    assert dm.version == 1 and dm.weights[1][0].shape == (256, 256)
    with pytest.raises(NotImplementedError):
        dm.fit(None)
    with pytest.raises(ValueError, match="cheetah_cost_fn"):
        controllers.MPCcontroller(env, dm, 5, lambda s, a, n: 0, 10)
    c = controllers.MPCcontroller(env, dm, horizon=3, cost_fn=cost_functions.cheetah_cost_fn, num_simulated_paths=7)
    assert not c._sampling_overridden()
    a = c.sample_random_actions()
    assert a.shape == (3, 7, 6) and a.dtype == np.float64           # controllers.py:53 layout
    c.sample_random_actions = lambda: a
    assert c._sampling_overridden()

    class Sub(controllers.MPCcontroller):
        def sample_random_actions(self):
            return a
    assert Sub(env, dm, 3, cost_functions.cheetah_cost_fn, 7)._sampling_overridden()
    assert controllers.RandomController(env).get_action(None).shape == (6,)
    pol = policy.MlpPolicy(None, env, 128, 2)
    var.update({f"pi/pi/pol/{n}/kernel:0": g[f"pol_w{i}"] for i, n in enumerate(["fc1", "fc2", "final"])})
    var.update({f"pi/pi/pol/{n}/bias:0": g[f"pol_b{i}"] for i, n in enumerate(["fc1", "fc2", "final"])})
    var.update({"pi/pi/pol/logstd:0": g["pol_logstd"], "pi/pi/obfilter/runningsum:0": g["ob_sum"],
                "pi/pi/obfilter/runningsumsq:0": g["ob_sumsq"], "pi/pi/obfilter/count:0": g["ob_count"]})
    pol.load_vars(var)
    from oracle import mpc_oracle as O
    m, s = O.OraclePolicy.rms_from_sums(g["ob_sum"], g["ob_sumsq"], float(g["ob_count"]))
    assert np.array_equal(pol.filt_mean, m) and np.array_equal(pol.filt_std, s) and pol.logstd.shape == (6,)
    pc = controllers.MPCcontrollerPolicyNet(env, dm, pol, explore=0.5, self_exp=False, horizon=3,
                                            cost_fn=cost_functions.cheetah_cost_fn, num_simulated_paths=7)
    assert pc.proposal == "policy_mix"


def test_samplers_host_logic():
    """sampling.sample mirrors utils.py:91-126; sample_vectorized makes one batched call per time step."""
    from bc_mpc_b200 import sampling

    class E:
        def __init__(self, k):
            self.k, self.t = k, 0

        def reset(self):
            self.t = 0
            return np.full(3, float(self.k))

        def step(self, a):
            self.t += 1
            return np.full(3, self.k + 0.1 * self.t) + a.sum(), float(self.t), False, {}

    class Ctl:
        calls = batched = 0

        def get_action(self, s):
            self.calls += 1
            return np.array([s[0], -s[0]])

        def get_actions(self, S):
            self.batched += 1
            return np.stack([S[:, 0], -S[:, 0]], axis=1)

    c = Ctl()
    one = sampling.sample(E(2), c, num_paths=3, horizon=4)
    assert len(one) == 3 and c.calls == 12 and len(one[0]["observations"]) == 4
    assert np.array_equal(one[0]["next_observations"][0], one[0]["observations"][1])
    vec = sampling.sample_vectorized([E(k) for k in range(5)], c, horizon=4)
    assert c.batched == 4 and len(vec) == 5
    ref = sampling.sample(E(2), c, num_paths=1, horizon=4)[0]
    for key in ref:
        assert np.allclose(np.asarray(vec[2][key]), np.asarray(ref[key]))
    assert sampling.path_cost(lambda s, a, n: float(np.sum(n) - np.sum(s)), vec[1]) == pytest.approx(
        sum(float(np.sum(n) - np.sum(s)) for s, n in zip(vec[1]["observations"], vec[1]["next_observations"])))


def test_segment_generators_host_logic():
    """sampling.traj_segment_generator follows utils.py:163-242; the vectorised form makes one planner call per step."""
    from bc_mpc_b200 import sampling

    class Space:
        def sample(self):
            return np.zeros(2, np.float32)

    class E:
        action_space = Space()

        def __init__(self, k):
            self.k, self.t = k, 0

        def reset(self):
            self.t = 0
            return np.full(3, float(self.k))

        def step(self, a):
            self.t += 1
            return np.full(3, self.k + 0.5 * self.t) + float(np.sum(a)), 1.0 + self.k, False, {}

    class Pi:
        def act(self, ob, stochastic=True):
            return np.array([ob[0], 1.0], np.float32), float(ob[0]) * 0.1

    class Ctl:
        single = batched = 0

        def get_action(self, s):
            self.single += 1
            return np.array([0.25 * s[0], 0.0])

        def get_actions(self, S):
            self.batched += 1
            return np.stack([0.25 * S[:, 0], np.zeros(len(S))], axis=1)

    c = Ctl()
    seg = sampling.traj_segment_generator(Pi(), c, None, None, E(2), True, False, 5)
    assert c.single == 5 and seg["ob"].shape == (5, 3) and seg["new"].tolist() == [1, 0, 0, 0, 0]
    assert np.allclose(seg["mpcac"][0], [0.5, 0.0]) and np.allclose(seg["ac"][0], [2.0, 1.0])
    assert np.allclose(seg["prevac"][1], seg["ac"][0]) and np.array_equal(seg["nxt_ob"][:-1], seg["ob"][1:])
    assert seg["ep_lens"] == [5] and seg["ep_rets"] == [15.0] and seg["nextvpred"] == pytest.approx(float(seg["ob"][4][0]) * 0.1)
    segs = sampling.traj_segments_vectorized(Pi(), c, [E(k) for k in range(4)], True, 5)
    assert c.batched == 5 and len(segs) == 4
    for key in ("ob", "rew", "nxt_ob", "vpred", "new", "ac", "prevac", "mpcac"):
        assert np.allclose(segs[2][key], seg[key])
    nompc = sampling.traj_segment_generator(Pi(), c, None, None, E(1), False, False, 3)
    assert np.allclose(nompc["mpcac"], nompc["ac"])


def test_shard_bounds():
    from bc_mpc_b200.distributed import shard_bounds
    assert list(shard_bounds(1000, 4)) == [0, 250, 500, 750, 1000]
    b = shard_bounds(1003, 8)
    assert b[0] == 0 and b[-1] == 1003 and np.all(np.diff(b) >= 125) and np.all(np.diff(b) <= 126)
    assert list(shard_bounds(3, 4)) == [0, 1, 2, 3, 3]


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from bc_mpc_b200 import _lib as L
    from bc_mpc_b200.distributed import gather_records, shard_bounds
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    B = 3
    bounds = shard_bounds(10, world)
    rec = np.zeros(B, L.RECORD_DTYPE)
    for b in range(B):
        rec[b]["cost"] = [5.0, 1.0][rank] if b == 0 else 2.0     # state 0: rank 1 wins; others tie
        rec[b]["index"] = bounds[rank] + b
        rec[b]["action"][:6] = rank * 10 + b
    t = torch.from_numpy(np.frombuffer(rec.tobytes(), np.uint8).copy())
    g = gather_records(t)
    arr = np.frombuffer(g.numpy().tobytes(), L.RECORD_DTYPE).reshape(world, B)
    q.put((rank, arr["cost"].tolist(), arr["index"].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_record_exchange():
    """N>1 host path on CPU: the all-gather layout is rank-major [world, B] and a min-loc with the
    global index as tie-break picks the same winner on every rank."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert res[0][1:] == res[1][1:]                       # every rank holds the same gathered table
    cost, index = np.array(res[0][1]), np.array(res[0][2])
    assert cost.shape == (2, 3) and list(index[:, 0]) == [0, 5]
    win = [min(range(2), key=lambda r: (cost[r, b], index[r, b])) for b in range(3)]
    assert win == [1, 0, 0]                                # ties go to the lowest global index
