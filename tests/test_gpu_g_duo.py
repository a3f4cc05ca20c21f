"""rollout_duo_kernel (csrc/rollout_duo.cuh): two tiles per CTA taking turns on one hidden-1 buffer, for hidden 257..510
when a launch has more than one wave of CTA-pair units.  It reads the weight images of rollout_tmem_kernel and must give
bit-identical results: same k-order in every accumulator, same epilogues (dynamics.py:106-119, cost_functions.py:10-30,
controllers.py:53,63-85)."""
import os

import numpy as np
import pytest

from oracle import mpc_oracle as O
from tests._util import f32_uniform, make_pair

pytestmark = pytest.mark.gpu


def _pair(monkeypatch, mode, seed, S, K, H, **kw):
    # BCMPC_DUO: 0 = never, 1 = when the launch has more CTA-pair units than SM pairs, 2 = always where it applies
    monkeypatch.setenv("BCMPC_DUO", mode)
    return make_pair(seed, (S, S), K, H, precision="bf16", **kw)


def _skip_single_cta():
    if os.environ.get("BCMPC_TC_CG") == "1":
        pytest.skip("the two-tile kernel is a CTA-pair kernel")


@pytest.mark.parametrize("S", [500, 300, 384, 446])
def test_duo_kernel_tracks_oracle_and_equals_the_wide_kernel(monkeypatch, S):
    """Per-path states / actions / costs of the debug flavour against the oracle and, bit for bit, against rollout_tmem_kernel;
    K = 1000 -> 4 CTA-pair units on one pair each when forced (every pipe alone ... no: units spread over pairs), so also K large
    enough that pairs get several units (both pipes active, odd unit counts: one pipe finishes alone)."""
    _skip_single_cta()
    K, H = 1000, 9
    rng, norm, state, pl, mpc = _pair(monkeypatch, "2", 41, S, K, H)
    actions = f32_uniform(rng, (H, K, 6))
    act, dbg = mpc.get_action(state, actions, return_debug=True)
    so, ao, co = pl.rollout_debug(state, actions)
    sd = np.where(norm["std_deltas"] > 0, norm["std_deltas"], 1.0)
    e = np.abs((so - dbg["states"]) / sd)
    rec = pl.plan(state, inj=actions)[0]
    assert np.array_equal(ao, actions) and e.max() < 0.3 and np.corrcoef(co, dbg["costs"])[0, 1] > 0.98
    assert rec["index"] == int(np.argmin(co)) and np.array_equal(rec["action"][:6], actions[0, rec["index"]])
    _, _, _, wide, _ = _pair(monkeypatch, "0", 41, S, K, H)
    sw, aw, cw = wide.rollout_debug(state, actions)
    assert np.array_equal(sw, so) and np.array_equal(cw, co)
    pl.close(); wide.close()


@pytest.mark.parametrize("B,K", [(1, 128 * 2 * 74 * 3), (5, 128 * 2 * 37), (3, 700), (160, 256), (2, 33)])
def test_duo_kernel_batches_equal_the_wide_kernel(monkeypatch, B, K):
    """Production flavour, many units per CTA pair (3 per pair; 2.5 per pair: odd counts, one pipe finishes alone; fewer units
    than pairs; one unit per state; a ragged single tile): records bit-identical to rollout_tmem_kernel and to the debug-flavour
    replay of the device-drawn actions."""
    _skip_single_cta()
    H = 4
    rng, norm, state, pl, mpc = _pair(monkeypatch, "2", 43, 500, K, H, max_states=B)
    states = norm["mean_obs"] + norm["std_obs"] * rng.standard_normal((B, 20))
    got = pl.plan(states, seed=7)
    _, _, _, wide, _ = _pair(monkeypatch, "0", 43, 500, K, H, max_states=B)
    ref = wide.plan(states, seed=7)
    assert np.array_equal(got["index"], ref["index"]) and np.array_equal(got["cost"], ref["cost"]) and np.array_equal(got["action"], ref["action"])
    if B * K <= 200000:
        replay = pl.plan(states, inj=pl.sample_injection(7, B))
        assert np.array_equal(got["index"], replay["index"]) and np.array_equal(got["cost"], replay["cost"])
    pl.close(); wide.close()


def test_duo_kernel_threshold_mode(monkeypatch):
    """BCMPC_DUO=1: a 32-state x K = 4096 batch (512 CTA-pair units) runs on rollout_duo_kernel, a single K = 1000 plan stays on
    rollout_tmem_kernel; both agree with the forced settings."""
    _skip_single_cta()
    monkeypatch.setenv("BCMPC_DUO", "1")
    rng, norm, state, pl, mpc = make_pair(44, (500, 500), 4096, 5, precision="bf16", max_states=32)
    states = norm["mean_obs"] + norm["std_obs"] * rng.standard_normal((32, 20))
    auto = pl.plan(states, seed=3)
    _, _, _, wide, _ = _pair(monkeypatch, "0", 44, 500, 4096, 5, max_states=32)
    ref = wide.plan(states, seed=3)
    assert np.array_equal(auto["index"], ref["index"]) and np.array_equal(auto["cost"], ref["cost"])
    pl.close(); wide.close()


def test_duo_kernel_plain_tanh_net(monkeypatch):
    """The plain tanh NNDynamicsModel (dynamics.py:60, explicit biases) on the two-tile kernel equals the wide kernel."""
    _skip_single_cta()
    from bc_mpc_b200 import Planner
    rng, norm, weights, state = O.synthetic_problem(45, (500, 500))
    res = {}
    for mode in ("2", "0"):
        monkeypatch.setenv("BCMPC_DUO", mode)
        pl = Planner(20, 6, 2, 500, "tanh", num_paths=128 * 2 * 74 * 2 + 100, horizon=4, precision="bf16", max_states=2)
        pl.set_dynamics(weights)
        pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
        states = norm["mean_obs"] + norm["std_obs"] * np.random.default_rng(5).standard_normal((2, 20))
        res[mode] = pl.plan(states, seed=11)
        pl.close()
    assert np.array_equal(res["2"]["index"], res["0"]["index"]) and np.array_equal(res["2"]["cost"], res["0"]["cost"])
