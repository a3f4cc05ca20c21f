"""Pipeline trace of the tcgen05 rollout kernel (CTA (0,0)): when the MMA warps issued every group, when the epilogue
groups saw each accumulator chunk, when the per-step state update ran.  Clock = SM cycles (clock64).
Usage (GPU box): [TRACE_PROPOSAL=policy_gaussian] python tools/tc_trace.py [S]  -> per-step timeline of steps 1 and 2.
Needs a developer build of the library: BCMPC_NVCC_EXTRA=-DBCMPC_TRACE python -m bc_mpc_b200._build (the trace code
costs ~6 % of the step time, so production builds leave it out; rebuild without the flag afterwards)."""
import os
import struct
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PATH = os.environ.setdefault("BCMPC_TC_TRACE", "/tmp/bcmpc_trace.bin")
from bc_mpc_b200 import Planner  # noqa: E402
from oracle import mpc_oracle as O  # noqa: E402

proposal = os.environ.get("TRACE_PROPOSAL", "uniform")


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    tiles = int(os.environ.get("TRACE_TILES", "2"))
    rng, norm, weights, state = O.synthetic_problem(0, (S, S))
    pl = Planner(20, 6, 2, S, "relu", num_paths=128 * tiles, horizon=6, precision="bf16", proposal=proposal)
    pl.set_dynamics(weights)
    if proposal != "uniform":
        pw, logstd, rms_mean, rms_std = O.synthetic_policy(rng, hidden=(128, 128), norm=norm)
        pl.set_policy(pw, logstd, rms_mean, rms_std, clip=5.0)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
    for i in range(3):
        pl.plan(state, seed=i)
    raw = open(PATH, "rb").read()
    steps, slots, nc, nstg, o_l1, o_l2, o_l3, o_epi, o_fin, o_prod = struct.unpack("10i", raw[:40])
    t = np.frombuffer(raw[40:40 + 8 * steps * slots], np.uint64).reshape(steps, slots).astype(np.int64)
    for h in (1, 2):
        t0 = t[h, o_l1]
        print(f"==== step {h}: cycles relative to the first layer-1 MMA group of the step")
        print(f"  step length (first layer-1 group of this step -> of the next): {t[h + 1, o_l1] - t0}")
        print("  layer-1 issuer (warp 3): pair | issue-start issue-end")
        for c0 in range(0, nc, 2):
            print(f"    c{c0}+{min(2, nc - c0)} | {t[h, o_l1 + c0] - t0:7d} {t[h, o_l1 + c0 + 1] - t0:7d}")
        print("  layer-2 issuers (even chunks warp 1, odd chunks warp 3): chunk | first-group issue-start  last-group issue-end (dur)")
        for c in range(nc):
            a, b = t[h, o_l2 + 2 * c], t[h, o_l2 + 2 * c + 1]
            print(f"    c{c} | {a - t0:7d} {b - t0:7d} ({b - a:5d})")
        print("  layer-3 issuer (warp 2): chunk | issue-start issue-end")
        for c in range(nc):
            print(f"    c{c} | {t[h, o_l3 + 2 * c] - t0:7d} {t[h, o_l3 + 2 * c + 1] - t0:7d}")
        print("  epilogue (quadrant-0 warp of the group): chunk | accumulator seen  +drained  +dest free  +stored (total)")
        for e in range(2 * nc):
            a, b, c, d = (t[h, o_epi + 4 * e + j] for j in range(4))
            grp = ((e // 2) & 1) if e < nc else (e & 1)
            print(f"    {'L2' if e >= nc else 'L1'} c{e % nc} g{grp} | {a - t0:7d}  +{b - a:4d}  +{c - b:4d}  +{d - c:4d} ({d - a:4d})")
        f0, f1 = t[h, o_fin], t[h, o_fin + 1]
        print(f"  final: output accumulator seen {f0 - t0}, this step's input was ready at {f1 - t0}")
        if proposal != "uniform":
            names = ["policy input published", "policy L1 accumulator seen", "policy L1 chunk stored", "policy L2 accumulator seen",
                     "policy L2 chunk stored", "policy output seen"]
            print("  proposal network (before the first layer-1 group): " + ", ".join(f"{n} {t[h, o_fin + 2 + i] - t0}" for i, n in enumerate(names)))
        print("  producer: stage | slot-wait-start copy-issued")
        for i in range(nc * nstg):
            print(f"    {i:2d} | {t[h, o_prod + 2 * i] - t0:7d} {t[h, o_prod + 2 * i + 1] - t0:7d}")


if __name__ == "__main__":
    main()
