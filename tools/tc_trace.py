"""Pipeline trace of the tcgen05 rollout kernel (CTA (0,0)): when the MMA warps issued every group, when the epilogue
groups saw each accumulator chunk, when the per-step state update ran.  Clock = SM cycles (clock64).
Usage (GPU box): BCMPC_TC_CG=2 python tools/tc_trace.py [S]   -> prints a per-step timeline of steps 1 and 2."""
import os
import struct
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PATH = os.environ.setdefault("BCMPC_TC_TRACE", "/tmp/bcmpc_trace.bin")
from bc_mpc_b200 import Planner  # noqa: E402
from oracle import mpc_oracle as O  # noqa: E402


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    tiles = int(os.environ.get("TRACE_TILES", "2"))
    rng, norm, weights, state = O.synthetic_problem(0, (S, S))
    pl = Planner(20, 6, 2, S, "relu", num_paths=128 * tiles, horizon=6, precision="bf16")
    pl.set_dynamics(weights)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
    for i in range(3):
        pl.plan(state, seed=i)
    raw = open(PATH, "rb").read()
    steps, slots, n_ops, nc, o_mma, o_l3, o_epi, o_fin, o_prod, n_stream = struct.unpack("10i", raw[:40])
    t = np.frombuffer(raw[40:40 + 8 * steps * slots], np.uint64).reshape(steps, slots).astype(np.int64)
    ops = np.frombuffer(raw[40 + 8 * steps * slots:], np.int32).reshape(n_ops, 6)   # kind, chunk, kb0, nblk, buffer, stage-first
    for h in (1, 2):
        t0 = t[h, o_mma]
        print(f"==== step {h}: cycles relative to the first layer-1 MMA group of the step")
        print(f"  step length (first group of this step -> first group of the next): {t[h + 1, o_mma] - t0}")
        print("  layer-1/2 issuer: idx group | issue-start  issue-end (dur)  gap-since-previous-end")
        prev = t0
        for i in range(n_ops):
            kd, ch, kb0, nblk, bf, sf = ops[i]
            name = f"L1 c{ch}+{nblk} b{bf}" if kd == 0 else f"L2 c{ch} kb{kb0}+{nblk} b{bf}{' W' if sf else ''}"
            a, b = t[h, o_mma + 2 * i], t[h, o_mma + 2 * i + 1]
            print(f"    {i:2d} {name:18s} | {a - t0:7d} {b - t0:7d} ({b - a:4d})  gap {a - prev:5d}")
            prev = b
        print("  layer-3 issuer: chunk | issue-start issue-end")
        for c in range(nc):
            print(f"    {c} | {t[h, o_l3 + 2 * c] - t0:7d} {t[h, o_l3 + 2 * c + 1] - t0:7d}")
        print("  epilogue (quadrant-0 warp of the group): chunk | accumulator seen  +drained  +dest free  +stored (total)")
        for e in range(2 * nc):
            a, b, c, d = (t[h, o_epi + 4 * e + j] for j in range(4))
            print(f"    {'L2' if e >= nc else 'L1'} c{e % nc} g{e & 1} | {a - t0:7d}  +{b - a:4d}  +{c - b:4d}  +{d - c:4d} ({d - a:4d})")
        f0, f1 = t[h, o_fin], t[h, o_fin + 1]
        print(f"  final: output accumulator seen {f0 - t0}, next input ready {f1 - t0} ({f1 - f0})")
        print("  producer: stage | slot-wait-start copy-issued")
        for i in range(n_stream):
            print(f"    {i:2d} | {t[h, o_prod + 2 * i] - t0:7d} {t[h, o_prod + 2 * i + 1] - t0:7d}")


if __name__ == "__main__":
    main()
