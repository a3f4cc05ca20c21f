"""Pipeline trace of rollout_duo_kernel (CTA 0): per pipe and pipe step, when the issuers and the pipe's epilogue group passed
the marks of csrc/rollout_duo.cuh (kDuTr*).  Needs a developer build: BCMPC_NVCC_EXTRA=-DBCMPC_TRACE python -m bc_mpc_b200._build
Usage (GPU box): python tools/duo_trace.py [S]"""
import os
import struct
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PATH = os.environ.setdefault("BCMPC_TC_TRACE", "/tmp/bcmpc_trace.bin")
os.environ["BCMPC_DUO"] = "2"
from bc_mpc_b200 import Planner  # noqa: E402
from oracle import mpc_oracle as O  # noqa: E402

NAMES = {0: "L1 first issue", 1: "L1 last issue", 2: "L2 first issue", 3: "L2 last issue end", 4: "L3 last issue", 8: "A0 published", 9: "A1 free seen",
         10: "L1 c0 stored", 11: "A1 all stored", 12: "L2 c0 stored", 13: "L2 all stored", 14: "D3 seen", 15: "state updated"}


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    rng, norm, weights, state = O.synthetic_problem(0, (S, S))
    pl = Planner(20, 6, 2, S, "relu", num_paths=128 * 2 * 74 * int(os.environ.get("DUO_TRACE_UNITS", "2")), horizon=8, precision="bf16")     # two units per CTA pair: both pipes
    pl.set_dynamics(weights)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
    for i in range(3):
        pl.plan(state, seed=i)
    raw = open(PATH, "rb").read()
    hdr = struct.unpack("10i", raw[:40])
    assert hdr[0] == -2, hdr
    steps, slots, nc = hdr[1:4]
    t = np.frombuffer(raw[40:40 + 8 * 2 * steps * slots], np.uint64).reshape(2, steps, slots).astype(np.int64)
    t0 = t[0, 1, 2]
    print(f"cycles relative to pipe 0's first layer-2 issue of step 1; hidden chunks {nc}")
    print(f"turn length: pipe0 L2 first issue step1 -> pipe1 step1: {t[1, 1, 2] - t[0, 1, 2]}, -> pipe0 step2: {t[0, 2, 2] - t[0, 1, 2]}, pipe1 step1 -> step2: {t[1, 2, 2] - t[1, 1, 2]}")
    for s in (1, 2, 3):
        for pipe in (0, 1):
            print(f"  step {s} pipe {pipe}: " + ", ".join(f"{NAMES[k]} {t[pipe, s, k] - t0}" for k in sorted(NAMES)))
            if slots >= 56:
                print("      L1 chunk: issued / seen / stored : " + "  ".join(f"c{c} {t[pipe, s, 16 + c] - t0}/{t[pipe, s, 24 + c] - t0}/{t[pipe, s, 32 + c] - t0}" for c in range(nc)))
                print("      L2 chunk: issue start / seen     : " + "  ".join(f"c{c} {t[pipe, s, 48 + c] - t0}/{t[pipe, s, 40 + c] - t0}" for c in range(nc)))


if __name__ == "__main__":
    main()
