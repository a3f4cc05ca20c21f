"""Pipeline trace of the cluster-split rollout kernel (CTA 0 of cluster 0): per accumulator use, when the issuer waited /
issued / committed and when the epilogue groups drained, exchanged and stored.  Clock = SM cycles (clock64).
Usage (GPU box): python tools/deep_trace.py [hidden=1024,1024,1024] [units=40]
Needs a developer build: BCMPC_NVCC_EXTRA=-DBCMPC_TRACE python -m bc_mpc_b200._build (rebuild without the flag afterwards)."""
import os
import struct
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PATH = os.environ.setdefault("BCMPC_TC_TRACE", "/tmp/bcmpc_deep_trace.bin")
from bc_mpc_b200 import Planner  # noqa: E402
from oracle import mpc_oracle as O  # noqa: E402


def main():
    hidden = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1024,1024,1024").split(","))
    units = int(sys.argv[2]) if len(sys.argv) > 2 else 40          # 256-path units: > 33 keeps every cluster busy
    rng, norm, weights, state = O.synthetic_problem(0, hidden)
    pl = Planner(20, 6, len(hidden), hidden[0], "relu", num_paths=256 * units, horizon=6, precision="bf16")
    pl.set_dynamics(weights)
    pl.set_normalization(norm["mean_obs"], norm["std_obs"], norm["mean_action"], norm["std_action"], norm["mean_deltas"], norm["std_deltas"])
    for i in range(3):
        pl.plan(state, seed=i)
    raw = open(PATH, "rb").read()
    steps, roles, uses, stamps, nch, L = struct.unpack("6i", raw[:24])
    t = np.frombuffer(raw[24:24 + 8 * steps * roles * uses * stamps], np.uint64).reshape(steps, roles, uses, stamps).astype(np.int64)
    n_use = nch + (L - 1) * 2 * nch + 1

    def name(u):
        if u < nch:
            return f"L1 c{u}"
        if u == n_use - 1:
            return "OUT"
        v = u - nch
        return f"l{2 + v // (2 * nch)} {'FO'[v % 2]}{(v % (2 * nch)) // 2}"

    for h in (1, 2):
        t0 = t[h, 0, 0, 1]
        print(f"==== step {h} (cycles relative to the issue of the step's first use); step length {t[h + 1, 0, 0, 1] - t0}")
        print("  use        | issuer: wait-begin  acc-free  weights-landed  input-stored  token/issue  issued | epilogue: start  drained  xchg-wait-done  sent/recv-done  stored (OUT: recv-wait recv-done state-done)")
        for u in range(n_use):
            # issuer stamps: 0 wait-begin, 3 accumulator free, 4 weight stages landed, 5 layer input stored, 1 token (issue starts), 2 issued
            order = [0, 3, 4, 5, 1, 2]
            w = 0 if t[h, 0, u, 0] else 3
            iss = [int(t[h, w, u, k] - t0) if t[h, w, u, k] else None for k in order]
            g = 1 if (u < nch and u % 2 == 0) or (u >= nch and (u == n_use - 1 or (u - nch) % 2 == 0)) else 2
            e = t[h, g, u, :6]
            e = [int(x - t0) if x else None for x in e]
            print(f"  {name(u):10s} | w{1 if w == 0 else 3} " + " ".join(f"{x:7d}" if x is not None else "      -" for x in iss) +
                  f" | g{g - 1} " + " ".join(f"{x:7d}" if x is not None else "      -" for x in e))
        print(f"  next input published at {t[h + 1, 1, 0, 7] - t0}")


if __name__ == "__main__":
    main()
