#!/usr/bin/env python
"""Per-translation-unit SASS opcode counts of the built library (cuobjdump -sass on bc_mpc_b200/_obj/*.o): the mnemonics that
prove a Blackwell-native kernel (UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk, UTCBAR =
tcgen05.commit, STAS = st.async into a peer CTA's shared memory, SYNCS = mbarrier) and the legacy ones that must be absent
(HMMA = mma.sync).  Runs without a GPU:  python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UBLKCP", "UTMALDG", "STAS", "SYNCS", "UCGABAR", "HMMA", "FFMA", "MUFU"]


def main():
    print("# SASS opcode counts per translation unit (sum over its kernel instantiations); see tools/sass_summary.py")
    print(f"{'object':22s} {'kernels':>7s} " + " ".join(f"{w:>10s}" for w in WATCH))
    for obj in sorted(glob.glob(os.path.join(ROOT, "bc_mpc_b200", "_obj", "*.o"))):
        out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        n_k = len(re.findall(r"^\s*Function :", out, re.M))
        ops = collections.Counter()
        for m in re.finditer(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", out, re.M):
            ops[m.group(1)] += 1
        row = [sum(v for k, v in ops.items() if k.startswith(w)) for w in WATCH]
        print(f"{os.path.basename(obj):22s} {n_k:7d} " + " ".join(f"{v:10d}" for v in row))


if __name__ == "__main__":
    main()
