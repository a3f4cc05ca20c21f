#!/usr/bin/env python
"""Per-source-line warp-stall summary of an ncu report captured with --import-source on (kernels built with -lineinfo):
  python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]
Prints the top lines by stall samples with their dominant stall reasons (read here, without a GPU)."""
import csv
import io
import subprocess
import sys


def main():
    rep, top_n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    path, hdr, lines = None, None, {}
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            path = r[1].split("/")[-1]
        elif len(r) > 10 and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0].isdigit():
            i_s = hdr.index("# Samples")
            stalls = {h: int(v) for h, v in zip(hdr, r) if h.startswith("stall_") and "Not Issued" not in h and v.isdigit() and int(v)}
            lines[(path, int(r[0]))] = (int(r[i_s]), r[1].strip(), stalls)
    tot = sum(v[0] for v in lines.values())
    print(f"total samples {tot}")
    for (f, ln), (n, src, st) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top_n]:
        top = ", ".join(f"{k[6:]} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"{100 * n / tot:5.1f}%  {f}:{ln:<5d} {src[:100]}   [{top}]")


if __name__ == "__main__":
    main()
