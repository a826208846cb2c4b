#!/usr/bin/env python
"""Per-source-line summary of an ncu source page (`--import-source on` capture), read offline.
usage: python tools/ncu_source_lines.py rep.ncu-rep [min_samples] -> file:line, stall samples, warp-instructions executed"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    min_s = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file, hdr = None, None
    tot_s = tot_i = 0
    agg = []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
            continue
        if hdr is None or r[0] in ("Function Name",) or len(r) < len(hdr):
            continue
        if r[2] != "-":  # SASS row
            continue
        try:
            s, i = int(r[iS]), int(r[iI])
        except ValueError:
            continue
        agg.append((cur_file, int(r[0]), r[1].strip()[:100], s, i))
        tot_s += s
        tot_i += i
    print(f"# total samples {tot_s}, warp-instructions {tot_i}")
    for f, ln, src, s, i in agg:
        if s >= min_s:
            print(f"{f}:{ln:<5d} samples {s:6d} ({100.0 * s / tot_s:5.1f}%)  inst {i:10d} ({100.0 * i / tot_i:5.1f}%)  {src}")


if __name__ == "__main__":
    main()
