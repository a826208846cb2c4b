#!/usr/bin/env python
"""Developer script: split an ncu SASS source page of sweep_kernel into phases at the barrier instructions and
report, per phase, stall samples, warp-instructions executed, shared-memory wavefronts and the stall mix.
usage: python tools/ncu_sass_phases.py rep.ncu-rep"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
phase, phases = 0, []
cur = dict(name="start", samples=0, inst=0, wave=0, ideal=0, stall={s: 0 for s in stalls}, first=None, ops={})
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[col["Source"]].strip()
    s, i = int(r[col["# Samples"]]), int(r[col["Instructions Executed"]])
    w, wi = int(r[col["L1 Wavefronts Shared"]] or 0), int(r[col["L1 Wavefronts Shared Ideal"]] or 0)
    cur["samples"] += s
    cur["inst"] += i
    cur["wave"] += w
    cur["ideal"] += wi
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    cur["ops"][op] = cur["ops"].get(op, 0) + i
    for st in stalls:
        cur["stall"][st] += int(r[col[st]] or 0)
    if src.startswith("BAR") or " BAR." in src or src.startswith("EXIT"):
        cur["end"] = src
        phases.append(cur)
        cur = dict(name=f"after {src[:40]}", samples=0, inst=0, wave=0, ideal=0, stall={s: 0 for s in stalls}, ops={})
phases.append(cur)
tot = sum(p["samples"] for p in phases)
for p in phases:
    if p["samples"] < 20 and p["inst"] < 1000:
        continue
    top = sorted(p["stall"].items(), key=lambda kv: -kv[1])[:5]
    ops = sorted(p["ops"].items(), key=lambda kv: -kv[1])[:6]
    print(f"{p['name'][:44]:44s} samples {p['samples']:6d} ({100.0 * p['samples'] / tot:4.1f}%) inst {p['inst']:10d} smem wavefronts {p['wave']:9d} (ideal {p['ideal']:9d})")
    print("      stalls: " + ", ".join(f"{k[6:]} {v}" for k, v in top) + "   ops: " + ", ".join(f"{k} {v}" for k, v in ops))
