#!/usr/bin/env python
"""Top stalled SASS instructions of a kernel in an .ncu-rep: python tools/ncu_source_top.py rep [n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") or h.startswith("Stall")]
data = []
for idx, r in enumerate(rows[2:]):
    if len(r) < len(hdr): continue
    try: n = int(r[ci["# Samples"]])
    except ValueError: continue
    data.append((idx, n, r))
tot = sum(n for _, n, _ in data)
print("instructions", len(data), "samples", tot)
for idx, n, r in sorted(data, key=lambda x: -x[1])[:topn]:
    ex = r[ci["Instructions Executed"]]
    print(f"{idx:5d} {n:7d} {100*n/tot:5.1f}%  exec={ex:>10s}  {r[ci['Source']].strip()[:100]}")
