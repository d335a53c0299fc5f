#!/usr/bin/env python
"""Is tcgen05.mma kind::mxf4 exact for bit counting, and how fast is it?  (run on the GPU box)

Prints one JSON object: the issue-rate ceilings of kind::i8 and kind::mxf4 and, per probe case of
fp4_probe.cu, expected / smallest / largest accumulator and the number of wrong accumulators."""
import ctypes as C, json, os, struct, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import stormbitmaps_b200 as sb
from stormbitmaps_b200 import _lib

lib = sb.load()
out = {"device": sb.device_info(0)}
for kind, name in [(5, "umma_i8_cta_group2"), (6, "umma_mxf4_cta_group1"), (7, "umma_mxf4_cta_group2")]:
    rate, _ = sb.microbench(kind)
    out[name] = {"ops_per_s": rate, "tops": rate / 1e12}
cases = []
for pattern in range(4):
    for n_full in (0, 1, 2, 255, 256, 2048, 2049, 16384, 131072, 262143):
        for n_single in (0, 1, 3, 63):
            if n_full == 0 and n_single == 0:
                continue
            if 64 * n_full + n_single >= 2 ** 24:
                continue
            cases.append((n_full, n_single, pattern))
arr = np.asarray(cases, dtype=np.uint32)
res = np.zeros((len(cases), 4), dtype=np.float32)
_lib.check(lib.STORM_b200_fp4_probe(arr.ctypes.data_as(_lib.u32p), len(cases), res.ctypes.data_as(C.POINTER(C.c_float))), "fp4 probe")
bad = []
rows = []
for (n_full, n_single, pattern), r in zip(cases, res):
    mism = int(np.frombuffer(r[3].tobytes(), dtype=np.uint32)[0])
    row = {"n_full": n_full, "n_single": n_single, "pattern": pattern, "expect": float(r[0]), "min": float(r[1]), "max": float(r[2]), "mismatches": mism}
    rows.append(row)
    if mism:
        bad.append(row)
out["cases"] = len(cases)
out["inexact_cases"] = bad
out["all_exact"] = not bad
out["largest_checked"] = max(r["expect"] for r in rows)
print(json.dumps(out, indent=1))
