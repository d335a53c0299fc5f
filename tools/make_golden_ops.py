#!/usr/bin/env python
"""Mint tests/golden/golden_ops_v1.json: union / diff totals of every golden_v1 case, computed by the
UNMODIFIED reference (STORM_wrapper_diag, storm.c:132-150, driven with the kernels its own choosers
STORM_get_union_count_func / STORM_get_diff_count_func return, libalgebra.h:3142-3236) through
oracle/_ref/libstorm_ref.so, plus a SHA-256 of the per-pair matrices from the reference kernels.

Run where /root/reference exists (`make -C oracle ref` first); the output is committed."""
import hashlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O
from conftest import case_rows

O.build()
orc, ref = O.Oracle(), O.Reference()
golden = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_v1.json")))
u64p = O.u64p
out = {"about": "union (op 1) and diff (op 2) totals per golden_v1 case from the unmodified reference: "
                "STORM_wrapper_diag(n, vals, W, STORM_get_{union,diff}_count_func(W)); pairs_sha256 = SHA-256 of the "
                "row-major uint32 strict-upper per-pair matrix from the same reference kernels (cases with N <= 300)",
       "cases": []}
for case in golden["cases"]:
    rows = case_rows(orc, case)
    if not rows:
        continue
    vals = O.positions_to_dense(rows, case["M"])
    n, w = vals.shape
    rec = {"name": case["name"]}
    for op, name in ((1, "union"), (2, "diff")):
        rec[name] = ref.wrapper_diag_op(vals, op)
        assert rec[name] == orc.wrapper_diag_op(vals, op), (case["name"], name)
        if n <= 300:
            m = np.zeros((n, n), dtype=np.uint32)
            for i in range(n):
                for j in range(i + 1, n):
                    m[i, j] = ref.lib.REF_count_op(vals[i].ctypes.data_as(u64p), vals[j].ctypes.data_as(u64p), w, op)
            assert (m == orc.rect_counts_op(vals, 0, n, 0, n, op)).all()
            rec[name + "_pairs_sha256"] = hashlib.sha256(m.tobytes()).hexdigest()
    out["cases"].append(rec)
    print(rec["name"], rec["union"], rec["diff"], flush=True)
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "golden_ops_v1.json"), "w"), indent=1)
