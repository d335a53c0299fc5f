#!/usr/bin/env python
"""Ingest paths of the contiguous model: STORM_contig_add row by row against
STORM_b200_contig_add_bulk, on C1-shaped rows; JSON lines (run on the GPU box).

    python tools/ingest_bench.py [rows:bits:draws ...]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import stormbitmaps_b200 as sb
from oracle import oracle as O

cases = [tuple(int(x) for x in a.split(":")) for a in sys.argv[1:]] or [(10000, 65536, 32768), (10000, 65536, 655), (4000, 1048576, 10486)]
orc = O.Oracle()
for rows, bits, draws in cases:
    pos = [orc.gen_row_positions(1, i, draws, bits) for i in range(rows)]
    off = np.zeros(rows + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(p) for p in pos])
    flat = np.concatenate(pos).astype(np.uint32)
    rec = {"rows": rows, "bits": bits, "draws": draws, "positions": int(flat.size)}
    with sb.StormContiguous(bits) as c:                      # warm-up: device state, arenas
        c.add_bulk(flat[:int(off[8])], off[:9])
    with sb.StormContiguous(bits) as a:
        t0 = time.perf_counter()
        for p in pos:
            a.add(p)
        rec["add_rows_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        ta = a.pairw_intersect_cardinality()
        rec["first_query_after_add_s"] = time.perf_counter() - t0           # includes the lazy upload
    with sb.StormContiguous(bits) as b:
        t0 = time.perf_counter()
        b.add_bulk(flat, off)
        rec["add_bulk_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        tb = b.pairw_intersect_cardinality()
        rec["first_query_after_bulk_s"] = time.perf_counter() - t0
    rec["totals_agree"] = ta == tb
    rec["bulk_positions_per_s"] = flat.size / rec["add_bulk_s"]
    print(json.dumps(rec), flush=True)
