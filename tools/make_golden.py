#!/usr/bin/env python
"""Mint the golden fixtures in tests/golden/ from the UNMODIFIED reference.

Run in the development container (needs /root/reference to build
oracle/_ref/libstorm_ref.so):

    python tools/make_golden.py

The reference ships no golden vectors (every input is
std::random_device seeded and no totals are published), so the fixtures are
minted here: seeded inputs from the repo's portable generator
(oracle/storm_oracle.c, orc_gen_row_positions) are pushed through every
reference entry point on the path, and the values are committed.  For each case
the file records

  exact      -- STORM_wrapper_diag with the reference's own per-pair kernel
                (storm.c:132-150 + libalgebra.h:3094-3140), cross-checked against
                a numpy Gram matrix and the column-count closed form;
  ref        -- what each reference struct-API entry point returned;
  ref_defect -- which known reference defect explains a
                `ref` value that differs from `exact` (D1: STORM_t bitmap x list
                probe; D2: contiguous list path beyond 512 rows / 16384 positions; D11:
                contiguous list path with adjacent duplicates in a sparse row);
  pairs_sha256 -- digest of the (N,N) uint32 strict-upper-triangle count matrix
                computed with the reference kernel pair by pair.

tests/test_oracle.py re-derives all of it on CPU from the oracle alone (the GPU
box has no /root/reference) and tests/test_parity_gpu.py checks the CUDA path
against the same numbers.
"""
from __future__ import annotations

import hashlib
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "golden_v1.json")


def ref_pair_matrix(ref: O.Reference, vals: np.ndarray) -> np.ndarray:
    n = vals.shape[0]
    out = np.zeros((n, n), dtype=np.uint32)
    for i in range(n):
        for j in range(i + 1, n):
            out[i, j] = ref.pair_count(vals[i], vals[j])
    return out


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def in_child(fn):
    """Run fn() in a forked child and return its JSON-able result, or None if the
    child died (the reference has undefined behaviour on some inputs, D2/D9)."""
    r, w = os.pipe()
    pid = os.fork()
    if pid == 0:
        os.close(r)
        try:
            os.write(w, json.dumps(fn()).encode())
        finally:
            os._exit(0)
    os.close(w)
    buf = b""
    while True:
        chunk = os.read(r, 1 << 16)
        if not chunk:
            break
        buf += chunk
    os.close(r)
    _, status = os.waitpid(pid, 0)
    if status != 0 or not buf:
        return None
    return json.loads(buf)


def ref_struct_api(ref, M, rows, bsize, vals, skip_contig=False):
    refv = {}
    with O.RefContig(ref, M) as rc, O.RefStorm(ref) as rs:
        for p in rows:
            rc.add(p)
            rs.add(p)
        refv["contig_rows"] = rc.n_rows()
        refv["storm_rows"] = rs.n_rows()
        refv["scalar_cutoff"] = rc.scalar_cutoff()
        refv["contig"] = None if skip_contig else rc.pairw()
        refv["contig_blocked"] = None if skip_contig else rc.pairw_blocked(bsize)
        refv["contig_list"] = None if skip_contig else rc.pairw_list()
        refv["contig_blocked_list"] = None if skip_contig else rc.pairw_blocked_list(bsize)
        refv["storm"] = rs.pairw()
        refv["storm_blocked_auto"] = rs.pairw_blocked(0)
        refv["storm_serialized_size"] = rs.serialized_size()
        refv["wrapper_diag_blocked"] = ref.wrapper_diag_blocked(vals, bsize) if len(rows) else 0
    return refv


def run_case(orc, ref, name, M, rows, bsize, want_pairs):
    """rows: list of uint32 position arrays (sorted unique; may be empty)."""
    vals = O.positions_to_dense(rows, M)
    exact = ref.wrapper_diag(vals) if len(rows) else 0
    checks = {"numpy": O.numpy_total(vals) if len(rows) * M <= 40_000_000 else None,
              "colcount": orc.colcount_total(vals) if len(rows) else 0,
              "oracle": orc.wrapper_diag(vals) if len(rows) else 0}
    for k, v in checks.items():
        assert v is None or v == exact, (name, k, v, exact)

    refv = in_child(lambda: ref_struct_api(ref, M, rows, bsize, vals))
    crashed = refv is None
    if crashed:
        # the reference's contiguous list path read through a corrupted position
        # arena (defect D2 is undefined behaviour) and took the process down
        refv = in_child(lambda: ref_struct_api(ref, M, rows, bsize, vals, skip_contig=True))
        assert refv is not None, name
    cutoff = refv["scalar_cutoff"]

    # explain every divergence with a known defect, or fail
    n_unique = [len(np.unique(p)) for p in rows if len(p)]
    any_sparse = any(u < cutoff for u in n_unique)
    stored = sum(u for u in n_unique if u < cutoff)
    d2_possible = any_sparse and (len(n_unique) > 512 or stored + max(n_unique, default=0) >= 16384)
    # D11 (found while minting): storm.c:1119-1129 writes a sparse row's position
    # list at the ORIGINAL index of each value, so adjacent duplicates leave
    # uninitialised holes while only the first n_unique entries are consumed.
    d11_possible = any(len(p) != len(np.unique(p)) and len(np.unique(p)) < cutoff for p in rows)
    kinds = set()
    for p in rows:
        if len(p):
            ids, cnt = np.unique(np.asarray(p) // 65536, return_counts=True)
            kinds |= {"bitmap" if c >= 4096 else "list" for c in cnt}
    d1_possible = kinds == {"bitmap", "list"}
    defects = {}
    for k in ("contig", "contig_blocked", "contig_list", "contig_blocked_list"):
        if refv[k] is None:
            assert d2_possible or d11_possible, (name, k, "crash without D2/D11 precondition")
            defects[k] = "D2-crash" if d2_possible else "D11-crash"
        elif refv[k] != exact:
            assert d2_possible or d11_possible, (name, k, refv[k], exact)
            defects[k] = "D2" if d2_possible else "D11"
    for k in ("storm", "storm_blocked_auto"):
        if refv[k] != exact:
            assert d1_possible, (name, k, refv[k], exact)
            defects[k] = "D1"
    assert refv["wrapper_diag_blocked"] == exact

    # the restatement must agree with the reference, defect D1 included
    with O.OracleContig(orc, M) as oc, O.OracleStorm(orc) as os_:
        for p in rows:
            oc.add(p)
            os_.add(p)
        assert oc.pairw() == oc.pairw_blocked(bsize) == oc.pairw_list() == oc.pairw_blocked_list(bsize) == exact, name
        assert os_.pairw(False) == os_.pairw_blocked(0, False) == exact, name
        assert os_.pairw(True) == refv["storm"], (name, "D1 emulation", os_.pairw(True), refv["storm"])
        assert os_.pairw_blocked(0, True) == refv["storm_blocked_auto"], name
        assert os_.serialized_size() == refv["storm_serialized_size"], name

    case = {"name": name, "M": M, "N": len(rows), "bsize": bsize, "exact": exact,
            "ref": refv, "ref_defect": defects}
    if want_pairs:
        pm = ref_pair_matrix(ref, vals)
        assert int(pm.sum(dtype=np.uint64)) == exact
        assert (pm == orc.rect_counts(vals, 0, len(rows), 0, len(rows))).all()
        case["pairs_sha256"] = sha(pm)
        case["pairs_row_sums_head"] = [int(x) for x in pm.sum(axis=1, dtype=np.uint64)[:8]]
    return case


def main():
    O.build()
    orc, ref = O.Oracle(), O.Reference()
    cases = []

    # --- seeded uniform cases (benchmark.cpp:749-797 recipe) ----------------
    seeded = [
        # name, M, N, n_draws, seed, bsize, pairs?
        ("ci_4092x1000_d2046", 4092, 1000, 2046, 7, 62, False),      # .travis.yml:199 shape
        ("ci_4096x100_d2048", 4096, 100, 2048, 8, 50, True),         # appveyor.yml:36 shape
        ("c1s_65536x300_d32768", 65536, 300, 32768, 42, 31, True),
        ("c1s_65536x300_d6553", 65536, 300, 6553, 42, 31, True),
        ("c1s_65536x300_d655", 65536, 300, 655, 42, 31, False),
        ("c1s_65536x300_d262", 65536, 300, 262, 42, 31, False),
        ("c1s_65536x300_d199_D2", 65536, 300, 199, 42, 31, False),
        ("c1s_65536x300_d65_D2", 65536, 300, 65, 42, 31, False),
        ("c1s_65536x300_d5", 65536, 300, 5, 42, 31, False),
        ("c1s_65536x300_d1", 65536, 300, 1, 42, 31, False),
        ("c1s_65536x1000_d16384", 65536, 1000, 16384, 1, 31, False),
        ("odd_1000x257_d128", 1000, 257, 128, 3, 7, True),            # M % 64 != 0, N = 2 tiles + 1
        ("odd_200x513_d40", 200, 513, 40, 4, 5, False),               # W = 4 (avx2-lookup tier)
        ("odd_100x70_d30", 100, 70, 30, 5, 3, True),                  # W = 2 (scalar tier), cutoff 0
        ("c2s_524288x120_d262144", 524288, 120, 262144, 11, 5, False),  # all bitmap blocks
        ("c2s_524288x120_d20971", 524288, 120, 20971, 11, 5, False),    # all list blocks (worst tier)
        ("c2s_524288x120_d524", 524288, 120, 524, 11, 5, False),
        ("c4s_1048576x64_d10486", 1048576, 64, 10486, 12, 5, False),    # C4-like 1 %
    ]
    for name, M, N, d, seed, bsize, pairs in seeded:
        rows = [orc.gen_row_positions(seed, i, d, M) for i in range(N)]
        c = run_case(orc, ref, name, M, rows, bsize, pairs)
        c["gen"] = {"kind": "uniform", "seed": seed, "n_draws": d}
        cases.append(c)
        print(name, c["exact"], c["ref_defect"])

    # --- mixed per-row density: exercises bitmap x list dispatch (D1) -------
    mixed = [
        ("mix_65536x200_1000_30000", 65536, 200, 1000, 30000, 21),
        ("mix_524288x150_1_262144", 524288, 150, 1, 262144, 22),
        ("mix_524288x100_40000_262144", 524288, 100, 40000, 262144, 23),
        ("mix_524288x150_1_20000", 524288, 150, 1, 20000, 24),
    ]
    for name, M, N, lo, hi, seed in mixed:
        draws = []
        for i in range(N):
            u = (orc.lib.orc_splitmix64(seed * 7919 + i) >> 11) / float(1 << 53)
            draws.append(max(1, int(math.exp(math.log(lo) + (math.log(hi) - math.log(lo)) * u))))
        rows = [orc.gen_row_positions(seed, i, draws[i], M) for i in range(N)]
        c = run_case(orc, ref, name, M, rows, 5, False)
        c["gen"] = {"kind": "per_row_draws", "seed": seed, "draws": draws}
        cases.append(c)
        print(name, c["exact"], c["ref"]["storm"], c["ref_defect"])

    # --- explicit edge cases (positions stored verbatim) ---------------------
    explicit = [
        ("edge_single_row", 256, [[1, 5, 9]]),
        ("edge_two_rows_disjoint", 256, [[0, 1, 2], [3, 4, 5]]),
        ("edge_two_rows_equal", 256, [[0, 63, 64, 255], [0, 63, 64, 255]]),
        ("edge_adjacent_dups", 65536, [[1, 1, 2, 2, 2, 70000 % 65536], [1, 2, 3, 3, 4464]]),
        ("edge_last_bit", 130, [[129], [0, 129], [128, 129], [64, 65, 129]]),
        ("edge_full_rows", 128, [list(range(128)), list(range(128)), list(range(0, 128, 2))]),
        ("edge_empty_row_middle", 300, [[1, 2, 3], [], [2, 3, 4], [], [3]]),
        ("edge_block_boundaries", 262144, [[65535, 65536, 131071, 131072, 262143],
                                           [0, 65535, 65536, 196608, 262143],
                                           [65536, 131072, 196607, 196608]]),
    ]
    for name, M, rows in explicit:
        rows = [np.asarray(sorted(p), dtype=np.uint32) for p in rows]
        c = run_case(orc, ref, name, M, rows, 5, True)
        c["gen"] = {"kind": "explicit", "rows": [[int(v) for v in p] for p in rows]}
        cases.append(c)
        print(name, c["exact"], c["ref"]["contig_rows"], c["ref"]["storm_rows"])

    # --- u16 list intersections (storm.c:4-73) -------------------------------
    u16_cases = []
    rng_rows = [(5, 9, 31), (8, 8, 32), (17, 300, 33), (4095, 4095, 34), (1, 4000, 35), (64, 64, 36), (0, 10, 37)]
    for n1, n2, seed in rng_rows:
        a = np.unique(orc.gen_row_positions(seed, 0, n1, 65536)).astype(np.uint16)
        b = np.unique(orc.gen_row_positions(seed, 1, n2, 65536)).astype(np.uint16)
        got = ref.intersect_u16(a, b) if len(a) and len(b) else 0
        assert got == orc.intersect_u16(a, b) == len(np.intersect1d(a, b))
        u16_cases.append({"seed": seed, "n1": n1, "n2": n2, "len1": int(a.size), "len2": int(b.size), "count": got})
    # leading-zero special case (storm.c:18-36)
    a = np.array([0, 1, 2, 3, 4, 5, 6, 7, 9, 11, 13, 15, 17, 19, 21, 23], dtype=np.uint16)
    b = np.array([0, 2, 4, 6, 8, 10, 12, 14, 15, 16, 17, 18, 19, 20, 21, 22], dtype=np.uint16)
    got = ref.intersect_u16(a, b)
    assert got == orc.intersect_u16(a, b)
    u16_cases.append({"explicit_a": a.tolist(), "explicit_b": b.tolist(), "count": got})

    doc = {
        "about": "Golden values minted from the unmodified reference (StormBitmaps @ 2eae567, libalgebra @ bff182e) "
                 "by tools/make_golden.py; inputs come from orc_gen_row_positions (oracle/storm_oracle.c).",
        "reference_kernel_w1024": ref.kernel_name(1024),
        "generator_probe": {"splitmix64(0)": int(orc.lib.orc_splitmix64(0)),
                            "draw(42,0,0,65536)": int(orc.lib.orc_draw_position(42, 0, 0, 65536)),
                            "draw(42,7,3,1048576)": int(orc.lib.orc_draw_position(42, 7, 3, 1048576)),
                            "geno_thr(1,0)": int(orc.lib.orc_geno_threshold(1, 0)),
                            "geno_thr(1,12345)": int(orc.lib.orc_geno_threshold(1, 12345)),
                            "geno_row0_sha256": sha(orc.gen_dense_geno(1, 4, 4096))},
        "cases": cases,
        "u16": u16_cases,
    }
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(doc, f, indent=1)
    print("wrote", OUT, len(cases), "cases")


if __name__ == "__main__":
    main()
