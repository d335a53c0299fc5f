"""CPU tests: the oracle (oracle/storm_oracle.c) against the golden vectors that
tools/make_golden.py minted from the unmodified reference, plus the oracle's own
internal consistency (three independent formulations of the same total).

None of these touch a GPU or /root/reference.
"""
import hashlib

import numpy as np
import pytest

from conftest import case_rows
from oracle import oracle as O


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_generator_is_pinned(orc, golden):
    """The portable generator must reproduce the probe values recorded at mint time."""
    p = golden["generator_probe"]
    assert int(orc.lib.orc_splitmix64(0)) == p["splitmix64(0)"]
    assert int(orc.lib.orc_draw_position(42, 0, 0, 65536)) == p["draw(42,0,0,65536)"]
    assert int(orc.lib.orc_draw_position(42, 7, 3, 1048576)) == p["draw(42,7,3,1048576)"]
    assert int(orc.lib.orc_geno_threshold(1, 0)) == p["geno_thr(1,0)"]
    assert int(orc.lib.orc_geno_threshold(1, 12345)) == p["geno_thr(1,12345)"]
    assert _sha(orc.gen_dense_geno(1, 4, 4096)) == p["geno_row0_sha256"]


def _case_ids(golden_path=None):
    import json, os
    with open(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.json")) as f:
        return [c["name"] for c in json.load(f)["cases"]]


@pytest.mark.parametrize("name", _case_ids())
def test_oracle_matches_reference_golden(orc, golden, name):
    case = next(c for c in golden["cases"] if c["name"] == name)
    M, exact, ref = case["M"], case["exact"], case["ref"]
    rows = case_rows(orc, case)
    vals = O.positions_to_dense(rows, M)
    n = len(rows)

    # raw-buffer loop, closed form (independent), numpy Gram (independent, small only)
    assert orc.wrapper_diag(vals) == exact
    assert orc.colcount_total(vals) == exact
    if n * M <= 8_000_000:
        assert O.numpy_total(vals) == exact
    assert ref["wrapper_diag_blocked"] == exact

    # contiguous container: every entry point is exact in the oracle; the
    # reference agrees wherever its list-path defects (D2/D11) cannot trigger
    with O.OracleContig(orc, M) as oc:
        for p in rows:
            oc.add(p)
        got = {"contig": oc.pairw(), "contig_blocked": oc.pairw_blocked(case["bsize"]),
               "contig_list": oc.pairw_list(), "contig_blocked_list": oc.pairw_blocked_list(case["bsize"])}
    for k, v in got.items():
        assert v == exact, k
        if k not in case["ref_defect"]:
            assert ref[k] == exact, k

    # STORM_t container: exact, and bug-compatible with the reference under D1
    with O.OracleStorm(orc) as s:
        for p in rows:
            s.add(p)
        assert s.pairw(False) == exact
        assert s.pairw_blocked(0, False) == exact
        assert s.pairw_blocked(7, False) == exact
        assert s.pairw(True) == ref["storm"]
        assert s.pairw_blocked(0, True) == ref["storm_blocked_auto"]
        assert s.serialized_size() == ref["storm_serialized_size"]
    if "storm" not in case["ref_defect"]:
        assert ref["storm"] == exact

    if "pairs_sha256" in case:
        pm = orc.rect_counts(vals, 0, n, 0, n)
        assert _sha(pm) == case["pairs_sha256"]
        assert [int(x) for x in pm.sum(axis=1, dtype=np.uint64)[:8]] == case["pairs_row_sums_head"]
        assert int(pm.sum(dtype=np.uint64)) == exact


def test_u16_intersection_golden(orc, golden):
    for c in golden["u16"]:
        if "explicit_a" in c:
            a, b = np.array(c["explicit_a"], np.uint16), np.array(c["explicit_b"], np.uint16)
        else:
            a = np.unique(orc.gen_row_positions(c["seed"], 0, c["n1"], 65536)).astype(np.uint16)
            b = np.unique(orc.gen_row_positions(c["seed"], 1, c["n2"], 65536)).astype(np.uint16)
            assert (a.size, b.size) == (c["len1"], c["len2"])
        assert orc.intersect_u16(a, b) == c["count"] == len(np.intersect1d(a, b))


def test_rectangles_tile_the_triangle(orc):
    """Tile sums (what each GPU / CTA produces) add up to the total; off-diagonal
    tiles equal the column-count product form."""
    vals = orc.gen_dense_uniform(9, 200, 3000, 8192)
    total = orc.wrapper_diag(vals)
    T = 64
    acc = 0
    for i0 in range(0, 200, T):
        for j0 in range(i0, 200, T):
            i1, j1 = min(i0 + T, 200), min(j0 + T, 200)
            t = orc.rect_total(vals, i0, i1, j0, j1)
            if j0 > i0:
                assert t == orc.colcount_rect(vals, i0, i1, j0, j1)
            assert t == int(orc.rect_counts(vals, i0, i1, j0, j1).sum(dtype=np.uint64))
            acc += t
    assert acc == total


def test_square_and_geno_generators(orc):
    a = orc.gen_dense_geno(3, 40, 5000)
    b = orc.gen_dense_geno(3, 24, 5000, row0=40)
    ab = np.concatenate([a, b])
    assert (orc.gen_dense_geno(3, 64, 5000) == ab).all()          # row0 offset is consistent
    sq = orc.wrapper_square(a, b)
    assert sq == orc.rect_total(ab, 0, 40, 40, 64) == orc.colcount_rect(ab, 0, 40, 40, 64)
    dens = np.unpackbits(a.view(np.uint8), axis=1).mean(axis=1)
    assert dens.min() >= 0.002 and dens.max() <= 0.53              # p in [0.005, 0.5]
    # padding bits beyond M stay clear
    assert (a[:, -1] >> np.uint64(5000 % 64)).max() == 0


def test_contig_error_conventions(orc):
    """storm.c:1032-1034,1136,1139-1141."""
    L = orc.lib
    assert L.orc_contig_add(None, None, 0) == -1
    with O.OracleContig(orc, 1000) as c:
        assert L.orc_contig_add(c.h, None, 3) == -2
        assert c.add([]) == 0                       # no row appended (D7)
        assert c.clear() == 0                       # nothing allocated yet
        assert c.add([1, 2, 3]) == 3
        assert c.add([2, 2, 3]) == 3                # returns n_values, not the unique count
        assert c.pairw() == 2
        assert c.clear() == 1
        assert c.pairw() == 0
    assert L.orc_contig_pairw(None) == 2**64 - 1
    assert L.orc_storm_pairw(None, 0) == 2**64 - 1
