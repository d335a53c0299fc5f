"""GPU parity tests (run with ``-m gpu`` on a B200).  Everything here calls the
CUDA path through the C-ABI of libstorm_b200.so and compares it, bit for bit,
with the CPU oracle and the golden values minted from the unmodified reference.

Nothing here reads /root/reference; the oracle is used only as the checker.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import case_rows
from oracle import oracle as O

pytestmark = pytest.mark.gpu

KERNELS = os.environ.get("STORM_TEST_KERNELS", "popc,csa,umma,fp4,b1").split(",")


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def sb():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    import stormbitmaps_b200 as sb
    sb.load()
    if os.environ.get("STORM_UMMA_CG"):
        sb.set_umma_cta_group(int(os.environ["STORM_UMMA_CG"]))
    info = sb.device_info(0)
    assert info["cc"] // 10 == 10, f"libstorm_b200 targets sm_100a, got {info}"
    return sb


def _case_names():
    with open(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.json")) as f:
        return [c["name"] for c in json.load(f)["cases"]]


# --------------------------------------------------------------------------- #
# golden vectors through the reference-facing API
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("name", _case_names())
def test_golden_through_storm_h_api(sb, orc, golden, name):
    case = next(c for c in golden["cases"] if c["name"] == name)
    M, exact = case["M"], case["exact"]
    rows = case_rows(orc, case)
    with sb.StormContiguous(M) as c, sb.Storm() as s:
        for p in rows:
            c.add(p)
            s.add(p)
        # dense model: all four entry points return the exact value (the reference's
        # list path diverges under D2/D11 -- recorded in the fixture, not reproduced)
        assert c.pairw_intersect_cardinality() == exact
        assert c.pairw_intersect_cardinality_blocked(case["bsize"]) == exact
        for route in ("auto", "tile", "probe", "stream"):                 # every kernel mix behind the *_list calls
            prev = sb.set_contig_list_route(route)
            try:
                assert c.pairw_intersect_cardinality_list() == exact, route
                assert c.pairw_intersect_cardinality_blocked_list(case["bsize"]) == exact, route
            finally:
                sb.set_contig_list_route(prev)
        # sparse model: exact (D1 not reproduced), any bsize
        assert s.pairw_intersect_cardinality() == exact
        assert s.pairw_intersect_cardinality_blocked(0) == exact
        assert s.serialized_size() == case["ref"]["storm_serialized_size"]
        # shards add up
        assert sum(c.pairw_shard(k, 3) for k in range(3)) == exact
        assert sum(s.pairw_shard(k, 3) for k in range(3)) == exact
        # wherever the reference itself is defect-free its own return value is matched too
        for k in ("contig", "contig_blocked", "contig_list", "contig_blocked_list", "storm", "storm_blocked_auto"):
            if k not in case["ref_defect"]:
                assert case["ref"][k] == exact
        n = len([p for p in rows if len(p)])          # contig skips empty rows (D7)
        if "pairs_sha256" in case and n == len(rows) and n >= 1:
            pm = c.pairw_rect(0, n, 0, n)
            assert _sha(pm) == case["pairs_sha256"]
            assert _sha(s.pairw_rect(0, n, 0, n)) == case["pairs_sha256"]
    vals = O.positions_to_dense(rows, M)
    if len(rows):
        assert sb.wrapper_diag(vals) == exact


# --------------------------------------------------------------------------- #
# device-resident rows: every kernel variant against the oracle
# --------------------------------------------------------------------------- #
def _device_rows(sb, vals):
    import torch
    n, w = vals.shape
    rows, _ = sb.alloc_rows(n, w * 64)
    rows[:, :w] = torch.from_numpy(vals.view(np.int64)).cuda()
    return rows


SHAPES = [
    # (M, N, n_draws, seed)
    (4096, 300, 1500, 101),       # W = 64: one UMMA k-block exactly
    (65536, 257, 20000, 102),     # two tiles + one row
    (8192, 1000, 4000, 103),
    (1000, 130, 300, 104),        # M % 64 != 0
    (192, 50, 60, 105),           # W = 3: ragged K tail
    (128, 600, 64, 106),          # W = 2: minimum UMMA width
    (64, 40, 20, 107),            # W = 1
]


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("M,N,draws,seed", SHAPES)
def test_device_total_and_pairs_match_oracle(sb, orc, kernel, M, N, draws, seed):
    import torch
    vals = orc.gen_dense_uniform(seed, N, draws, M)
    W = vals.shape[1]
    if kernel in ("umma", "fp4") and W < 2:
        pytest.skip("UMMA needs at least 128 bits per row")
    rows = _device_rows(sb, vals)
    exact = orc.wrapper_diag(vals)
    total = sb.pairw_device(rows, n_words=W, kernel=kernel)
    torch.cuda.synchronize()
    assert int(total.item()) == exact
    # sharded: partial sums add up and no shard is empty when there are enough tiles
    parts = [int(sb.pairw_device(rows, n_words=W, shard=k, n_shards=4, kernel=kernel).item()) for k in range(4)]
    assert sum(parts) == exact
    # per-pair counts, full matrix and an off-diagonal / ragged rectangle
    counts, t2 = sb.pairw_rect_device(rows, 0, N, 0, N, n_words=W, kernel=kernel)
    want = orc.rect_counts(vals, 0, N, 0, N)
    assert (counts.cpu().numpy().view(np.uint32) == want).all()
    assert int(t2.item()) == exact
    i0, i1, j0, j1 = N // 3, N // 3 + min(70, N // 2), N // 5, N - 1
    counts, t3 = sb.pairw_rect_device(rows, i0, i1, j0, j1, n_words=W, kernel=kernel)
    want = orc.rect_counts(vals, i0, i1, j0, j1)
    assert (counts.cpu().numpy().view(np.uint32) == want).all()
    assert int(t3.item()) == int(want.sum(dtype=np.uint64)) == orc.rect_total(vals, i0, i1, j0, j1)
    # all pairs (no triangle mask) == XY^T square
    a, b = vals[: N // 2], vals[N // 2:]
    _, sq = sb.square_device(rows[: N // 2], rows[N // 2:], n_words=W, kernel=kernel)
    assert int(sq.item()) == orc.wrapper_square(a, b)


@pytest.mark.parametrize("kernel", ["umma", "fp4"])
def test_per_pair_output_vector_and_scalar_stores(sb, orc, kernel):
    """Per-pair rectangles whose chunks take the 32-byte store path (ld a multiple of 8, whole 32-column chunks
    right of the diagonal), the scalar path (diagonal / ragged chunks, odd ld), and both inside one tile."""
    M, N = 2048 + 64, 1100
    vals = orc.gen_dense_uniform(77, N, 700, M)
    W = vals.shape[1]
    rows = _device_rows(sb, vals)
    for (i0, i1, j0, j1, strict) in [(0, 300, 320, 960, True), (0, 512, 0, 512, True), (0, 520, 0, 1000, True),
                                     (10, 300, 37, 360, True), (600, 1100, 0, 1096, False), (256, 700, 250, 1100, True)]:
        counts, total = sb.pairw_rect_device(rows, i0, i1, j0, j1, n_words=W, kernel=kernel, strict_upper=strict)
        want = orc.rect_counts(vals, i0, i1, j0, j1)
        if not strict:
            want = np.array([[orc.pair_count(vals[i], vals[j]) for j in range(j0, j1)] for i in range(i0, i0 + 3)], dtype=np.uint32)
            assert (counts[:3].cpu().numpy().view(np.uint32) == want).all(), (i0, i1, j0, j1)
            assert int(total.item()) == int(counts.cpu().numpy().view(np.uint32).sum(dtype=np.uint64))
            continue
        assert (counts.cpu().numpy().view(np.uint32) == want).all(), (i0, i1, j0, j1)
        assert int(total.item()) == int(want.sum(dtype=np.uint64))


@pytest.mark.parametrize("kernel", KERNELS)
def test_extreme_rows(sb, orc, kernel):
    """All-ones, all-zero and single-bit rows; counts reach M exactly."""
    import torch
    M, N = 2048, 140
    W = M // 64
    vals = np.zeros((N, W), dtype=np.uint64)
    vals[0:40] = np.uint64(2**64 - 1)
    vals[60, 0] = np.uint64(1)
    vals[61, W - 1] = np.uint64(1) << np.uint64(63)
    vals[62:100:2] = np.uint64(0xAAAAAAAAAAAAAAAA)
    rows = _device_rows(sb, vals)
    counts, total = sb.pairw_rect_device(rows, 0, N, 0, N, n_words=W, kernel=kernel)
    want = orc.rect_counts(vals, 0, N, 0, N)
    assert want.max() == M
    assert (counts.cpu().numpy().view(np.uint32) == want).all()
    assert int(total.item()) == orc.wrapper_diag(vals) == O.numpy_total(vals)


def test_synthetic_generators_match_oracle(sb, orc):
    import torch
    for M, N, draws, seed, row0 in [(65536, 64, 30000, 5, 0), (1000, 33, 200, 6, 1000), (131072, 16, 7, 7, 123456)]:
        rows, W = sb.alloc_rows(N, M)
        sb.synth_uniform_device(rows, M, draws, seed, row0)
        torch.cuda.synchronize()
        got = rows.cpu().numpy().view(np.uint64)
        assert (got[:, :W] == orc.gen_dense_uniform(seed, N, draws, M, row0=row0)).all()
        assert (got[:, W:] == 0).all()
    for M, N, seed, row0 in [(4096, 40, 1, 0), (5000, 17, 3, 99), (131072, 8, 2, 199990)]:
        rows, W = sb.alloc_rows(N, M)
        sb.synth_geno_device(rows, M, seed, row0)
        torch.cuda.synchronize()
        got = rows.cpu().numpy().view(np.uint64)
        assert (got[:, :W] == orc.gen_dense_geno(seed, N, M, row0=row0)).all()


def test_bulk_ingest_equals_row_by_row(sb, orc):
    M, N = 65536, 700            # crosses the 512-row growth boundary of the reference (D2 territory)
    rows = [orc.gen_row_positions(77, i, [3, 150, 4000, 0, 30000][i % 5], M) for i in range(N)]
    offs = np.zeros(N + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(r) for r in rows])
    flat = np.concatenate(rows).astype(np.uint32)
    vals = O.positions_to_dense([r for r in rows if len(r)], M)
    exact = orc.wrapper_diag(vals)
    with sb.StormContiguous(M) as a, sb.StormContiguous(M) as b:
        for r in rows:
            a.add(r)
        b.add_bulk(flat, offs)
        for c in (a, b):
            assert c.pairw_intersect_cardinality() == exact
            assert c.pairw_intersect_cardinality_list() == exact          # sparse rows -> probe kernel
            assert c.pairw_intersect_cardinality_blocked_list(31) == exact
        n = vals.shape[0]
        assert (a.pairw_rect(0, n, 0, n) == b.pairw_rect(0, n, 0, n)).all()
        # clear keeps capacity; reuse with a different density (benchmark.cpp:738-739)
        a.clear()
        for r in rows[:100]:
            a.add(r)
        assert a.pairw_intersect_cardinality() == orc.wrapper_diag(O.positions_to_dense([r for r in rows[:100] if len(r)], M))


def test_contig_list_routes_agree(sb, orc):
    """storm.c:1253-1258 switches per pair between the probe and the bitmap kernel; here that is a cost decision
    between three kernel mixes.  All give the exact total on all-sparse, mixed and all-dense containers, shards add
    up, and the stream kernel is only taken when every row is a list."""
    M = 65536
    for name, draws in (("all_sparse", [1, 5, 60, 150, 199, 3]), ("mixed", [5, 150, 4000, 30000, 90, 250]), ("all_dense", [300, 5000, 40000])):
        rows = [orc.gen_row_positions(31, i, draws[i % len(draws)], M) for i in range(700)]
        exact = orc.wrapper_diag(O.positions_to_dense(rows, M))
        with sb.StormContiguous(M) as c:
            for r in rows:
                c.add(r)
            assert c.pairw_intersect_cardinality() == exact
            for route in ("auto", "tile", "probe", "stream"):
                prev = sb.set_contig_list_route(route)
                try:
                    assert c.pairw_intersect_cardinality_list() == exact, (name, route)
                    took = c.last_list_route()
                    if name == "all_dense":
                        assert took == "tile"
                    elif route == "stream":
                        assert took == ("stream" if name == "all_sparse" else "probe"), (name, took)
                    elif route != "auto":
                        assert took == route, (name, route, took)
                    assert c.pairw_intersect_cardinality_blocked_list(9) == exact, (name, route)
                finally:
                    sb.set_contig_list_route(prev)


def test_storm_t_dispatch_regimes(sb, orc):
    """One container holding tiny rows, list rows and bitmap rows: every branch of
    the per-block dispatch (storm.c:618-656) is exercised against the exact oracle."""
    M = 3 * 65536 + 1000
    draws = [1, 5, 40, 64, 65, 300, 3000, 9000, 20000, 60000, 150000, 0]
    rows = [orc.gen_row_positions(55, i, draws[i % len(draws)], M) for i in range(180)]
    with O.OracleStorm(orc) as ref_s, sb.Storm() as s:
        for p in rows:
            ref_s.add(p)
            s.add(p)
        exact = ref_s.pairw(False)
        assert s.pairw_intersect_cardinality() == exact
        assert s.pairw_intersect_cardinality_blocked(7) == exact
        vals = O.positions_to_dense(rows, M)
        assert (s.pairw_rect(0, 180, 0, 180) == orc.rect_counts(vals, 0, 180, 0, 180)).all()
        assert (s.pairw_rect(20, 90, 50, 171) == orc.rect_counts(vals, 20, 90, 50, 171)).all()
        with sb.Storm() as t:
            for p in rows[:50]:
                t.add(p)
            assert s.intersect_cardinality_square(t) == orc.wrapper_square(vals, vals[:50])
        s.clear()
        assert s.pairw_intersect_cardinality() == 0
        for p in rows[100:160]:
            s.add(p)
        assert s.pairw_intersect_cardinality() == orc.wrapper_diag(vals[100:160])


@pytest.mark.parametrize("route", ["sparse", "dense", "auto"])
def test_storm_t_routes_agree(sb, orc, route):
    """Whole-container STORM_t queries: the sparse merge/probe kernel and the densified +
    dense-tile route return the same exact total on every density regime of config C2
    (scaled to 120 rows) and on the C4-like 1 % rows."""
    M = 524288
    prev = sb.set_storm_route(route)
    try:
        for draws in (262144, 20971, 5242, 524, 104, 5, 1):
            rows = [orc.gen_row_positions(77, i, draws, M) for i in range(120)]
            with O.OracleStorm(orc) as ref_s, sb.Storm() as s:
                for p in rows:
                    ref_s.add(p)
                    s.add(p)
                exact = ref_s.pairw(False)
                assert s.pairw_intersect_cardinality_blocked(0) == exact, (route, draws)
                took = s.last_route()
                if route != "auto":
                    assert took == route
                else:                                  # cost model: dense for well-filled rows, merge/probe for
                    if draws >= 20971:                 # nearly empty ones; in between it depends on the row count
                        assert took == "dense", (draws, took)
                    elif draws <= 5:
                        assert took == "sparse", (draws, took)
                    else:
                        assert took in ("dense", "sparse")
                # shards add up on either route
                assert sum(s.pairw_shard(r, 3) for r in range(3)) == exact
                # a second query reuses the resident mirror; a mutation invalidates it
                assert s.pairw_intersect_cardinality() == exact
                s.add(rows[0])
                ref_s.add(rows[0])
                assert s.pairw_intersect_cardinality() == ref_s.pairw(False)
    finally:
        sb.set_storm_route(prev)


@pytest.mark.parametrize("M", [524288, 40 * 65536 + 77])
def test_storm_t_split_route(sb, orc, M):
    """Containers that hold a few heavy rows (bitmap blocks, or more values than a row group of the stream kernel takes)
    among light ones: the split route -- light rows among themselves through the stream kernel, every pair with a heavy
    row through the block merge/probe kernel with the heavy rows as rows i -- gives the oracle's total, alone, in
    shards and after a mutation, on every route.  Heavy rows first, last, adjacent, and with bitmap blocks built
    from duplicates (few bits)."""
    draws = [3, 40, 700, 0, 1, 2500, 64, 300]
    rows = [orc.gen_row_positions(123, i, draws[i % len(draws)], M) for i in range(300)]
    heavy = {0: 150000, 1: 9000, 77: 60000, 78: 12000, 150: 200000, 299: 30000}
    for r, d in heavy.items():
        rows[r] = orc.gen_row_positions(124, r, d, M)
    rows[200] = np.concatenate([np.full(5000, 9, dtype=np.uint32), np.array([10, 65536 + 4], dtype=np.uint32)])   # bitmap block, 2 bits
    vals = O.positions_to_dense(rows, M)
    exact = orc.wrapper_diag(vals)
    for route in ("split", "auto", "sparse", "dense"):
        prev = sb.set_storm_route(route)
        try:
            with sb.Storm() as s:
                for p in rows:
                    s.add(p)
                assert s.pairw_intersect_cardinality() == exact, (M, route)
                if route != "auto":
                    assert s.last_route() == route, (route, s.last_route())
                assert sum(s.pairw_shard(r, 5) for r in range(5)) == exact, (M, route)
                s.add(rows[150])                                  # one more heavy row: both halves of the mirror rebuilt
                s.add(rows[2])
                more = np.concatenate([vals, vals[150:151], vals[2:3]])
                assert s.pairw_intersect_cardinality() == orc.wrapper_diag(more), (M, route)
        finally:
            sb.set_storm_route(prev)
    # only heavy rows / only light rows: the split route does not apply and the sparse kernels answer
    prev = sb.set_storm_route("split")
    try:
        for sel in ([rows[r] for r in heavy], rows[2:60]):
            with sb.Storm() as s:
                for p in sel:
                    s.add(p)
                assert s.pairw_intersect_cardinality() == orc.wrapper_diag(O.positions_to_dense(sel, M))
                assert s.last_route() == "sparse"
    finally:
        sb.set_storm_route(prev)


def test_storm_t_stream_kernel_position_ranges(sb, orc):
    """Rows of a few hundred to a few thousand values: the light mirror is cut into position ranges (whole blocks) so that
    32 rows fit the stream kernel's table range by range.  Even rows, rows whose values sit in ONE block (a group then
    holds few rows), empty rows and rows that only touch the last range; totals, shards, mutation."""
    M = 16 * 65536
    rng = np.random.default_rng(5)
    rows = [orc.gen_row_positions(131, i, [900, 2500, 40, 1500, 0, 3000][i % 6], M) for i in range(400)]
    for r in range(7, 400, 13):                                   # all values inside block 3 (3 800 draws: below the bitmap limit)
        rows[r] = (3 * 65536 + np.unique(rng.integers(0, 65536, 3800))).astype(np.uint32)
    for r in range(11, 400, 17):                                  # only the last range
        rows[r] = (15 * 65536 + np.unique(rng.integers(0, 65536, 700))).astype(np.uint32)
    vals = O.positions_to_dense(rows, M)
    exact = orc.wrapper_diag(vals)
    prev = sb.set_storm_route("dense")
    try:
        with sb.Storm() as s:
            for p in rows:
                s.add(p)
            assert s.pairw_intersect_cardinality() == exact           # mirror built for the densified route: no light arrays yet
            assert s.last_route() == "dense"
            sb.set_storm_route("sparse")                              # ... they are added when a route first reads them
            assert s.pairw_intersect_cardinality() == exact
            assert s.last_route() == "sparse"
            assert sum(s.pairw_shard(k, 7) for k in range(7)) == exact
            assert (s.pairw_rect(0, 400, 0, 400) == orc.rect_counts(vals, 0, 400, 0, 400)).all()   # (row-major flat form, untouched)
            s.add(rows[7])
            assert s.pairw_intersect_cardinality() == orc.wrapper_diag(np.concatenate([vals, vals[7:8]]))
        # values that do not ascend INSIDE a block (blocks still in order): the range-major mirror does not rely on it
        with sb.Storm() as s:
            for p in rows:
                q = p.copy()
                for blk in np.unique(q >> 16):
                    sel = np.flatnonzero((q >> 16) == blk)
                    q[sel] = rng.permutation(q[sel])
                s.add(q)
            assert s.pairw_intersect_cardinality() == exact
    finally:
        sb.set_storm_route(prev)


@pytest.mark.parametrize("M", [2 * 65536, 1048576, 20 * 65536, 21 * 65536 + 5])
def test_storm_t_flat_probe_kernel_equals_block_kernel(sb, orc, M):
    """Sparse route on containers without bitmap blocks: the row-group stream kernel (totals), the flat probe kernel
    (whole-row shared bitmap, partner positions probed; rows up to 20 blocks wide) and the block merge/probe kernel
    give the oracle's total, shard sums, per-pair rectangles and XY^T totals; empty rows, single values, duplicates
    and last-bit values included."""
    draws = [0, 1, 3, 17, 64, 65, 200, 1000, 3500, 2, 0, 700]
    rows = [orc.gen_row_positions(91, i, draws[i % len(draws)], M) for i in range(150)]
    rows[5] = np.array([M - 1], dtype=np.uint32)
    rows[6] = np.array([0, 0, 7, 7, 65535, 65536, M - 1, M - 1], dtype=np.uint32)      # adjacent duplicates collapse
    vals = O.positions_to_dense(rows, M)
    exact = orc.wrapper_diag(vals)
    prev = sb.set_storm_route("sparse")
    try:
        for flat in ("stream", "flat", "block"):
            was = sb.set_sparse_flat(flat)
            try:
                with sb.Storm() as s, sb.Storm() as t:
                    for p in rows:
                        s.add(p)
                    for p in rows[40:95]:
                        t.add(p)
                    assert s.pairw_intersect_cardinality() == exact, (M, flat)
                    assert s.last_route() == "sparse"
                    assert sum(s.pairw_shard(r, 4) for r in range(4)) == exact, (M, flat)
                    assert (s.pairw_rect(0, 150, 0, 150) == orc.rect_counts(vals, 0, 150, 0, 150)).all(), (M, flat)
                    assert (s.pairw_rect(3, 77, 30, 149) == orc.rect_counts(vals, 3, 77, 30, 149)).all(), (M, flat)
                    assert s.intersect_cardinality_square(t) == orc.wrapper_square(vals, vals[40:95]), (M, flat)
                    s.add(rows[8])                                   # a mutation rebuilds both mirrors
                    assert s.pairw_intersect_cardinality() == orc.wrapper_diag(np.concatenate([vals, vals[8:9]])), (M, flat)
            finally:
                sb.set_sparse_flat(was)
    finally:
        sb.set_storm_route(prev)


# --------------------------------------------------------------------------- #
# full-size configurations: size-independent properties
# --------------------------------------------------------------------------- #
def _colcount_total_torch(rows, W):
    """sum_k C(c_k, 2) with torch ops only (independent of the library's kernels)."""
    import torch
    total = 0
    for b in range(64):
        c = ((rows[:, :W] >> b) & 1).sum(dim=0, dtype=torch.int64)
        total += int((c * (c - 1) // 2).sum().item())
    return total


def test_c1_full_size_total_and_sampled_tiles(sb, orc):
    """BASELINE config C1: benchmark 65536 10000, 32768 draws per row."""
    import torch
    M, N, draws, seed = 65536, 10000, 32768, 1
    rows, W = sb.alloc_rows(N, M)
    sb.synth_uniform_device(rows, M, draws, seed)
    closed = _colcount_total_torch(rows, W)
    for kernel in KERNELS:
        total = sb.pairw_device(rows, n_words=W, kernel=kernel)
        assert int(total.item()) == closed, kernel
        parts = [int(sb.pairw_device(rows, n_words=W, shard=k, n_shards=8, kernel=kernel).item()) for k in range(8)]
        assert sum(parts) == closed, kernel
        # shards hold equal numbers of TILES (the unit of kernel time); their pair totals differ where a shard collects the
        # diagonal tiles and the ragged last column block (16 of 256 rows on C1), so this is a loose check only
        assert min(parts) > 0.65 * max(parts), "shards are balanced"
    # sampled tiles (diagonal, interior, last ragged) against the reference kernel restated in the oracle
    host = rows[:, :W].cpu().numpy().view(np.uint64)
    for (i0, i1, j0, j1) in [(0, 40, 0, 40), (4990, 5030, 9960, 10000), (9970, 10000, 9970, 10000), (100, 130, 7000, 7040)]:
        want = orc.rect_counts(host, i0, i1, j0, j1)
        for kernel in KERNELS:
            got, _ = sb.pairw_rect_device(rows, i0, i1, j0, j1, n_words=W, kernel=kernel)
            assert (got.cpu().numpy().view(np.uint32) == want).all(), (kernel, i0, j0)


def test_c3_shape_subsample_and_idempotence(sb, orc):
    """C3 row width (131072 bits, genotype-like) on a row subsample: closed form,
    oracle on a sub-block, and identical results on repeated / re-ordered queries."""
    import torch
    M, N, seed = 131072, 6000, 2
    rows, W = sb.alloc_rows(N, M)
    sb.synth_geno_device(rows, M, seed)
    closed = _colcount_total_torch(rows, W)
    res = {k: int(sb.pairw_device(rows, n_words=W, kernel=k).item()) for k in KERNELS}
    assert all(v == closed for v in res.values()), res
    assert int(sb.pairw_device(rows, n_words=W).item()) == closed            # AUTO, repeated
    perm = torch.randperm(N, device="cuda")
    assert int(sb.pairw_device(rows[perm].contiguous(), n_words=W).item()) == closed   # row order is irrelevant
    host = orc.gen_dense_geno(seed, 96, M, row0=3000)
    got, t = sb.pairw_rect_device(rows, 3000, 3096, 3000, 3096, n_words=W)
    assert (got.cpu().numpy().view(np.uint32) == orc.rect_counts(host, 0, 96, 0, 96)).all()
    assert int(t.item()) == orc.wrapper_diag(host)


# --------------------------------------------------------------------------- #
# FP4 tensor form (tcgen05.mma kind::mxf4, fp32 accumulators): exactness limits
# --------------------------------------------------------------------------- #
def test_fp4_accumulation_selftest_and_probe(sb):
    """The hardware property the FP4 form rests on, measured on this device: fp32 accumulators driven
    to 2^24 - 1 by +64 and +1 steps stay exact for every operand encoding the kernel uses."""
    import ctypes as C
    from stormbitmaps_b200 import _lib
    lib = sb.load()
    assert lib.STORM_b200_fp4_selftest() == 1
    cases = np.asarray([(n_full, n_single, p) for p in range(4)
                        for (n_full, n_single) in [(0, 1), (1, 0), (3, 5), (4096, 63), (131071, 63), (262143, 63)]],
                       dtype=np.uint32)
    res = np.zeros((len(cases), 4), dtype=np.float32)
    _lib.check(lib.STORM_b200_fp4_probe(cases.ctypes.data_as(_lib.u32p), len(cases),
                                        res.ctypes.data_as(C.POINTER(C.c_float))), "fp4 probe")
    for (n_full, n_single, p), r in zip(cases, res):
        want = 64.0 * n_full + n_single
        assert r[0] == want and r[1] == want and r[2] == want, (n_full, n_single, p, r[:3])
        assert r[3:4].view(np.uint32)[0] == 0


def test_fp4_largest_counts_and_limit(sb, orc):
    """Counts of 2^22 (all-ones rows of 4 Mi bits) come out exact; beyond 2^24 bits per row the FP4
    form refuses and AUTO answers with the int8 form."""
    import torch
    N, M = 260, 1 << 22
    rows, W = sb.alloc_rows(N, M)
    rows[:, :W] = -1
    rows[7, 5] = 0x0123456789ABCDEF                  # one row that is not all ones
    host = rows[:, :W].cpu().numpy().view(np.uint64)
    exact = orc.wrapper_diag(host)
    for kernel in ("fp4", "umma", "auto"):
        assert int(sb.pairw_device(rows, n_words=W, kernel=kernel).item()) == exact, kernel
    got, _ = sb.pairw_rect_device(rows, 0, 16, 0, 16, n_words=W, kernel="fp4")
    assert (got.cpu().numpy().view(np.uint32) == orc.rect_counts(host, 0, 16, 0, 16)).all()
    del rows
    N, M = 3, (1 << 24) + 64
    rows, W = sb.alloc_rows(N, M)
    rows[:, :W] = -1
    assert sb.resolved_kernel_name("auto", W) == "umma"
    assert sb.resolved_kernel_name("auto", 2048) == "fp4"
    with pytest.raises(sb.StormError):
        sb.pairw_device(rows, n_words=W, kernel="fp4")
    assert int(sb.pairw_device(rows, n_words=W, kernel="auto").item()) == 3 * M


@pytest.mark.parametrize("N,M", [(300, 8192), (700, 65536 + 64), (2000, 4096), (5000, 8192), (1100, 1 << 17)])
def test_stream_k_split_is_exact(sb, N, M):
    """Total-only queries split (tile, K chunk) units over the persistent CTAs: fewer tiles than SMs (K slices
    of a tile on different CTAs), a ragged last chunk, and a tail wave, against whole-tile scheduling."""
    rows, W = sb.alloc_rows(N, M)
    sb.synth_uniform_device(rows, M, max(1, M // 3), 21)
    closed = _colcount_total_torch(rows, W)
    for on in (True, False):
        prev = sb.set_umma_stream_k(on)
        try:
            for kernel in ("umma", "fp4"):
                assert int(sb.pairw_device(rows, n_words=W, kernel=kernel).item()) == closed, (on, kernel)
            parts = [int(sb.pairw_device(rows, n_words=W, shard=r, n_shards=3, kernel="fp4").item()) for r in range(3)]
            assert sum(parts) == closed, (on, parts)
        finally:
            sb.set_umma_stream_k(bool(prev))


@pytest.mark.parametrize("N,M", [(2000, 4096), (5000, 8192), (3000, 1024), (4100, 192), (1500, 65536), (900, 1 << 19)])
def test_accumulator_chaining_is_exact(sb, orc, N, M):
    """Total-only queries drain the tensor-memory accumulator once per run of interior tiles (up to
    2^24 / (32 M) tiles in the FP4 form): same total as one drain per tile, for both tensor forms, with
    diagonal and ragged edge tiles (which end a run) in between, through shards, and with all-ones rows
    (every accumulator element at the top of its range)."""
    rows, W = sb.alloc_rows(N, M)
    sb.synth_uniform_device(rows, M, max(1, M // 2), 33)
    closed = _colcount_total_torch(rows, W)
    if N * W <= 3_000_000:
        host = rows[:, :W].cpu().numpy().view(np.uint64)
        assert orc.wrapper_diag(np.ascontiguousarray(host[:600])) == int(sb.pairw_device(rows[:600], n_words=W).item())
    for on in (True, False):
        prev = sb.set_umma_chain(on)
        try:
            for kernel in ("umma", "fp4"):
                assert int(sb.pairw_device(rows, n_words=W, kernel=kernel).item()) == closed, (on, kernel)
            parts = [int(sb.pairw_device(rows, n_words=W, shard=r, n_shards=3, kernel="fp4").item()) for r in range(3)]
            assert sum(parts) == closed, (on, parts)
        finally:
            sb.set_umma_chain(bool(prev))
    rows[:, :W] = -1                                     # every pair count = 64 W: the accumulator's worst case
    full = N * (N - 1) // 2 * 64 * W
    for kernel in ("umma", "fp4"):
        assert int(sb.pairw_device(rows, n_words=W, kernel=kernel).item()) == full, kernel


def test_wave_sync_is_only_a_hint(sb):
    """The wave counter of the persistent tensor kernels changes when CTAs load, never what they compute."""
    N, M = 5000, 8192
    rows, W = sb.alloc_rows(N, M)
    sb.synth_uniform_device(rows, M, 3000, 9)
    closed = _colcount_total_torch(rows, W)
    for on in (False, True):
        prev = sb.set_umma_wave_sync(on)
        try:
            for kernel in ("umma", "fp4"):
                assert int(sb.pairw_device(rows, n_words=W, kernel=kernel).item()) == closed, (on, kernel)
        finally:
            sb.set_umma_wave_sync(bool(prev))


@pytest.mark.parametrize("kernel", ["fp4", "umma", "popc"])
def test_tile_ranges_and_reserved_sms_add_up(sb, orc, kernel):
    """STORM_b200_pairw_tiles_device over any partition of the raster = the full total; the band plan of the
    pipelined multi-GPU host query (distributed.stream_plan) is such a partition, and a band's tiles give the
    right answer with only the rows of bands <= b resident (later rows poisoned with ones).  Leaving SMs to a
    collective (STORM_b200_set_umma_reserved_sms) never changes a result."""
    import torch
    from stormbitmaps_b200 import distributed as D
    N, M = 4500, 4096
    vals = orc.gen_dense_uniform(33, N, 1500, M)
    exact = orc.wrapper_diag(vals)
    rows, W = _device_rows(sb, vals), vals.shape[1]
    kid = sb.resolve_kernel(kernel, W)
    n_tiles = sb.tile_count(N, kernel)[0]
    cuts = sorted({0, 1, n_tiles // 3, n_tiles // 3 + 1, n_tiles - 1, n_tiles})
    parts = [int(sb.pairw_tiles_device(rows, a, b, n_words=W, kernel=kid).item()) for a, b in zip(cuts, cuts[1:])]
    assert sum(parts) == exact
    with pytest.raises(sb.StormError):
        sb.pairw_tiles_device(rows, 0, n_tiles + 1, n_words=W, kernel=kid)
    for world in (2, 3):
        plan = D.stream_plan(N, world, kid, 4)
        got = 0
        for rsv in (0, 2, 200):
            prev = sb.set_umma_reserved_sms(rsv)
            try:
                got = 0
                staged = torch.full_like(rows, -1)                     # rows that have "not arrived yet": all ones
                for (r0, r1, t0, t1) in plan:
                    staged[r0:r1] = rows[r0:r1]
                    for r in range(world):
                        tb, te = D.rank_tiles(t0, t1, r, world)
                        got += int(sb.pairw_tiles_device(staged, tb, te, n_words=W, kernel=kid).item())
                assert got == exact, (world, rsv)
            finally:
                sb.set_umma_reserved_sms(prev)


def test_c_caller_linked_with_the_library_prints_the_reference_totals(sb):
    """tests/drivers/dropin_driver.c -- benchmark.cpp's call sequence in plain C99 -- linked with
    libstorm_b200.so prints, field for field, what it prints when linked with the reference's storm.c
    (tests/golden/dropin_driver_v1.json): both models, blocked variants, the raw-buffer wrapper, clear + reuse."""
    import json, tempfile
    from test_abi import build_dropin_driver, run_dropin_driver, HOST_FIELDS, QUERY_FIELDS
    golden = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dropin_driver_v1.json")))["cases"]
    with tempfile.TemporaryDirectory() as d:
        exe = build_dropin_driver(d)
        for args, want in golden.items():
            got = run_dropin_driver(exe, args)
            for k in HOST_FIELDS + QUERY_FIELDS:
                assert got[k] == want[k], (args, k, got[k], want[k])


# --------------------------------------------------------------------------- #
# round 2: FP4 exactness at the top of the range, through the production kernel
# --------------------------------------------------------------------------- #
def test_fp4_random_increment_probe_both_cta_groups(sb):
    """Data-dependent increments (0 .. 64 per instruction and element, production nibble encoding) at cta_group 1
    and 2, driven to 2^24 - 64: every one of the 128 (256) x 256 accumulators must hold the exact integer."""
    import ctypes as C
    lib = sb.load()
    for cg in (1, 2):
        for steps, seed in [(1, 1), (7, 2), (4097, 3), (100000, 4), (262143, 5), (262143, 6)]:
            res = np.zeros(4, dtype=np.float32)
            rc = lib.STORM_b200_fp4_probe_random(cg, steps, seed, res.ctypes.data_as(C.POINTER(C.c_float)))
            assert rc == 0, sb.last_error()
            assert res[0] == 64.0 * steps, (cg, steps, res)
            assert res[1] == 0.0 and res[2] == 0.0 and res[3:4].view(np.uint32)[0] == 0, (cg, steps, res)


def _random_density_rows(N, W, seed):
    """Rows of W words on the GPU with per-row densities ~25 / 50 / 75 / 94 / 3 % (torch generators; the oracle
    gets a host copy), rows 0 and 1 all ones."""
    import torch
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    def rnd(n):
        return torch.randint(-2**63, 2**63 - 1, (n, W), dtype=torch.int64, device="cuda", generator=g)
    rows = torch.empty((N, (W + 15) // 16 * 16), dtype=torch.int64, device="cuda")
    rows.zero_()
    k = N // 5
    rows[:k, :W] = rnd(k) & rnd(k)
    rows[k:2 * k, :W] = rnd(k)
    rows[2 * k:3 * k, :W] = rnd(k) | rnd(k)
    rows[3 * k:4 * k, :W] = rnd(k) | rnd(k) | rnd(k) | rnd(k)
    n5 = N - 4 * k
    rows[4 * k:, :W] = rnd(n5) & rnd(n5) & rnd(n5) & rnd(n5) & rnd(n5)
    rows[0, :W] = -1
    rows[1, :W] = -1
    return rows


@pytest.mark.parametrize("N,M", [(300, 1 << 23), (272, (1 << 24) - 64)])
def test_fp4_production_kernel_at_the_top_of_its_range(sb, orc, N, M):
    """dense_umma_kernel<2, FP4> (the AUTO default) on random-density rows of 2^23 and 2^24 - 64 bits: totals, shard
    sums and every per-pair count against the oracle.  Pair counts run up to M itself (the all-ones pair), i.e. the
    fp32 accumulators are driven by data-dependent increments to the top of their exact range."""
    import torch
    W = M // 64
    rows = _random_density_rows(N, W, 1234 + N)
    host = np.ascontiguousarray(rows[:, :W].cpu().numpy().view(np.uint64))
    want = orc.rect_counts(host, 0, N, 0, N)
    assert int(want.max()) == M
    exact = int(want.sum(dtype=np.uint64))
    assert sb.resolved_kernel_name("auto", W) == "fp4"
    for kernel in ("fp4", "auto", "umma"):
        assert int(sb.pairw_device(rows, n_words=W, kernel=kernel).item()) == exact, kernel
    assert sum(int(sb.pairw_device(rows, n_words=W, shard=k, n_shards=3, kernel="fp4").item()) for k in range(3)) == exact
    counts, total = sb.pairw_rect_device(rows, 0, N, 0, N, n_words=W, kernel="fp4")
    assert (counts.cpu().numpy().view(np.uint32) == want).all()
    assert int(total.item()) == exact


def test_fp4_chained_accumulators_at_the_top_of_their_range(sb, orc):
    """Total-only queries keep several interior tiles in one accumulator (chain_max = 2^24 / (32 M)): with M = 2^18 a
    run is two tiles long, an element reaches 2 M = 2^19 and the epilogue's 32-column float sum 2^24 exactly when the
    rows are all ones.  Random-density rows against the closed form and the oracle on sampled tiles, then all ones."""
    import torch
    N, M = 4700, 1 << 18
    W = M // 64
    rows = _random_density_rows(N, W, 99)
    closed = _colcount_total_torch(rows, W)
    for chain in (True, False):
        prev = sb.set_umma_chain(chain)
        try:
            assert int(sb.pairw_device(rows, n_words=W, kernel="fp4").item()) == closed, chain
            parts = [int(sb.pairw_device(rows, n_words=W, shard=r, n_shards=2, kernel="fp4").item()) for r in range(2)]
            assert sum(parts) == closed, (chain, parts)
        finally:
            sb.set_umma_chain(bool(prev))
    for (i0, i1, j0, j1) in [(0, 24, 0, 24), (900, 930, 4670, 4700), (2800, 2816, 2800, 2830)]:
        sub = np.ascontiguousarray(torch.cat([rows[i0:i1, :W], rows[j0:j1, :W]]).cpu().numpy().view(np.uint64))
        ni, nj = i1 - i0, j1 - j0
        want = orc.rect_counts(sub, 0, ni, ni, ni + nj)
        got, _ = sb.pairw_rect_device(rows, i0, i1, j0, j1, n_words=W, kernel="fp4", strict_upper=False)
        assert (got.cpu().numpy().view(np.uint32) == want).all(), (i0, j0)
    rows[:, :W] = -1
    assert int(sb.pairw_device(rows, n_words=W, kernel="fp4").item()) == N * (N - 1) // 2 * M


# --------------------------------------------------------------------------- #
# round 2: raw-buffer list wrappers, foreign per-pair functions
# --------------------------------------------------------------------------- #
def test_wrapper_diag_list_and_blocked(sb, orc):
    """STORM_wrapper_diag_list[_blocked] (storm.c:173-220, 281-369): host matrix + caller-built position arrays, any
    cutoff / block size; equal to the dense wrapper and to the exact value on sparse, mixed and dense rows."""
    M = 8192
    for draws in ([3, 20, 150, 40], [5, 600, 4000, 90, 2500], [3000, 6000]):
        pos = [orc.gen_row_positions(19, i, draws[i % len(draws)], M) for i in range(333)]
        vals = O.positions_to_dense(pos, M)
        exact = orc.wrapper_diag(vals)
        assert sb.wrapper_diag(vals) == exact
        for cutoff in (0, 40, 200, 10**6):
            assert sb.wrapper_diag_list(vals, pos, cutoff) == exact, (draws, cutoff)
            assert sb.wrapper_diag_list(vals, pos, cutoff, bsize=15) == exact, (draws, cutoff)
    assert sb.wrapper_diag_list(vals[:1], pos[:1], 40) == 0


def test_wrappers_identify_a_foreign_compute_func(sb, orc):
    """A per-pair function that is not this library's own (here: ctypes callbacks standing in for libalgebra's static
    kernels in the caller's translation unit) selects the set operation it computes; an unrecognisable one is an error."""
    import ctypes as C
    lib = sb.load()
    proto = C.CFUNCTYPE(C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_size_t)
    fns = {0: proto(lambda a, b, n: sum(bin(a[k] & b[k]).count("1") for k in range(n))),
           1: proto(lambda a, b, n: sum(bin(a[k] | b[k]).count("1") for k in range(n))),
           2: proto(lambda a, b, n: sum(bin(a[k] ^ b[k]).count("1") for k in range(n)))}
    vals = orc.gen_dense_uniform(5, 90, 700, 4096)
    p = vals.ctypes.data_as(C.POINTER(C.c_uint64))
    for op, f in fns.items():
        assert lib.STORM_wrapper_diag(90, p, 64, C.cast(f, C.c_void_p)) == orc.wrapper_diag_op(vals, op), op
        assert lib.STORM_wrapper_diag_blocked(90, p, 64, C.cast(f, C.c_void_p), 7) == orc.wrapper_diag_op(vals, op), op
    bogus = proto(lambda a, b, n: 7)
    assert lib.STORM_wrapper_diag(90, p, 64, C.cast(bogus, C.c_void_p)) == 2**64 - 1
    assert "neither" in sb.last_error()


# --------------------------------------------------------------------------- #
# round 2: device sets behind storm.h (replicas; several on one device when the box has one GPU)
# --------------------------------------------------------------------------- #
def _device_lists(sb):
    n = sb.load().STORM_b200_device_count()
    lists = [[0, 0], [0, 0, 0]]
    if n >= 2:
        lists += [list(range(n)), [1, 0]]
    return lists


def test_multi_device_contig_queries_equal_single_device(sb, orc):
    """STORM_B200_DEVICES / STORM_b200_set_device_list: every query entry point of the dense model and the raw-buffer
    wrappers return the single-device value when the tile raster is sharded over G replicas (G real GPUs where the
    box has them; replicas on one GPU otherwise -- the same host logic: band slices, peer pulls, per-replica
    shards, host-side sum)."""
    M = 65536
    draws = [5, 150, 4000, 30000, 90, 250, 20000]
    rows = [orc.gen_row_positions(71, i, draws[i % len(draws)], M) for i in range(1500)]
    vals = O.positions_to_dense(rows, M)
    exact = orc.wrapper_diag(vals)
    sparse_rows = [orc.gen_row_positions(72, i, [1, 5, 60, 150, 199, 3][i % 6], M) for i in range(900)]
    sparse_exact = orc.wrapper_diag(O.positions_to_dense(sparse_rows, M))
    for ids in _device_lists(sb):
        sb.set_device_list(ids)
        try:
            assert sb.get_devices() == ids
            with sb.StormContiguous(M) as c:
                for r in rows[:700]:
                    c.add(r)
                assert c.pairw_intersect_cardinality_blocked(31) == orc.wrapper_diag(vals[:700])   # first query: banded upload
                assert c.device_count() == len(ids)
                for r in rows[700:]:
                    c.add(r)                                                                   # rows added after a query
                assert c.pairw_intersect_cardinality() == exact
                assert c.pairw_intersect_cardinality_blocked(5) == exact                           # steady state
                for route in ("auto", "tile", "probe"):
                    prev = sb.set_contig_list_route(route)
                    try:
                        assert c.pairw_intersect_cardinality_list() == exact, (ids, route)
                    finally:
                        sb.set_contig_list_route(prev)
                assert sum(c.pairw_shard(k, 3) for k in range(3)) == exact                       # external shards x replicas
                n = vals.shape[0]
                assert (c.pairw_rect(10, 90, 40, 300) == orc.rect_counts(vals, 10, 90, 40, 300)).all()
                c.clear()
                for r in sparse_rows:
                    c.add(r)
                for route in ("auto", "stream", "probe", "tile"):
                    prev = sb.set_contig_list_route(route)
                    try:
                        assert c.pairw_intersect_cardinality_blocked_list(9) == sparse_exact, (ids, route)
                    finally:
                        sb.set_contig_list_route(prev)
            assert sb.wrapper_diag(vals) == exact
            assert sb.wrapper_diag(vals[:3]) == orc.wrapper_diag(vals[:3])
            big = orc.gen_dense_uniform(3, 5000, 3000, 131072)          # 82 MB: several bands
            assert sb.wrapper_diag(big) == orc.colcount_total(big)
            assert sb.wrapper_diag(vals, op="union") == orc.wrapper_diag_op(vals, 1)
        finally:
            sb.set_device_list(())
    assert len(sb.get_devices()) == 1


def test_device_threads_on_and_off_agree(sb, orc):
    """STORM_b200_set_device_threads: the per-device calls of a multi-device query go out from one host thread per
    device (default) or from the caller alone; same totals either way, for the dense model, the raw-buffer wrapper,
    STORM_t on both routes, and for many short queries back to back (workers polling, then asleep)."""
    import time
    M = 32768
    rows = [orc.gen_row_positions(81, i, [40, 3000, 9000, 7][i % 4], M) for i in range(1300)]
    vals = O.positions_to_dense(rows, M)
    exact = orc.wrapper_diag(vals)
    ids = _device_lists(sb)[-2] if len(_device_lists(sb)) > 2 else [0, 0, 0]
    sb.set_device_list(ids)
    try:
        for on in (True, False, True):
            sb.set_device_threads(on)
            with sb.StormContiguous(M) as c, sb.Storm() as s:
                for r in rows:
                    c.add(r)
                    s.add(r)
                for k in range(40):
                    assert c.pairw_intersect_cardinality() == exact, (on, k)
                    if k == 20:
                        time.sleep(0.05)                                   # the workers have gone to sleep
                assert sb.wrapper_diag(vals) == exact
                for route in ("dense", "sparse"):
                    prev = sb.set_storm_route(route)
                    try:
                        assert s.pairw_intersect_cardinality() == exact, (on, route)
                    finally:
                        sb.set_storm_route(prev)
    finally:
        sb.set_device_threads(True)
        sb.set_device_list(())


def test_bulk_ingest_on_replicas(sb, orc):
    M, N = 65536, 600
    rows = [orc.gen_row_positions(77, i, [3, 150, 4000, 0, 30000][i % 5], M) for i in range(N)]
    offs = np.zeros(N + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(r) for r in rows])
    flat = np.concatenate(rows).astype(np.uint32)
    exact = orc.wrapper_diag(O.positions_to_dense([r for r in rows if len(r)], M))
    sb.set_device_list([0, 0])
    try:
        with sb.StormContiguous(M) as c:
            c.add_bulk(flat[: offs[200]], offs[:201])
            for r in rows[200:300]:
                c.add(r)
            c.add_bulk(flat[offs[300]:], offs[300:] - offs[300])
            assert c.pairw_intersect_cardinality() == exact
            assert c.pairw_intersect_cardinality_list() == exact
    finally:
        sb.set_device_list(())


def test_rows_are_uploaded_while_they_are_added(sb, orc):
    """The host mirror is page-locked and finished rows go to the device in the background (every 8 MB), so the
    first query after row-by-row ingest pays (almost) no upload; results do not depend on where the batches fall,
    on growth of the mirror in between, on clear + reuse, or on a pageable source (the raw-buffer wrapper)."""
    M, N = 131072, 3000                                  # 16 KiB rows: 512 rows per batch, mirror grows 512 -> 4096
    vals = orc.gen_dense_uniform(8, N, 9000, M)
    exact = orc.colcount_total(vals)
    pos = [np.flatnonzero(np.unpackbits(vals[i].view(np.uint8), bitorder="little")).astype(np.uint32) for i in range(N)]
    with sb.StormContiguous(M) as c:
        for p in pos:
            c.add(p)
        assert c.pairw_intersect_cardinality_blocked(15) == exact
        first = c.last_timing()
        assert c.pairw_intersect_cardinality_blocked(15) == exact
        steady = c.last_timing()
        assert first["total_s"] < 20 * steady["total_s"] + 0.05, (first, steady)
        c.clear()
        for p in pos[:700]:
            c.add(p)
        assert c.pairw_intersect_cardinality() == orc.colcount_total(vals[:700])
    assert sb.wrapper_diag(vals) == exact                # numpy memory: pageable, goes through the staging ring


def test_dense_ingest_rehome_and_resident_multi_device_query(sb, orc):
    """STORM_b200_contig_add_dense (rows as bitmaps), STORM_b200_contig_rehome (same container, another device set) and
    STORM_b200_pairw_devices (matrix already resident on every device of a list) against the oracle."""
    import torch
    M, N = 65536, 1300
    vals = orc.gen_dense_uniform(61, N, 20000, M)
    exact = orc.wrapper_diag(vals)
    with sb.StormContiguous(M) as c:
        c.add_dense(vals[:500])
        assert c.pairw_intersect_cardinality() == orc.wrapper_diag(vals[:500])
        c.add_dense(vals[500:])
        assert c.pairw_intersect_cardinality_blocked(31) == exact and c.device_count() == 1
        for ids in _device_lists(sb):
            sb.set_device_list(ids)
            try:
                c.rehome()
                assert c.pairw_intersect_cardinality_blocked(31) == exact and c.device_count() == len(ids)
                assert c.pairw_intersect_cardinality_list() == exact
            finally:
                sb.set_device_list(())
        c.rehome()
        assert c.pairw_intersect_cardinality() == exact and c.device_count() == 1
    n_dev = torch.cuda.device_count()
    copies = []
    for d in range(n_dev):
        rows, W = sb.alloc_rows(N, M, device=f"cuda:{d}")
        rows[:, :W] = torch.from_numpy(vals.view(np.int64)).to(f"cuda:{d}")
        copies.append(rows)
    for k in sorted({1, n_dev, max(1, n_dev // 2)}):
        assert sb.pairw_devices(copies[:k], n_words=vals.shape[1]) == exact, k
    torch.cuda.set_device(0)


def test_storm_t_banded_dense_route_and_densified_rectangles(sb, orc):
    """Containers whose dense form is too large to keep whole are densified in row bands (forced here with small
    bands): triangle of a band + rectangles with the later bands = the exact total, shards included; per-pair
    rectangles of containers with bitmap blocks go through densified row ranges and the tile kernel."""
    M = 3 * 65536 + 1000
    draws = [1, 5, 40, 300, 3000, 9000, 20000, 60000, 150000, 0, 64, 65]
    rows = [orc.gen_row_positions(58, i, draws[i % len(draws)], M) for i in range(700)]
    vals = O.positions_to_dense(rows, M)
    exact = orc.wrapper_diag(vals)
    prev_route = sb.set_storm_route("dense")
    try:
        with sb.Storm() as s:
            for p in rows:
                s.add(p)
            assert s.pairw_intersect_cardinality() == exact
            for band in (256, 300, 512, 10000):
                prev = sb.set_storm_band_rows(band)
                try:
                    assert s.pairw_intersect_cardinality() == exact, band
                    assert s.pairw_intersect_cardinality_blocked(0) == exact, band
                    assert sum(s.pairw_shard(r, 3) for r in range(3)) == exact, band
                finally:
                    sb.set_storm_band_rows(prev)
            assert s.pairw_intersect_cardinality() == exact
    finally:
        sb.set_storm_route(prev_route)
    with sb.Storm() as s:                                   # AUTO: rectangles through densified rows (bitmap blocks present)
        for p in rows:
            s.add(p)
        assert (s.pairw_rect(0, 700, 0, 700) == orc.rect_counts(vals, 0, 700, 0, 700)).all()
        assert (s.pairw_rect(20, 290, 150, 671) == orc.rect_counts(vals, 20, 290, 150, 671)).all()
        assert (s.pairw_rect(500, 700, 3, 130) == orc.rect_counts(vals, 500, 700, 3, 130)).all()


def test_multi_device_storm_t_queries_equal_single_device(sb, orc):
    """Whole-container STORM_t queries on a device set: the block mirror goes to every replica, each answers its
    shard through the route every replica agrees on (sparse kernels, densified rows, densified in bands); totals,
    external shards x replicas, mutation after a query, clear + reuse."""
    M = 4 * 65536
    for name, draws in (("lists", [1, 5, 60, 150, 700, 3, 0]), ("mixed", [5, 150, 4000, 30000, 90, 70000, 250]), ("bitmaps", [20000, 50000, 90000])):
        rows = [orc.gen_row_positions(83, i, draws[i % len(draws)], M) for i in range(420)]
        vals = O.positions_to_dense(rows, M)
        exact = orc.wrapper_diag(vals)
        for ids in _device_lists(sb):
            sb.set_device_list(ids)
            try:
                for route in ("auto", "sparse", "dense"):
                    prev = sb.set_storm_route(route)
                    try:
                        with sb.Storm() as s:
                            for p in rows[:300]:
                                s.add(p)
                            assert s.pairw_intersect_cardinality() == orc.wrapper_diag(vals[:300]), (name, ids, route)
                            for p in rows[300:]:
                                s.add(p)
                            assert s.pairw_intersect_cardinality_blocked(0) == exact, (name, ids, route)
                            assert sum(s.pairw_shard(k, 2) for k in range(2)) == exact, (name, ids, route)
                            if route == "dense":
                                was = sb.set_storm_band_rows(256)
                                try:
                                    assert s.pairw_intersect_cardinality() == exact, (name, ids, "banded")
                                finally:
                                    sb.set_storm_band_rows(was)
                            assert (s.pairw_rect(5, 60, 30, 200) == orc.rect_counts(vals, 5, 60, 30, 200)).all()
                            s.clear()
                            for p in rows[100:250]:
                                s.add(p)
                            assert s.pairw_intersect_cardinality() == orc.wrapper_diag(vals[100:250]), (name, ids, route)
                    finally:
                        sb.set_storm_route(prev)
            finally:
                sb.set_device_list(())


@pytest.mark.parametrize("ids", [(), (0, 0)])
def test_two_host_threads_with_their_own_objects(sb, orc, ids):
    """The reference's threading contract is one host thread per object (storm.c has no locks).  Two threads, each with
    its own containers and its own raw-buffer calls, run concurrently (ctypes releases the GIL) and get exact results:
    the process-wide pieces they share -- prefix cache, wave counters, staging ring, wrapper scratch arenas, FP4
    self-test, per-thread tensor-map cache, and (on a device set: ids) the per-device worker threads -- are locked or
    thread-local."""
    import threading
    M = 65536
    sb.set_device_list(list(ids))
    jobs = []
    for t in range(2):
        rows = [orc.gen_row_positions(200 + t, i, [5, 150, 4000, 30000, 90][i % 5], M) for i in range(900 + 100 * t)]
        vals = O.positions_to_dense(rows, M)
        jobs.append((rows, vals, orc.wrapper_diag(vals)))
    errors = []

    def work(t):
        rows, vals, exact = jobs[t]
        try:
            for rep in range(3):
                with sb.StormContiguous(M) as c, sb.Storm() as s:
                    for p in rows:
                        c.add(p)
                        s.add(p)
                    assert c.pairw_intersect_cardinality_blocked(31) == exact
                    assert c.pairw_intersect_cardinality_list() == exact
                    assert s.pairw_intersect_cardinality() == exact
                    assert (c.pairw_rect(3, 200, 100, 700) == orc.rect_counts(vals, 3, 200, 100, 700)).all()
                assert sb.wrapper_diag(vals) == exact
        except Exception as e:                                # noqa: BLE001
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    sb.set_device_list(())
    assert not errors, errors
