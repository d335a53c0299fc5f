"""CPU tests of the drop-in boundary: the library loads without a GPU, exports
every symbol include/*.h declares, keeps the reference's public struct layout,
mirrors its host-side error conventions, and fails loudly (no CPU fallback) when
a query needs a device that is not there."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from stormbitmaps_b200 import build
    build.build()
    import stormbitmaps_b200 as sb
    return sb.load()


def _declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(STORM_[A-Za-z0-9_]+)\s*\(", text)) - {"STORM_ALIGN"}


def test_every_declared_symbol_is_exported_and_bound(lib):
    from stormbitmaps_b200 import _lib
    declared = _declared_functions("storm.h") | _declared_functions("storm_b200.h")
    assert len(declared) > 50
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"
    unbound = sorted(declared - set(_lib.SIGNATURES))
    assert not unbound, f"declared but without a ctypes signature: {unbound}"
    exported = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    assert "orc_" not in exported, "the product library must not contain oracle code"


LAYOUT_C = r"""
#include <stdio.h>
#include <stddef.h>
#include "storm.h"
#define P(T, f) printf(#T "." #f " %zu\n", offsetof(T, f))
int main(void) {
  printf("sizeof.STORM_contiguous_t %zu\n", sizeof(STORM_contiguous_t));
  P(STORM_contiguous_t, data); P(STORM_contiguous_t, scalar); P(STORM_contiguous_t, n_scalar);
  P(STORM_contiguous_t, bitmaps); P(STORM_contiguous_t, n_data); P(STORM_contiguous_t, m_data);
  P(STORM_contiguous_t, tot_scalar); P(STORM_contiguous_t, m_scalar); P(STORM_contiguous_t, vector_length);
  P(STORM_contiguous_t, n_bitmaps_vector); P(STORM_contiguous_t, intsec_func); P(STORM_contiguous_t, alignment);
  P(STORM_contiguous_t, scalar_cutoff); P(STORM_contiguous_t, b200);
  printf("sizeof.STORM_contiguous_bitmap_t %zu\n", sizeof(STORM_contiguous_bitmap_t));
  P(STORM_contiguous_bitmap_t, data); P(STORM_contiguous_bitmap_t, scalar); P(STORM_contiguous_bitmap_t, n_scalar);
  printf("sizeof.STORM_t %zu\n", sizeof(STORM_t));
  P(STORM_t, conts); P(STORM_t, n_conts); P(STORM_t, m_conts); P(STORM_t, b200);
  printf("sizeof.STORM_bitmap_cont_t %zu\n", sizeof(STORM_bitmap_cont_t));
  P(STORM_bitmap_cont_t, bitmaps); P(STORM_bitmap_cont_t, block_ids); P(STORM_bitmap_cont_t, n_bitmaps);
  P(STORM_bitmap_cont_t, m_bitmaps); P(STORM_bitmap_cont_t, prev_inserted_value);
  printf("sizeof.STORM_bitmap_t %zu\n", sizeof(STORM_bitmap_t));
  printf("alignof.STORM_bitmap_t %zu\n", _Alignof(STORM_bitmap_t));
  P(STORM_bitmap_t, data); P(STORM_bitmap_t, scalar); P(STORM_bitmap_t, n_bits_set); P(STORM_bitmap_t, n_missing);
  P(STORM_bitmap_t, m_scalar); P(STORM_bitmap_t, id);
  return 0;
}
"""

# offsets of the reference structs on x86-64 (gcc C99/C11 and g++).
REFERENCE_LAYOUT = {
    "STORM_contiguous_t.data": 0, "STORM_contiguous_t.scalar": 8, "STORM_contiguous_t.n_scalar": 16,
    "STORM_contiguous_t.bitmaps": 24, "STORM_contiguous_t.n_data": 32, "STORM_contiguous_t.m_data": 40,
    "STORM_contiguous_t.tot_scalar": 48, "STORM_contiguous_t.m_scalar": 56, "STORM_contiguous_t.vector_length": 64,
    "STORM_contiguous_t.n_bitmaps_vector": 72, "STORM_contiguous_t.intsec_func": 80,
    "STORM_contiguous_t.alignment": 88, "STORM_contiguous_t.scalar_cutoff": 92,
    "STORM_contiguous_t.b200": 96,                      # appended after the reference's 96 bytes
    "sizeof.STORM_contiguous_bitmap_t": 24, "STORM_contiguous_bitmap_t.data": 0,
    "STORM_contiguous_bitmap_t.scalar": 8, "STORM_contiguous_bitmap_t.n_scalar": 16,
    "STORM_t.conts": 0, "STORM_t.n_conts": 8, "STORM_t.m_conts": 12, "STORM_t.b200": 16,
    "sizeof.STORM_bitmap_cont_t": 32, "STORM_bitmap_cont_t.bitmaps": 0, "STORM_bitmap_cont_t.block_ids": 8,
    "STORM_bitmap_cont_t.n_bitmaps": 16, "STORM_bitmap_cont_t.m_bitmaps": 20,
    "STORM_bitmap_cont_t.prev_inserted_value": 24,
    "sizeof.STORM_bitmap_t": 128, "alignof.STORM_bitmap_t": 64, "STORM_bitmap_t.data": 0,
    "STORM_bitmap_t.scalar": 64, "STORM_bitmap_t.n_bits_set": 76, "STORM_bitmap_t.n_missing": 84,
    "STORM_bitmap_t.m_scalar": 88, "STORM_bitmap_t.id": 92,
}


@pytest.mark.parametrize("compiler,std", [("gcc", "-std=c11"), ("g++", "-std=c++17")])
def test_public_struct_layout_matches_reference(compiler, std):
    with tempfile.TemporaryDirectory() as d:
        ext = ".c" if compiler == "gcc" else ".cpp"
        src = os.path.join(d, "layout" + ext)
        code = LAYOUT_C if compiler == "gcc" else LAYOUT_C.replace("_Alignof", "alignof")
        open(src, "w").write(code)
        exe = os.path.join(d, "layout")
        subprocess.check_call([compiler, std, "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        got = dict((k, int(v)) for k, v in (l.split() for l in subprocess.check_output([exe], text=True).splitlines()))
    for k, v in REFERENCE_LAYOUT.items():
        assert got[k] == v, (k, got[k], v)


def test_host_side_conventions_without_gpu(lib):
    """storm.c:1001-1018,1031-1137,1139-1147,844-875: construction is pure host work."""
    import stormbitmaps_b200 as sb
    u32p = C.POINTER(C.c_uint32)
    assert lib.STORM_contig_add(None, None, 0) == -1
    assert lib.STORM_contig_clear(None) == -1
    assert lib.STORM_contig_pairw_intersect_cardinality(None) == 2**64 - 1
    assert lib.STORM_pairw_intersect_cardinality(None) == 2**64 - 1
    assert lib.STORM_add(None, None, 0) == -1
    c = sb.StormContiguous(65536)
    assert lib.STORM_contig_add(c._h, None, 3) == -2
    assert c.clear() == 0                                  # nothing allocated yet
    assert c.add([]) == 0                                  # empty list: no row (D7)
    assert c.add([1, 5, 9]) == 3
    assert c.add([5, 5, 9, 10]) == 4                       # returns n_values, duplicates collapse
    with pytest.raises(sb.StormError):
        c.add([65536])                                     # out of range is rejected, not written
    with pytest.raises(sb.StormError, match="vector_length"):
        c.add([7, 300, 64000, 70000, 12])                  # ... also after in-range values: the half-written row is undone
    # public fields are populated like the reference's
    class Contig(C.Structure):
        _fields_ = [("data", C.POINTER(C.c_uint64)), ("scalar", u32p), ("n_scalar", u32p), ("bitmaps", C.c_void_p),
                    ("n_data", C.c_uint64), ("m_data", C.c_uint64), ("tot_scalar", C.c_uint64), ("m_scalar", C.c_uint64),
                    ("vector_length", C.c_uint64), ("n_bitmaps_vector", C.c_uint32), ("intsec_func", C.c_void_p),
                    ("alignment", C.c_uint32), ("scalar_cutoff", C.c_uint32)]
    s = Contig.from_address(c._h)
    assert (s.n_data, s.vector_length, s.n_bitmaps_vector, s.scalar_cutoff) == (2, 65536, 1024, 200)
    assert s.data[0] == (1 << 1) | (1 << 5) | (1 << 9) and s.data[1024] == (1 << 5) | (1 << 9) | (1 << 10)
    assert [s.n_scalar[0], s.n_scalar[1]] == [3, 3] and s.tot_scalar == 6
    assert [s.scalar[i] for i in range(6)] == [1, 5, 9, 5, 9, 10]
    assert all(s.data[2 * 1024 + w] == 0 for w in range(1024))               # nothing left behind by the rejected rows
    assert c.add([64000, 3, 3]) == 3                       # unsorted input is accepted like the reference's (storm.c:1103-1115)
    assert s.n_data == 3 and s.data[2 * 1024] == 1 << 3 and s.data[2 * 1024 + 1000] == 1 << (64000 - 64000 // 64 * 64)
    assert s.n_scalar[2] == 2
    fn = C.CFUNCTYPE(C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_size_t)(s.intsec_func)
    a = (C.c_uint64 * 2)(0b1011, 1 << 63)
    b = (C.c_uint64 * 2)(0b0110, 1 << 63)
    assert fn(a, b, 2) == 2                                # the kept host function pointer is callable
    assert c.clear() == 1 and s.n_data == 0
    c.free()
    # cutoff rule min(200, M/200) (storm.c:1016)
    for M, want in [(65536, 200), (4096, 20), (100, 0), (1 << 20, 200)]:
        k = sb.StormContiguous(M)
        assert Contig.from_address(k._h).scalar_cutoff == want
        k.free()


def test_storm_t_host_builder_matches_oracle_sizes(lib, orc, golden):
    """Block splitting / list-vs-bitmap threshold / serialized size (storm.c:692-758,372-394,963-973)."""
    import stormbitmaps_b200 as sb
    from conftest import case_rows
    for name in ("mix_65536x200_1000_30000", "edge_block_boundaries", "edge_empty_row_middle", "c2s_524288x120_d20971"):
        case = next(c for c in golden["cases"] if c["name"] == name)
        with sb.Storm() as s:
            for p in case_rows(orc, case):
                assert s.add(p) == 1
            assert s.serialized_size() == case["ref"]["storm_serialized_size"], name


def test_queries_fail_loudly_without_a_device(lib):
    import stormbitmaps_b200 as sb
    if lib.STORM_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    c = sb.StormContiguous(256)
    c.add([1, 2]); c.add([2, 3])
    with pytest.raises(sb.StormError, match="no CUDA device"):
        c.pairw_intersect_cardinality()
    with pytest.raises(sb.StormError):
        sb.wrapper_diag(np.ones((4, 4), dtype=np.uint64))
    s = sb.Storm()
    s.add([1, 2]); s.add([2, 3])
    with pytest.raises(sb.StormError):
        s.pairw_intersect_cardinality()


# ---- a C program written against storm.h, linked with libstorm_b200.so -----------------------------------
HOST_FIELDS = ("rows", "words", "cutoff", "naive", "fptr01", "serialized", "round2_rows", "null_query", "null_add")
QUERY_FIELDS = ("contig", "contig_blocked", "contig_list", "contig_blocked_list", "storm", "storm_blocked", "wrapper", "wrapper_blocked",
                "round2_contig", "round2_storm")


def build_dropin_driver(workdir):
    """tests/drivers/dropin_driver.c (the calls benchmark.cpp makes, nothing but <storm.h>) against this repo's
    header and shared library -- the link-time drop-in of INTEGRATION.md section 2."""
    from stormbitmaps_b200 import build
    build.build()
    exe = os.path.join(workdir, "dropin_driver")
    pkg = os.path.join(ROOT, "stormbitmaps_b200")
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "drivers", "dropin_driver.c"), "-L", pkg, "-lstorm_b200",
                           "-Wl,-rpath," + pkg, "-o", exe])
    return exe


def run_dropin_driver(exe, args):
    line = subprocess.check_output([exe] + args.split(), text=True, timeout=300).strip()
    return dict(kv.split("=") for kv in line.split())


def test_c_caller_compiles_links_and_keeps_host_semantics(lib):
    """The same C source gives the same host-visible state (public fields, return codes, kept function pointer)
    as when it is linked with the reference's storm.c (tests/golden/dropin_driver_v1.json, minted by
    tools/make_golden_driver.py); without a device every query is the sentinel, never a CPU result."""
    import json
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "dropin_driver_v1.json")))["cases"]
    have_gpu = lib.STORM_b200_device_count() > 0
    with tempfile.TemporaryDirectory() as d:
        exe = build_dropin_driver(d)
        for args, want in golden.items():
            got = run_dropin_driver(exe, args)
            for k in HOST_FIELDS:
                assert got[k] == want[k], (args, k, got[k], want[k])
            for k in QUERY_FIELDS:
                assert got[k] == (want[k] if have_gpu else str(2**64 - 1)), (args, k, got[k])


def test_route_model_decisions_on_the_measured_cases(lib):
    """The STORM_t route cost model is pure arithmetic: its decisions on the cases it was fitted to
    (profiles/r02_sparse_routes_ranges_v2.jsonl: 10 000 x 524 288 sparse route up to 3 000 values per row, dense from
    4 000; C4 dense) and its estimates within a factor 1.5 of the measured milliseconds."""
    import stormbitmaps_b200 as sb
    measured = {1: 0.027, 5: 0.032, 104: 0.216, 300: 0.619, 524: 1.20, 1000: 2.27, 1500: 3.18, 2097: 4.56}   # stream kernel, ms
    for nnz, ms in measured.items():
        m = sb.storm_route_model(10000, 8192, nnz, min(8, nnz), nnz + 40)
        assert m["route"] == "sparse", (nnz, m)
        assert ms / 1.7 < m["sparse_s"] * 1e3 < ms * 1.5, (nnz, m)
        assert 6.36 / 1.2 < m["dense_s"] * 1e3 < 6.36 * 1.2, m                      # densified route: 6.36 ms measured
    for nnz in (4000, 5242, 20971):
        assert sb.storm_route_model(10000, 8192, nnz, 8, nnz + 200)["route"] == "dense", nnz
    assert 8.83 / 1.2 < sb.storm_route_model(10000, 8192, 4000, 8, 4300)["sparse_s"] * 1e3 < 8.83 * 1.2   # 8.83 ms on the stream kernel
    assert sb.storm_route_model(20000, 2048, 16, 2, 40)["route"] == "sparse"        # 0.12 ms against 6.15 ms measured
    assert sb.storm_route_model(100000, 16384, 10486, 16, 11000)["route"] == "dense"   # C4: rows too long for the stream kernel
    assert sb.storm_route_model(10000, 8192, 104, 8, 150, n_bitmap_blocks=3)["route"] == "dense"   # bitmap blocks: block kernel only
    # without the FP4 form the tile kernel is half as fast and the crossover moves up
    assert sb.storm_route_model(10000, 8192, 4000, 8, 4300, fp4=False)["route"] == "sparse"


def test_split_route_model(lib):
    """The split route's estimate (pure arithmetic): a few heavy rows among light ones cost little more than the light
    rows alone through the stream kernel and far less than densifying everything; with a quarter of the rows heavy
    (C2's log-uniform mixed level) the densified route stays cheaper."""
    import stormbitmaps_b200 as sb
    light = sb.storm_route_model(10000, 8192, 300, 8, 340)                          # no heavy rows: 1.19 ms measured
    few = sb.storm_split_model(10000, 10, 9990 * 300.0, 9990 * 300.0 + 10 * 100000.0, 8, 80)
    assert light["sparse_s"] < few < light["sparse_s"] + 0.5e-3, (light, few)
    assert few < light["dense_s"] / 3
    # 16 Mbit rows: dense costs N^2 W / 2 / rate = 0.2 s, the split route still milliseconds
    wide = sb.storm_route_model(10000, 262144, 310, 8, 200000, n_bitmap_blocks=80)
    assert sb.storm_split_model(10000, 10, 9990 * 300.0, 9990 * 300.0 + 10 * 100000.0, 256, 80) < wide["dense_s"] / 20
    mixed = sb.storm_split_model(10000, 2800, 7200 * 900.0, 7200 * 900.0 + 2800 * 75000.0, 8, 2800 * 6.0)
    assert mixed > 5 * sb.storm_route_model(10000, 8192, 21000, 8, 262144, n_bitmap_blocks=16800)["dense_s"]
    assert sb.storm_split_model(1, 0, 0, 0, 1, 0) < 0


def test_per_device_thread_pool_without_a_device(lib):
    """The pool that issues the per-device calls of a multi-device query (devices.cu: DevicePool), driven through its
    test hook on the CPU: every job runs exactly once for 1 .. 9 "devices", thousands of dispatches back to back
    (workers polling) and after pauses (workers asleep), from two caller threads at once; a failing job's code and
    message reach the caller; with the threads off the same jobs run on the caller."""
    import threading, time
    import stormbitmaps_b200 as sb
    for n in range(0, 10):
        assert lib.STORM_b200_selftest_device_threads(n, -1) == 0, n
    for k in range(3000):
        assert lib.STORM_b200_selftest_device_threads(2 + k % 7, -1) == 0
        if k % 1000 == 999:
            time.sleep(0.02)
    assert lib.STORM_b200_selftest_device_threads(8, 5) == -1
    assert "job 5 failed on purpose" in sb.last_error()
    assert lib.STORM_b200_selftest_device_threads(8, 0) == -1 and "job 0 failed" in sb.last_error()
    bad = []

    def hammer():
        for k in range(2000):
            if lib.STORM_b200_selftest_device_threads(3 + k % 5, -1) != 0:
                bad.append(k)
    ts = [threading.Thread(target=hammer) for _ in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not bad
    assert sb.set_device_threads(False) is True
    try:
        assert lib.STORM_b200_selftest_device_threads(8, -1) == 0
        assert lib.STORM_b200_selftest_device_threads(8, 2) == -1 and "job 2 failed" in sb.last_error()
    finally:
        sb.set_device_threads(True)


def test_header_declares_every_reference_prototype():
    """Drop-in means every function the reference's storm.h declares is declared (and exported) here.  The name list
    is a committed fixture (minted with: grep -oE '\\bSTORM_[a-zA-Z0-9_]+ *\\(' /root/reference/storm.h | tr -d ' (' | sort -u, minus the STORM_ALIGN macro);
    where the reference tree is present the fixture itself is checked against it."""
    want = set(open(os.path.join(ROOT, "tests", "golden", "reference_storm_h_prototypes.txt")).read().split())
    assert len(want) == 42
    ours = _declared_functions("storm.h")
    assert not (want - ours), f"reference prototypes missing from include/storm.h: {sorted(want - ours)}"
    ref_header = "/root/reference/storm.h"
    if os.path.isfile(ref_header):
        text = re.sub(r"/\*.*?\*/", "", open(ref_header).read(), flags=re.S)
        live = set(re.findall(r"\b(STORM_[A-Za-z0-9_]+)\s*\(", text)) - {"STORM_ALIGN"}
        assert live == want, (sorted(live - want), sorted(want - live))


def test_per_pair_host_helpers(lib, orc):
    """storm.h:56-61, 207-210, 220-221: the reference's CPU helpers for ONE pair, restated on the host
    (host_pairs.cu).  Exact set semantics, checked against numpy and the oracle's u16 restatement."""
    u16p, u32p, u64p = C.POINTER(C.c_uint16), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    rng = np.random.default_rng(7)
    for na, nb in [(0, 5), (1, 1), (7, 9), (8, 8), (64, 4000), (4095, 4095), (3, 4095), (500, 17)]:
        a = np.unique(rng.integers(0, 65536, na)).astype(np.uint16)
        b = np.unique(rng.integers(0, 65536, nb)).astype(np.uint16)
        if len(a) and len(b):
            b[0] = a[0] = min(a[0], b[0]); a.sort(); b.sort(); a = np.unique(a); b = np.unique(b)
        want = len(np.intersect1d(a, b))
        got = lib.STORM_intersect_vector16_cardinality(a.ctypes.data_as(u16p), b.ctypes.data_as(u16p), len(a), len(b))
        assert got == want == (orc.intersect_u16(a, b) if len(a) and len(b) else want), (na, nb)
    a = np.array([1, 4, 9, 16, 25, 36], dtype=np.uint32); b = np.array([0, 4, 5, 16, 36, 99], dtype=np.uint32)
    out = np.zeros(12, dtype=np.uint32)
    n = lib.STORM_intersect_vector32_unsafe(a.ctypes.data_as(u32p), b.ctypes.data_as(u32p), 6, 6, out.ctypes.data_as(u32p))
    assert n == 6 and out[:6].tolist() == [1, 1, 3, 3, 5, 4]              # (index in a, index in b) pairs
    assert lib.STORM_intersect_vector32_unsafe(a.ctypes.data_as(u32p), b.ctypes.data_as(u32p), 6, 6, None) == 0
    # list probe: the shorter list goes into the other row's bitmap (storm.c:108-129)
    M = 4096
    l1 = np.array([3, 64, 700, 4095], dtype=np.uint32); l2 = np.array([3, 5, 64, 65, 700, 701], dtype=np.uint32)
    b1 = np.zeros(M // 64, dtype=np.uint64); b2 = np.zeros(M // 64, dtype=np.uint64)
    for v in l1: b1[v >> 6] |= np.uint64(1) << np.uint64(v & 63)
    for v in l2: b2[v >> 6] |= np.uint64(1) << np.uint64(v & 63)
    for (x1, x2, y1, y2) in [(b1, b2, l1, l2), (b2, b1, l2, l1)]:
        assert lib.STORM_intersect_bitmaps_scalar_list(x1.ctypes.data_as(u64p), x2.ctypes.data_as(u64p), y1.ctypes.data_as(u32p),
                                                       y2.ctypes.data_as(u32p), len(y1), len(y2)) == 3
    # blocks and rows: every (list | bitmap) x (list | bitmap) combination, exact (D1 not reproduced)
    def row(values):
        h = lib.STORM_bitmap_cont_new()
        v = np.asarray(values, dtype=np.uint32)
        assert lib.STORM_bitmap_cont_add(h, v.ctypes.data_as(u32p), len(v)) == 1
        return h
    dense = np.sort(rng.choice(3 * 65536, 40000, replace=False)).astype(np.uint32)       # bitmap blocks
    sparse = np.sort(rng.choice(3 * 65536, 900, replace=False)).astype(np.uint32)        # list blocks
    sparse2 = np.unique(np.concatenate([sparse[::3], dense[::50]])).astype(np.uint32)
    rows = {"dense": dense, "sparse": sparse, "sparse2": sparse2, "dense2": dense[5000:35000]}
    handles = {k: row(v) for k, v in rows.items()}
    scratch = np.zeros(64, dtype=np.uint32)
    f = C.cast(lib.STORM_get_intersect_count_func(1024), C.c_void_p)
    for ka in rows:
        for kb in rows:
            want = len(np.intersect1d(rows[ka], rows[kb]))
            assert lib.STORM_bitmap_cont_intersect_cardinality(handles[ka], handles[kb]) == want, (ka, kb)
            assert lib.STORM_bitmap_cont_intersect_cardinality_premade(handles[ka], handles[kb], f, scratch.ctypes.data_as(u32p)) == want
    for h in handles.values():
        lib.STORM_bitmap_cont_free(h)
    # one block built with bits + list (storm.c:467-519): duplicates enter the list once
    blk = lib.STORM_bitmap_new()
    v = np.array([5, 5, 9, 70, 9, 65535], dtype=np.uint32)
    assert lib.STORM_bitmap_add_with_scalar(blk, v.ctypes.data_as(u32p), len(v)) == 6
    blk2 = lib.STORM_bitmap_new()
    w = np.array([9, 65535, 100], dtype=np.uint32)
    assert lib.STORM_bitmap_add_scalar_only(blk2, w.ctypes.data_as(u32p), 3) == 3
    assert lib.STORM_bitmap_intersect_cardinality(blk, blk2) == 2 == lib.STORM_bitmap_intersect_cardinality_func(blk, blk2, f)
    assert lib.STORM_bitmap_intersect_cardinality(blk, None) == 0
    lib.STORM_bitmap_free(blk); lib.STORM_bitmap_free(blk2)


def test_foreign_compute_func_is_classified_not_assumed(lib):
    """storm.c:132-150 applies whatever per-pair function it is handed.  A foreign function is identified by its
    values on probe vectors: a union / diff / intersect count selects that operation, anything else is an error --
    never a silent intersect total."""
    import stormbitmaps_b200 as sb
    proto = C.CFUNCTYPE(C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_size_t)
    bogus = proto(lambda a, b, n: 42)
    vals = np.ones((4, 4), dtype=np.uint64)
    rc = lib.STORM_wrapper_diag(4, vals.ctypes.data_as(C.POINTER(C.c_uint64)), 4, C.cast(bogus, C.c_void_p))
    assert rc == 2**64 - 1 and "neither" in sb.last_error()
    xor_count = proto(lambda a, b, n: sum(bin(a[k] ^ b[k]).count("1") for k in range(n)))
    rc = lib.STORM_wrapper_diag(4, vals.ctypes.data_as(C.POINTER(C.c_uint64)), 4, C.cast(xor_count, C.c_void_p))
    if lib.STORM_b200_device_count() == 0:
        assert rc == 2**64 - 1 and "neither" not in sb.last_error()      # classified; then fails for want of a device
    else:
        assert rc == 0                                                    # four identical rows: every xor count is 0


def test_device_set_api_without_gpu(lib):
    import stormbitmaps_b200 as sb
    prev = lib.STORM_b200_set_devices(2)
    assert prev >= 0
    assert lib.STORM_b200_set_devices(prev if prev > 1 else 1) == 2
    ids = (C.c_int * 2)(0, 0)
    assert lib.STORM_b200_set_device_list(ids, 2) == 0
    assert lib.STORM_b200_set_device_list(None, 0) == 0                   # back to the default: the current device
    if lib.STORM_b200_device_count() == 0:
        got = (C.c_int * 4)()
        assert lib.STORM_b200_get_devices(got, 4) < 0 and "no CUDA device" in sb.last_error()


def test_add_dense_builds_the_same_container_as_add(lib, orc):
    """STORM_b200_contig_add_dense (rows as bitmaps) leaves the public fields exactly as STORM_contig_add on the rows'
    sorted positions does: mirror words, per-row counts, position lists of the rows below the cutoff, skipped empty rows."""
    import stormbitmaps_b200 as sb
    from oracle import oracle as O
    M = 5000                                              # not a multiple of 64: the tail bits are checked
    draws = [0, 1, 3, 24, 25, 26, 400, 3000]
    rows = [orc.gen_row_positions(9, i, draws[i % len(draws)], M) for i in range(64)]
    vals = O.positions_to_dense(rows, M)

    class Contig(C.Structure):
        _fields_ = [("data", C.POINTER(C.c_uint64)), ("scalar", C.POINTER(C.c_uint32)), ("n_scalar", C.POINTER(C.c_uint32)), ("bitmaps", C.c_void_p),
                    ("n_data", C.c_uint64), ("m_data", C.c_uint64), ("tot_scalar", C.c_uint64), ("m_scalar", C.c_uint64),
                    ("vector_length", C.c_uint64), ("n_bitmaps_vector", C.c_uint32), ("intsec_func", C.c_void_p),
                    ("alignment", C.c_uint32), ("scalar_cutoff", C.c_uint32)]
    a, b = sb.StormContiguous(M), sb.StormContiguous(M)
    for r in rows:
        a.add(r)
    b.add_dense(vals[:40])
    b.add_dense(vals[40:])
    sa, sb_ = Contig.from_address(a._h), Contig.from_address(b._h)
    assert sa.n_data == sb_.n_data == sum(1 for r in rows if len(r)) and sa.tot_scalar == sb_.tot_scalar
    W = sa.n_bitmaps_vector
    assert [sa.data[i] for i in range(sa.n_data * W)] == [sb_.data[i] for i in range(sa.n_data * W)]
    assert [sa.n_scalar[i] for i in range(sa.n_data)] == [sb_.n_scalar[i] for i in range(sa.n_data)]
    assert [sa.scalar[i] for i in range(sa.tot_scalar)] == [sb_.scalar[i] for i in range(sa.tot_scalar)]
    bad = vals[:1].copy()
    bad[0, -1] |= np.uint64(1) << np.uint64(63)           # bit 5055 >= vector_length
    with pytest.raises(sb.StormError, match="vector_length"):
        b.add_dense(bad)
    a.free(); b.free()
