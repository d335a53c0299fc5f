"""Union / symmetric-difference cardinalities: the oracle against the
values the unmodified reference produced (tests/golden/golden_ops_v1.json, minted by
tools/make_golden_ops.py), and -- on a GPU -- the CUDA path through the C-ABI against both."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import case_rows
from oracle import oracle as O

OPS = {"union": 1, "diff": 2}
GOLDEN_OPS = os.path.join(os.path.dirname(__file__), "golden", "golden_ops_v1.json")


def _ops_cases():
    with open(GOLDEN_OPS) as f:
        return json.load(f)["cases"]


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", [c["name"] for c in _ops_cases()])
def test_oracle_setops_match_reference_golden(orc, golden, name):
    rec = next(c for c in _ops_cases() if c["name"] == name)
    case = next(c for c in golden["cases"] if c["name"] == name)
    vals = O.positions_to_dense(case_rows(orc, case), case["M"])
    n = vals.shape[0]
    pops = np.array([int(np.unpackbits(r.view(np.uint8)).sum()) for r in vals], dtype=np.int64)
    inter = orc.wrapper_diag(vals)
    for op, opid in OPS.items():
        assert orc.wrapper_diag_op(vals, opid) == rec[op]
        # the identity the GPU path uses: |a|b| = |a|+|b|-|a&b|, |a^b| = |a|+|b|-2|a&b|
        assert rec[op] == (n - 1) * int(pops.sum()) - opid * inter
        if op + "_pairs_sha256" in rec:
            assert _sha(orc.rect_counts_op(vals, 0, n, 0, n, opid)) == rec[op + "_pairs_sha256"]
    assert orc.wrapper_diag_op(vals, 0) == inter == case["exact"]


@pytest.fixture(scope="module")
def sb():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    import stormbitmaps_b200 as sb
    sb.load()
    return sb


def _device_rows(sb, vals):
    import torch
    n, w = vals.shape
    rows, _ = sb.alloc_rows(n, w * 64)
    rows[:, :w] = torch.from_numpy(vals.view(np.int64)).cuda()
    return rows


@pytest.mark.gpu
@pytest.mark.parametrize("name", [c["name"] for c in _ops_cases()])
def test_gpu_setops_match_reference_golden(sb, orc, golden, name):
    """STORM_wrapper_diag handed the union / diff kernel pointer, exactly as a C caller of the reference
    would (storm.c:132-150 + libalgebra.h:3142-3236), and the device-buffer entry points."""
    rec = next(c for c in _ops_cases() if c["name"] == name)
    case = next(c for c in golden["cases"] if c["name"] == name)
    vals = O.positions_to_dense(case_rows(orc, case), case["M"])
    n, W = vals.shape
    rows = _device_rows(sb, vals)
    for op in OPS:
        assert sb.wrapper_diag(vals, op=op) == rec[op]
        assert int(sb.pairw_op_device(rows, op, n_words=W).item()) == rec[op]
        if op + "_pairs_sha256" in rec:
            counts, total = sb.pairw_rect_op_device(rows, op, 0, n, 0, n, n_words=W)
            assert _sha(counts.cpu().numpy().view(np.uint32)) == rec[op + "_pairs_sha256"]
            assert int(total.item()) == rec[op]
    assert sb.wrapper_diag(vals, op="intersect") == case["exact"]


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["popc", "umma", "fp4"])
def test_gpu_setops_rectangles_and_square(sb, orc, kernel):
    import torch
    M, N = 8192, 700
    vals = orc.gen_dense_uniform(31, N, 2500, M)
    W = vals.shape[1]
    rows = _device_rows(sb, vals)
    pops = sb.row_popcounts_device(rows, n_words=W).cpu().numpy()
    assert (pops == np.unpackbits(vals.view(np.uint8), axis=1).sum(axis=1)).all()
    for op, opid in OPS.items():
        assert int(sb.pairw_op_device(rows, op, n_words=W, kernel=kernel).item()) == orc.wrapper_diag_op(vals, opid)
        for (i0, i1, j0, j1) in [(0, 64, 0, 64), (100, 333, 50, 699), (500, 700, 0, 300), (10, 11, 0, 700)]:
            counts, total = sb.pairw_rect_op_device(rows, op, i0, i1, j0, j1, n_words=W, kernel=kernel)
            want = orc.rect_counts_op(vals, i0, i1, j0, j1, opid)
            assert (counts.cpu().numpy().view(np.uint32) == want).all(), (op, i0, j0)
            assert int(total.item()) == int(want.sum(dtype=np.uint64))
        a, b = vals[:300], vals[300:]
        assert sb.wrapper_square(a, b, op=op) == orc.wrapper_square_op(a, b, opid)
    # a foreign function pointer (or NULL) means intersect, as documented in storm.h
    assert sb.wrapper_diag(vals) == orc.wrapper_diag(vals)
