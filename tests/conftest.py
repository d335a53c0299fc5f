"""pytest configuration: registers the `gpu` marker and shared fixtures.

`-m "not gpu"` runs on a CPU-only box (oracle vs golden vectors, host logic,
C-ABI export checks, gloo multi-process sharding); `-m gpu` are the parity
tests proper and call the CUDA path through the C-ABI.
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "golden_v1.json")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden():
    with open(GOLDEN) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle as O
    O.build()
    return O.Oracle()


def case_rows(orc, case):
    """Regenerate a golden case's position lists from its recorded recipe."""
    g, M, N = case["gen"], case["M"], case["N"]
    if g["kind"] == "uniform":
        return [orc.gen_row_positions(g["seed"], i, g["n_draws"], M) for i in range(N)]
    if g["kind"] == "per_row_draws":
        return [orc.gen_row_positions(g["seed"], i, d, M) for i, d in enumerate(g["draws"])]
    if g["kind"] == "explicit":
        return [np.asarray(p, dtype=np.uint32) for p in g["rows"]]
    raise ValueError(g["kind"])
