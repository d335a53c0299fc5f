"""CPU tests of bench.py's contract: the reference arm (the compiled reference on the host cores) prints one
JSON line with the keys the driver reads, and the product arm has no CPU path -- without a device it exits
non-zero instead of printing a number."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    from oracle import oracle as O
    O.build()
    if not O.have_reference():
        pytest.skip("oracle/_ref was not built here (no /root/reference)")
    p = _run("--impl", "reference", "--workload", "custom", "--rows", "400", "--bits", "4096", "--steps", "2", "--warmup", "1")
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "xxt_wordpair_and_popcnt_per_s" and line["unit"] == "wp/s"
    assert line["higher_is_better"] is True and line["steps"] == 2 and line["warmup"] == 1 and line["n_gpus"] == 1
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["rows"] == 400 and line["config"]["bits"] == 4096


def test_product_arm_has_no_cpu_path():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    p = _run("--workload", "custom", "--rows", "300", "--bits", "4096", "--steps", "1", "--warmup", "1", timeout=200)
    assert p.returncode != 0
    assert not any(l.startswith("{") for l in p.stdout.splitlines())        # no bench line without a GPU
