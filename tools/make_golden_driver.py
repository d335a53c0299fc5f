#!/usr/bin/env python
"""Mint tests/golden/dropin_driver_v1.json: the output of tests/drivers/dropin_driver.c compiled against the
UNMODIFIED reference (its storm.h and storm.c where they lie under /root/reference) for a few argument sets.
The same source compiled against this repo's include/storm.h + libstorm_b200.so must print the same line
(tests/test_abi.py on the CPU for the host-side fields, tests/test_parity_gpu.py on the GPU for every total).

    python tools/make_golden_driver.py [/root/reference]
"""
import json, os, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
CASES = ["4096 40 300 7", "65536 300 6553 42", "65536 60 150 3", "524288 50 30000 5", "1000 130 128 9", "131072 64 1 11"]

with tempfile.TemporaryDirectory() as d:
    exe = os.path.join(d, "driver_ref")
    subprocess.check_call(["gcc", "-std=gnu99", "-O2", "-march=native", "-DNDEBUG", "-w", "-I", REF,
                           os.path.join(ROOT, "tests", "drivers", "dropin_driver.c"), os.path.join(REF, "storm.c"), "-o", exe])
    out = {"_comment": "dropin_driver.c linked with the reference's storm.c (gcc -std=gnu99 -O2 -march=native); "
                       "key=value fields of its one output line per argument set 'M N draws seed'.  Under "
                       "-fsanitize=address,undefined -UNDEBUG the reference is clean on these inputs except for its "
                       "shift by more than 63 in the list probe (storm.c:123, SURVEY.md D10: x86 masks the count, and the "
                       "values equal the naive count) on the two all-sparse sets; every field is asserted equal to the "
                       "driver's own naive popcount before it is written",
           "reference": "StormBitmaps @ 2eae567", "cases": {}}
    for args in CASES:
        line = subprocess.check_output([exe] + args.split(), text=True).strip()
        fields = dict(kv.split("=") for kv in line.split())
        assert (fields["naive"] == fields["contig"] == fields["contig_blocked"] == fields["contig_list"] == fields["contig_blocked_list"]
                == fields["storm"] == fields["storm_blocked"] == fields["wrapper"] == fields["wrapper_blocked"]), line
        out["cases"][args] = fields
path = os.path.join(ROOT, "tests", "golden", "dropin_driver_v1.json")
json.dump(out, open(path, "w"), indent=1)
print(path, len(out["cases"]), "cases")
