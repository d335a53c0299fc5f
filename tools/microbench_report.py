#!/usr/bin/env python
"""Issue-rate probes and the clock64 calibration on the GPU box -> one JSON object (profiles/rNN_microbench.json).

Per-clock figures divide by the probe's OWN clock64 delta times the calibrated SM cycles per clock64 tick (a 250 ms
spin on the idle device, nvidia-smi's clocks.sm sampled beside it), not by an assumed clock."""
import json
import os
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stormbitmaps_b200 as sb  # noqa: E402

sb.load()
info = sb.device_info(0)
out = {"device": info}

# calibration: nvidia-smi beside the spin
path = tempfile.mktemp(suffix=".csv")
smi = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits", "-lms", "20"],
                       stdout=open(path, "w"), stderr=subprocess.DEVNULL)
time.sleep(0.1)
ticks = [sb.microbench(8)[1] for _ in range(3)]
smi.terminate(); smi.wait()
rows = [l.split(",") for l in open(path) if "," in l]
smi_mhz = sorted(float(r[0]) for r in rows)[len(rows) // 2] if rows else None
smi_max = max(float(r[1]) for r in rows) if rows else None
ratio = smi_mhz / ticks[-1] if smi_mhz else None
out["clock64_calibration"] = {"ticks_per_us": ticks, "nvidia_smi_sm_mhz_median": smi_mhz, "nvidia_smi_sm_max_mhz": smi_max,
                              "sm_cycles_per_tick": ratio, "rounded": round(ratio) if ratio else None}
k = float(round(ratio)) if ratio and abs(ratio - round(ratio)) < 0.04 else (ratio or 1.0)

for kind, name in [(0, "popc32"), (1, "lop3_32"), (2, "iadd32"), (3, "mix_1popc_2lop3")]:
    rate, mhz = sb.microbench(kind)
    out[name] = {"thread_instr_per_s": rate, "clock64_ticks_per_us": mhz, "sm_mhz": mhz * k,
                 "per_clk_per_sm": rate / (mhz * k * 1e6) / info["sm_count"]}
for kind, name, pipe in [(4, "umma_i8_cta_group1", 8192), (5, "umma_i8_cta_group2", 8192),
                         (6, "umma_mxf4_cta_group1", 16384), (7, "umma_mxf4_cta_group2", 16384)]:
    rate, mhz = sb.microbench(kind)
    mac = rate / 2 / (mhz * k * 1e6) / info["sm_count"] if mhz else None
    out[name] = {"ops_per_s": rate, "tops": rate / 1e12, "issue_loop_clock64_ticks_per_us": mhz, "issue_loop_sm_mhz": mhz * k,
                 "mac_per_clk_per_sm": mac, "pipe_mac_per_clk_per_sm": pipe, "frac_of_pipe": mac / pipe if mac else None}
print(json.dumps(out, indent=1))
