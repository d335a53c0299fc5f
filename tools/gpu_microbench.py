#!/usr/bin/env python
"""Print the CUDA-core issue rates measured by STORM_b200_microbench (run on the GPU box)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stormbitmaps_b200 as sb
info = sb.device_info(0)
out = {"device": info}
for kind, name in [(0, "popc32"), (1, "lop3_32"), (2, "iadd32"), (3, "mix_1popc_2lop3")]:
    rate, mhz = sb.microbench(kind)
    out[name] = {"thread_instr_per_s": rate, "sm_mhz": mhz,
                 "per_clk_per_sm": rate / (mhz * 1e6) / info["sm_count"]}
for kind, name in [(4, "umma_i8_cta_group1"), (5, "umma_i8_cta_group2")]:
    rate, _ = sb.microbench(kind)
    out[name] = {"int8_ops_per_s": rate, "tops": rate / 1e12,
                 "mac_per_clk_per_sm_at_1965mhz": rate / 2 / 1.965e9 / info["sm_count"]}
print(json.dumps(out, indent=1))
