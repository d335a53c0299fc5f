#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): python tools/ncu_summary.py rep [out.md]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = [
 "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__cycles_active.avg",
 "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
 "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
 "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
 "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_op_imma_cycles_active.avg.pct_of_peak_sustained_active",
 "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
 "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
 "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
 "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct",
 "lts__t_sector_hit_rate.pct", "smsp__average_warp_latency_issue_stalled_long_scoreboard", "sm__sass_inst_executed_op_shared_st.sum",
]
out = []
for k, row in enumerate(rows[2:]):
    d = dict(zip(hdr, row))
    out.append(f"### launch {k}: {d.get('Kernel Name','?')[:90]}")
    for key in KEYS:
        if key in d:
            out.append(f"- `{key}` = {d[key]} {units[hdr.index(key)]}")
    # every tensor / stall metric that exists
    for h in hdr:
        if ("tensor" in h and "pct_of_peak_sustained_active" in h) or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
            if h not in KEYS:
                out.append(f"- `{h}` = {d[h]} {units[hdr.index(h)]}")
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
