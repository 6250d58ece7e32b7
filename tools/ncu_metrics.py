#!/usr/bin/env python
"""Dump the metrics DESIGN.md / profiles/r01_summary.md quote from an .ncu-rep (ncu --set full) into
JSON:  python tools/ncu_metrics.py report.ncu-rep out_metrics.json [out_traffic.json workload kernel-label]"""
import csv, io, json, subprocess, sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, vals = rows[0], rows[1], rows[-1]
out = {}
for k in KEEP:
    if k in hdr:
        i = hdr.index(k)
        out[k] = {"value": vals[i], "unit": units[i]}
out["kernel"] = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
json.dump(out, open(sys.argv[2], "w"), indent=1)
if len(sys.argv) > 5:
    def to_bytes(m):
        v, u = float(out[m]["value"]), out[m]["unit"].lower()
        return int(round(v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]))
    json.dump({"workload": sys.argv[4], "kernel": sys.argv[5], "dram_bytes_read": to_bytes("dram__bytes_read.sum"),
               "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
               "source": "ncu --set full --clock-control none, " + sys.argv[2]}, open(sys.argv[3], "w"), indent=1)
print(json.dumps({k: v["value"] for k, v in out.items() if isinstance(v, dict)}, indent=1)[:1500])
