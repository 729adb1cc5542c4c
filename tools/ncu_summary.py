#!/usr/bin/env python
"""Condenses an .ncu-rep (read here with `ncu -i ... --page raw --csv`) into the metrics profiles/ keeps."""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
args = [a for a in sys.argv[1:] if not a.startswith("--json=")]
json_out = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--json=")), None)
for rep in args:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"== {rep}")
    for r in data:
        name = r[hdr.index("Kernel Name")]
        print(f"kernel: {name}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:95s} {r[i]:>18s} {units[i]}")
        if json_out:                                     # the figures bench.py quotes (first kernel of the first report)
            import json
            g = lambda k: float(r[hdr.index(k)].replace(",", "")) if k in hdr else None
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd, wr = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
            rd = rd * scale.get(units[hdr.index("dram__bytes_read.sum")], 1.0) if rd is not None else None
            wr = wr * scale.get(units[hdr.index("dram__bytes_write.sum")], 1.0) if wr is not None else None
            json.dump({"report": rep, "kernel": name, "dram_bytes_read": rd, "dram_bytes_write": wr,
                       "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                       "inst_executed": g("smsp__inst_executed.sum"), "duration": g("gpu__time_duration.sum"),
                       "duration_unit": units[hdr.index("gpu__time_duration.sum")]}, open(json_out, "w"), indent=1)
            json_out = None
