#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the handful of metrics the roofline uses. Usage: ncu_summary.py rep [out.txt]"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg"]
out = []
for r in rows[2:]:
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            out.append("%-70s %s %s" % (w, r[i], units[i]))
    out.append("-" * 100)
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write("# ncu --set full --clock-control none, summarised by tools/ncu_summary.py from %s\n" % rep + text + "\n")
