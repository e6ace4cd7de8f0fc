#!/usr/bin/env python
"""Sum DRAM traffic and duration over the tensor-core contraction launches of ONE bench step from an `ncu --set full` report
(tools/run_ncu_profiles.sh; launches_per_step 0 = autodetect the period) -> profiles/r0N_contraction_traffic.json, which bench.py reports as roofline.traffic."""
import csv, json, subprocess, sys
rep, launches_per_step, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {n: hdr.index(n) for n in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
names = [r[col["Kernel Name"]] for r in rows[2:]]
if launches_per_step <= 0:
    # autodetect: the smallest period of the kernel-name sequence (every step launches the same contraction kernels in the same order)
    launches_per_step = next((P for P in range(8, len(names) // 2 + 1) if names[:P] == names[P:2 * P]), len(names))
body = rows[2:][:launches_per_step]
tot_b = tot_ms = 0.0
per = {}
for r in body:
    b = (float(r[col["dram__bytes_read.sum"]]) * scale[units[col["dram__bytes_read.sum"]]] + float(r[col["dram__bytes_write.sum"]]) * scale[units[col["dram__bytes_write.sum"]]])
    ms = float(r[col["gpu__time_duration.sum"]]) * tscale[units[col["gpu__time_duration.sum"]]]
    tot_b += b
    tot_ms += ms
    k = r[col["Kernel Name"]].split("(")[0]
    e = per.setdefault(k, [0, 0.0, 0.0, 0.0])
    e[0] += 1; e[1] += b; e[2] += ms; e[3] += float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]) * ms
res = {"source": rep, "launches": len(body), "dram_bytes_per_step": tot_b, "ms_under_ncu": tot_ms,
       "per_kernel": {k: {"launches": v[0], "dram_bytes": v[1], "ms": v[2], "tensor_pipe_active_pct_time_weighted": v[3] / v[2] if v[2] else 0} for k, v in per.items()}}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k: res[k] for k in ("launches", "dram_bytes_per_step", "ms_under_ncu")}))
