"""Writes profiles/<name>.json (default ncu_bench_traffic) from an `ncu --set full` capture (tools/profile.sh ->
gpurun_out/prof.ncu-rep): DRAM read+write bytes per launch of the traversal kernel, its limiter percentages, and the
hash of the kernel source it was captured from -- bench.py quotes the file only while that hash matches its own build.
usage: ncu_traffic.py [report.ncu-rep [name [kernel-regex]]]"""
import csv
import json
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import kernel_source_sha  # noqa: E402

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/prof.ncu-rep"
name = sys.argv[2] if len(sys.argv) > 2 else "ncu_bench_traffic"
only = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
vals = []
body = [r for r in rows[2:] if only is None or only.search(r[h.index("Kernel Name")])]
for r in body:
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = h.index(m)
        tot += float(r[i]) * scale[units[i]]
    vals.append((r[h.index("Kernel Name")], tot, float(r[h.index("gpu__time_duration.sum")])))
mean = sum(v[1] for v in vals) / len(vals)
limiters = {}
for metric, key in (("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1_data_pipe_pct_of_peak"),
                  ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_slots_pct_of_peak"),
                  ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pipe_pct_of_peak"),
                  ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_peak"),
                  ("sm__warps_active.avg.pct_of_peak_sustained_active", "active_warps_pct_of_peak"),
                  ("smsp__thread_inst_executed_per_inst_executed.ratio", "active_lanes_per_instruction")):
    if metric in h:
        limiters[key] = round(sum(float(r[h.index(metric)]) for r in body) / len(body), 2)
wavefronts = None
if "l1tex__data_pipe_lsu_wavefronts.sum" in h:
    wavefronts = sum(float(r[h.index("l1tex__data_pipe_lsu_wavefronts.sum")]) for r in body) / len(body)
json.dump({"dram_bytes_per_launch": mean, "l1_wavefronts_per_launch": wavefronts, "launches": len(vals), "kernel": vals[0][0], "gpu_time_ms_under_ncu": [v[2] for v in vals],
           "limiters": limiters, "kernel_source_sha16": kernel_source_sha(),
           "source": f"ncu --set full --clock-control none, {path} (tools/profile.sh), dram__bytes_read.sum + dram__bytes_write.sum"},
          open(f"profiles/{name}.json", "w"), indent=1)
print(open(f"profiles/{name}.json").read())
