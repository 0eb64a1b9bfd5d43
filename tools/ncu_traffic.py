"""Writes profiles/ncu_bench_traffic.json from an `ncu --set full` capture of the bench command
(tools/profile.sh -> gpurun_out/prof.ncu-rep): DRAM read+write bytes per launch of the traversal kernel."""
import csv
import json
import subprocess
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/prof.ncu-rep"
out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
vals = []
for r in rows[2:]:
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = h.index(m)
        tot += float(r[i]) * scale[units[i]]
    vals.append((r[h.index("Kernel Name")], tot, float(r[h.index("gpu__time_duration.sum")])))
mean = sum(v[1] for v in vals) / len(vals)
limiters = {}
for name, key in (("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1_data_pipe_pct_of_peak"),
                  ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_slots_pct_of_peak"),
                  ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pipe_pct_of_peak"),
                  ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_peak"),
                  ("sm__warps_active.avg.pct_of_peak_sustained_active", "active_warps_pct_of_peak"),
                  ("smsp__thread_inst_executed_per_inst_executed.ratio", "active_lanes_per_instruction")):
    if name in h:
        limiters[key] = round(sum(float(r[h.index(name)]) for r in rows[2:]) / len(rows[2:]), 2)
json.dump({"dram_bytes_per_launch": mean, "launches": len(vals), "kernel": vals[0][0], "gpu_time_ms_under_ncu": [v[2] for v in vals],
           "limiters": limiters,
           "source": f"ncu --set full --clock-control none, {path} (tools/profile.sh), dram__bytes_read.sum + dram__bytes_write.sum"},
          open("profiles/ncu_bench_traffic.json", "w"), indent=1)
print(open("profiles/ncu_bench_traffic.json").read())
