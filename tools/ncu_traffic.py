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
json.dump({"dram_bytes_per_launch": mean, "launches": len(vals), "kernel": vals[0][0], "gpu_time_ms_under_ncu": [v[2] for v in vals],
           "source": f"ncu --set full --clock-control none, {path} (tools/profile.sh), dram__bytes_read.sum + dram__bytes_write.sum"},
          open("profiles/ncu_bench_traffic.json", "w"), indent=1)
print(open("profiles/ncu_bench_traffic.json").read())
