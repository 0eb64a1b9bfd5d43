#!/bin/bash
# what would a cheap re-binning buy on the L2-resident scene? traversal time of the bounce streams (8.7 M rays) unsorted
# and re-binned with 3 / 4 / 5 Morton bits per axis of the origin, per kernel (ncu launch list: cold, serialised)
mkdir -p gpurun_out
for cfg in "sort=0" "sort=1,sort_origin_bits=3" "sort=1,sort_origin_bits=4" "sort=1,sort_origin_bits=5" "sort=1,sort_origin_bits=4,sort_dir_bits=1" "sort=1,sort_origin_bits=3,sort_dir_bits=1"; do
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l23.csv python tools/prof_case.py $cfg 1 > /dev/null 2>&1
  python - "$cfg" <<'PY'
import csv, sys
rows = list(csv.reader(open("gpurun_out/l23.csv")))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hd = rows[h]
seq = []
for r in rows[h + 1:]:
    if len(r) == len(hd):
        d = dict(zip(hd, r))
        if d["Metric Name"] == "gpu__time_duration.sum":
            seq.append((d["Kernel Name"].split("(")[0].split("::")[-1][:28], float(d["Metric Value"]) / 1e3))
# the last two traversal launches: primary stream, bounce streams; everything between them belongs to the second
idx = [i for i, s in enumerate(seq) if "tracePackedKernel" in s[0]]
a, b = idx[-2], idx[-1]
pre = seq[a + 1:b]
print(sys.argv[1], "| primary trace %.0f us | bounce: sort kernels %.0f us (%d launches), trace %.0f us" % (seq[a][1], sum(t for _, t in pre), len(pre), seq[b][1]))
PY
done
