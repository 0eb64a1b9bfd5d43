#!/bin/bash
# Round 2, GPU call 3: quantised nodes (variant 4) on hardware -- parity, then what they buy on battlefield and on config 5,
# with ncu beside the exact format; the full GPU suite and the bench line with the new records.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02c3_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02c3_pytest_gpu.log
tail -12 gpurun_out/r02c3_pytest_gpu.log
timeout 300 python tools/sweep.py "variant=3;variant=4;variant=4,smem_stack=16;variant=3,sort=1;variant=4,sort=1" > gpurun_out/r02c3_sweep.log 2>&1; grep -E "^\{" gpurun_out/r02c3_sweep.log
timeout 300 python tools/prof_c5.py > gpurun_out/r02c3_c5_times.txt 2>&1; cat gpurun_out/r02c3_c5_times.txt
# ncu: battlefield bench batch, exact vs quantised (prof_case launches primary and secondary separately, twice)
for v in 3 4; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:tracePackedKernel -s 2 -c 2 -o gpurun_out/r02c3_prof_bf_v$v -f python tools/prof_case.py variant=$v 2 > gpurun_out/r02c3_prof_bf_v$v.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tracePackedKernel -s 1 -c 2 -o gpurun_out/r02c3_prof_c5 -f python tools/prof_c5.py > gpurun_out/r02c3_prof_c5.log 2>&1
for f in r02c3_prof_bf_v3 r02c3_prof_bf_v4 r02c3_prof_c5; do python tools/ncu_summary.py gpurun_out/$f.ncu-rep > gpurun_out/$f.txt 2>&1; done
grep -E "kernel|gpu__time_duration|data_pipe_lsu_wavefronts.sum|issue_active|pipe_alu.avg|dram__bytes_read|l1tex__throughput|lts__t_sector_hit|inst_executed.sum" gpurun_out/r02c3_prof_*.txt | cut -c1-200
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02c3_bench.json 2> gpurun_out/r02c3_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02c3_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02c3_bench.json"))
print(json.dumps({k: d[k] for k in ("value", "breakdown")})[:1500])
print(json.dumps(d.get("c5"))[:2500])
PY
