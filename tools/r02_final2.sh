#!/bin/bash
# Round 2, closing single-GPU pass: every GPU test, smoke, bench + reference arm (the ncu JSONs in profiles/ still describe this kernel source)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_final_pytest_gpu.log; tail -4 gpurun_out/r02_final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_final_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_final_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_final_bench_ref.json 2> gpurun_out/r02_final_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_final_bench.json"))
r = json.load(open("gpurun_out/r02_final_bench_ref.json"))
print("same config:", d["config"] == r["config"])
print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "parity_sample_bit_exact")}))
ro = d["roofline"]; print(json.dumps({k: ro[k] for k in ("bound", "achieved", "peak", "frac", "traffic", "algorithmic_frac")}))
print("c5", d["c5"]["value"], d["c5"]["roofline"]["frac"], d["c5"]["roofline"]["traffic"], d["c5"]["parity_sample_bit_exact"], d["c5"]["quantised_nodes"]["value"])
print("render", d["device_render"]["value"], "c3", d["c3"]["value"], "c4", d["c4"]["value"])
print("e2e", d["e2e"]["value"], d["e2e"]["frac_of_pcie_ceiling"], "ref", r["value"], r.get("reference_cpu_query_path"))
PY
