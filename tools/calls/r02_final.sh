#!/bin/bash
# Round 2, final single-GPU pass: every GPU test, smoke, the bench line and the reference arm, then the ncu evidence the
# bench line quotes (tools/profile.sh). Outputs -> gpurun_out/r02_final_*.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_final_pytest_gpu.log; tail -4 gpurun_out/r02_final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r02_final_smoke.log
bash tools/profile.sh > gpurun_out/r02_final_profile.log 2>&1; tail -3 gpurun_out/r02_final_profile.log
cp gpurun_out/ncu_bench_traffic.json gpurun_out/ncu_c5_traffic.json profiles/ 2>/dev/null
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_final_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_final_bench_ref.json 2> gpurun_out/r02_final_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_final_bench.json"))
r = json.load(open("gpurun_out/r02_final_bench_ref.json"))
print("same config:", d["config"] == r["config"], d["config"])
print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "parity_sample_bit_exact")}))
ro = d["roofline"]; print(json.dumps({k: ro[k] for k in ("bound", "achieved", "peak", "frac", "traffic", "algorithmic_frac", "l1_wavefronts_per_launch", "ncu_limiters", "traffic_source")})[:1200])
print(json.dumps(d["c5"]["roofline"])[:700])
print("e2e", d["e2e"]["value"], d["e2e"]["frac_of_pcie_ceiling"], "ref", r["value"])
PY
