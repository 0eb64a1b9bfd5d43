#!/bin/bash
# GPU call 6 (1 GPU): L2 prefetch of pushed far children on config 5 (exact and quantised nodes), parity of the option
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "prefetch or larger_than_l2 or quantised_nodes_full" > gpurun_out/r02c6_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02c6_pytest.log
timeout 300 python tools/prof_c5.py > gpurun_out/r02c6_c5_times.txt 2>&1; cat gpurun_out/r02c6_c5_times.txt
PROF_C5_PREFETCH=1 timeout 300 python tools/prof_c5.py > gpurun_out/r02c6_c5_times_prefetch.txt 2>&1; cat gpurun_out/r02c6_c5_times_prefetch.txt
RACC_B200_SORT=0 timeout 300 python tools/prof_c5.py 10000000 30000000 > gpurun_out/r02c6_c5_times_unsorted.txt 2>&1; cat gpurun_out/r02c6_c5_times_unsorted.txt
RACC_B200_SORT=0 PROF_C5_PREFETCH=1 timeout 300 python tools/prof_c5.py 10000000 30000000 > gpurun_out/r02c6_c5_times_unsorted_prefetch.txt 2>&1; cat gpurun_out/r02c6_c5_times_unsorted_prefetch.txt
