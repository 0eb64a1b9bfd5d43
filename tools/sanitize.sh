#!/bin/bash
# compute-sanitizer passes over tools/sanitize_case.py (run under gpurun, one GPU). Logs -> gpurun_out/.
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_case.py 6000 > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize case ok|Error|error" gpurun_out/sanitize_$tool.log | head -8
done
