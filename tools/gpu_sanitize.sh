#!/bin/bash
# compute-sanitizer on a B200: memcheck, racecheck (shared-memory hazards incl. the TMA stages) and synccheck over
# tools/sanitize_cases.py; summaries land in gpurun_out/ (copy the tails into profiles/).
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "rc=$?" >> gpurun_out/sanitizer_$tool.txt
  tail -n 6 gpurun_out/sanitizer_$tool.txt
done
