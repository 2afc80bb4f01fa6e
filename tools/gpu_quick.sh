#!/bin/bash
# Quick A/B on one GPU: device-resident bench only, once per environment setting given as arguments
# (e.g. ./tools/gpu_quick.sh "" "CVVDP_B200_NO_LOCKSTEP=1").
mkdir -p gpurun_out
: > gpurun_out/quick.txt
timeout 120 python __graft_entry__.py --smoke >> gpurun_out/quick.txt 2>&1
for envs in "$@"; do
  echo "=== env: [$envs]" >> gpurun_out/quick.txt
  env $envs timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras >> gpurun_out/quick.txt 2>> gpurun_out/quick.err
done
python - <<'PY'
import json
for line in open("gpurun_out/quick.txt"):
    if line.startswith("=== env") or line.startswith("smoke"):
        print(line.strip())
    elif line.startswith("{"):
        b = json.loads(line)
        print(" value", b["value"], "ms", b["ms_per_step"], {k: v["ms_per_step"] for k, v in list(b["kernels"].items())[:6]})
PY
tail -3 gpurun_out/quick.err
