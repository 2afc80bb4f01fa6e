#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
for k in k_band2 k_temporal_reg; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/prof_$k \
      python bench.py --frames 32 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$k.txt 2>&1
done
tail -3 gpurun_out/pytest_gpu.txt
