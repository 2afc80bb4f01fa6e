#!/bin/bash
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.txt
if grep -q "smoke OK" gpurun_out/smoke.txt; then
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.txt 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
if [ "$1" = "ncu" ]; then
for k in k_band2 k_temporal_stg k_reduce; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o /tmp/prof_$k \
      python bench.py --frames 32 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$k.txt 2>&1
  ncu -i /tmp/prof_$k.ncu-rep --page raw --csv > gpurun_out/raw_$k.csv 2>/dev/null
  ncu -i /tmp/prof_$k.ncu-rep --page source --csv > gpurun_out/src_$k.csv 2>/dev/null
  ls -la /tmp/prof_$k.ncu-rep >> gpurun_out/ncu_$k.txt
done
fi
fi
du -sh gpurun_out
tail -n 2 gpurun_out/smoke.txt gpurun_out/pytest_gpu.txt gpurun_out/bench.err
