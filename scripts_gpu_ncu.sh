#!/bin/bash
# ncu evidence: launch list of the bench command + full captures of the top kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.txt 2> gpurun_out/ncu_list.err
for k in k_band k_temporal k_reduce; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/prof_$k \
      python bench.py --frames 32 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$k.txt 2>&1
done
timeout 600 python bench.py --dtype f32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_f32.txt 2> gpurun_out/bench_f32.err
ls -la gpurun_out
