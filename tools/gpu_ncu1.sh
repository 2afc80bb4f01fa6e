#!/bin/bash
# ncu --set full capture of one kernel (regex $1) from the 32-frame bench command (+ optional env in $2)
mkdir -p gpurun_out
k=$1
env $2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o /tmp/prof_$k \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$k.txt 2>&1
ncu -i /tmp/prof_$k.ncu-rep --page raw --csv > gpurun_out/raw_$k.csv 2>/dev/null
ncu -i /tmp/prof_$k.ncu-rep --page source --csv > gpurun_out/src_$k.csv 2>/dev/null
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv > gpurun_out/smi_$k.txt
tail -2 gpurun_out/ncu_$k.txt | cut -c1-300
