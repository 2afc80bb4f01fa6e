#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_temporal_2s -c 1 -f -o /tmp/prof_tyuv python tools/probe_yuv_e2e.py > gpurun_out/ncu_tyuv.txt 2>&1
ncu -i /tmp/prof_tyuv.ncu-rep --page raw --csv > gpurun_out/raw_tyuv.csv 2>/dev/null
ncu -i /tmp/prof_tyuv.ncu-rep --page source --csv > gpurun_out/src_tyuv.csv 2>/dev/null
ncu -i /tmp/prof_tyuv.ncu-rep --page details > gpurun_out/details_tyuv.txt 2>/dev/null
tail -5 gpurun_out/ncu_tyuv.txt
