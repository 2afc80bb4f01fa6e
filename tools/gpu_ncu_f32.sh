mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_temporal_2s -c 1 -f -o /tmp/prof_t32 python bench.py --dtype f32 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/ncu_t32.txt 2>&1
ncu -i /tmp/prof_t32.ncu-rep --page raw --csv > gpurun_out/raw_t32.csv 2>/dev/null
ncu -i /tmp/prof_t32.ncu-rep --page source --csv > gpurun_out/src_t32.csv 2>/dev/null
ncu -i /tmp/prof_t32.ncu-rep --page details > gpurun_out/details_t32.txt 2>/dev/null
python bench.py --dtype f32 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | cut -c1-1500
