#!/bin/bash
# One GPU call: smoke, GPU parity tests (incl. the staged reference on cuda), the default bench line (both arms,
# extras), optionally the ncu launch list and full captures of the dominant kernels ("ncu" as first argument).
mkdir -p gpurun_out
timeout 180 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/bench.txt 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
if [ "$1" = "ref" ] || [ "$1" = "ncu" ]; then
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.txt 2> gpurun_out/bench_ref.err; echo "ref rc=$?" >> gpurun_out/bench_ref.err
fi
if [ "$1" = "ncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/launches_bench.log 2>&1
for k in k_band2 k_temporal_2s k_reduce2; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o /tmp/prof_$k \
      python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/ncu_$k.txt 2>&1
  ncu -i /tmp/prof_$k.ncu-rep --page raw --csv > gpurun_out/raw_$k.csv 2>/dev/null
  ncu -i /tmp/prof_$k.ncu-rep --page source --csv > gpurun_out/src_$k.csv 2>/dev/null
done
fi
du -sh gpurun_out
tail -n 3 gpurun_out/smoke.txt gpurun_out/pytest_gpu.txt gpurun_out/bench.err
tail -c 1500 gpurun_out/bench.txt
