#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.txt 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
CVVDP_B200_DEBUG_TIMELINE=1 python tools/e2e_probe.py > gpurun_out/e2e_probe.txt 2>&1
tail -n 3 gpurun_out/pytest_gpu.txt; tail -n 3 gpurun_out/bench.err; tail -n 12 gpurun_out/e2e_probe.txt
