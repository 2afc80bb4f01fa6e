#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x --timeout 600 -k "yuv or video_file or resize" > gpurun_out/pytest_yuv.txt 2>&1; tail -3 gpurun_out/pytest_yuv.txt
python tools/probe_yuv_e2e.py 2>&1 | tail -9
