#!/bin/bash
mkdir -p gpurun_out
CVVDP_B200_DEBUG_TIMELINE=1 python tools/e2e_probe.py > gpurun_out/e2e_probe.txt 2>&1
tail -22 gpurun_out/e2e_probe.txt
