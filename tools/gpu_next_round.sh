#!/bin/bash
# First GPU call of the next round (about 6 GPU-minutes on one B200): confirm the GPU tests added at the end
# of round 1 without a GPU (feature mode, fp32/fp16 two-stage temporal kernel), then time the three opt-in A/B
# variants against the default on the same box, then one ncu capture of the narrow k_band3.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -3 gpurun_out/pytest_gpu.txt
./tools/gpu_quick.sh "" "CVVDP_B200_BAND3_NARROW=1" "CVVDP_B200_REDUCE_TY16=1" "CVVDP_B200_BAND3_NARROW=1 CVVDP_B200_REDUCE_TY16=1"
CVVDP_B200_BAND3_NARROW=1 ./tools/gpu_ncu1.sh k_band3 CVVDP_B200_BAND3_NARROW=1
