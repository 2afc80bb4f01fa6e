#!/bin/sh
# TEST INFRASTRUCTURE: compile the CUDA sources of the product against the CPU mock of the CUDA
# runtime (tests/emu/cuda_emu.h) so that kernel logic can be checked without a GPU.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
g++ -O2 -g -std=c++17 -fPIC -shared -DCVVDP_EMU ${CVVDP_EMU_DEFS} -I"$HERE" -x c++ \
    "$ROOT/colorvideovdp_b200/csrc/cvvdp_api.cu" -o "$HERE/libcvvdp_b200_emu.so" -Wall -Wno-unused-function -Wno-unknown-pragmas
