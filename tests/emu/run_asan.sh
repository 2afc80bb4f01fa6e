#!/bin/sh
# TEST INFRASTRUCTURE: the mock-device build under AddressSanitizer -- every global, "shared" (heap / static
# arrays of the mock) and staging access of the kernels is bounds-checked while the emu parity tests run.
#   sh tests/emu/run_asan.sh [pytest args, default: tests/test_emu_parity.py -q]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
LIB="$HERE/libcvvdp_b200_emu.so"
[ -f "$LIB" ] && cp "$LIB" "$LIB.plain"
g++ -O1 -g -fsanitize=address -fno-omit-frame-pointer -std=c++17 -fPIC -shared -DCVVDP_EMU -I"$HERE" -x c++ \
    "$ROOT/colorvideovdp_b200/csrc/cvvdp_api.cu" -o "$LIB" -Wno-unused-function -Wno-unknown-pragmas
cd "$ROOT"
status=0
LD_PRELOAD="$(g++ -print-file-name=libasan.so)" ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:halt_on_error=1 \
    python -m pytest ${@:-tests/test_emu_parity.py -q} || status=$?
if [ -f "$LIB.plain" ]; then mv "$LIB.plain" "$LIB"; touch "$LIB"; else rm -f "$LIB"; fi
exit $status
