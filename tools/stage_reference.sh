#!/bin/bash
# Installs the UNMODIFIED reference (gfxdisp/ColorVideoVDP, package `pycvvdp`) into the git-ignored
# baseline/_ref/ with pip, so that it travels to the GPU box with the gpurun snapshot and can serve as
#   * the JOD / Q_per_ch oracle at the BASELINE shapes (tests/test_gpu_reference.py, device='cuda'), and
#   * the CPU arm of bench.py (--impl reference, cpu_baseline.kind = "reference").
# /root/reference is read-only, so the build runs from a scratch copy.  No media, no examples: the
# package data are the calibration JSON files only.  Nothing from the reference enters the git history.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference}"
if [ ! -d "$SRC/pycvvdp" ]; then echo "reference tree not found at $SRC" >&2; exit 1; fi
TMP="$(mktemp -d /tmp/cvvdp_ref_XXXXXX)"
mkdir -p "$TMP/src"
cp -r "$SRC/pycvvdp" "$SRC/pyproject.toml" "$SRC/README.md" "$SRC/LICENSE" "$TMP/src/"
rm -rf "$ROOT/baseline/_ref"
mkdir -p "$ROOT/baseline/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$ROOT/baseline/_ref" "$TMP/src" 2>&1 | tail -3
rm -rf "$TMP"
ls "$ROOT/baseline/_ref"
