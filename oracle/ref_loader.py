"""Import the UNMODIFIED reference (pycvvdp) from /root/reference -- build-container only.

TEST INFRASTRUCTURE.  /root/reference does not exist on the GPU box, so nothing in the `-m gpu`
tests, smoke() or bench.py may call this; it is used by tests/golden/make_golden.py (fixture
generation) and by the container-only cross-check tests (skipped when the tree is absent).

`import pycvvdp` needs `ffmpeg` and `imageio` at module scope (pycvvdp/video_writer.py:2-3,
pycvvdp/video_source_file.py:8,12); neither is on the hot path, so empty stub modules are enough.
"""
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pycvvdp"))


def load():
    if not available():
        raise RuntimeError("reference tree not present (expected on the build container only)")
    for name in ("ffmpeg", "imageio", "imageio.v2"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import pycvvdp
    return pycvvdp
