"""Import the UNMODIFIED reference (gfxdisp/ColorVideoVDP, package `pycvvdp`) -- TEST INFRASTRUCTURE.

Two places are searched, in this order:
  1. `baseline/_ref/` in this repository: the reference installed there by `tools/stage_reference.sh`
     (`pip install --target`, git-ignored, NOT gpurun-ignored, so it travels to the GPU box).  This is what the
     `-m gpu` full-shape parity tests (reference on `cuda` as the JOD oracle of the BASELINE configurations) and
     `bench.py --impl reference` / `cpu_baseline` (reference on the host cores) import at run time.
  2. `/root/reference`: the read-only tree of the build container, used by `tests/golden/make_golden*.py`
     (fixture generation) and by the container-only cross-check tests.
Only `tests/`, `__graft_entry__.smoke()` and the reference legs of `bench.py` may import this module; the product
package never does.

`import pycvvdp` needs `ffmpeg` and `imageio` at module scope (pycvvdp/video_writer.py:2-3,
pycvvdp/video_source_file.py:8,12); neither is on the hot path, so empty stub modules are enough.
"""
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"
STAGED_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def staged():
    """The pip-installed copy under baseline/_ref exists (works on the GPU box)."""
    return os.path.isfile(os.path.join(STAGED_ROOT, "pycvvdp", "cvvdp_metric.py"))


def available():
    return staged() or os.path.isdir(os.path.join(REFERENCE_ROOT, "pycvvdp"))


def load(prefer_staged=True):
    if not available():
        raise RuntimeError("reference not present: run tools/stage_reference.sh in the build container")
    for name in ("ffmpeg", "imageio", "imageio.v2"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    root = STAGED_ROOT if (staged() and prefer_staged) else REFERENCE_ROOT
    if root not in sys.path:
        sys.path.insert(0, root)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import pycvvdp
    return pycvvdp


def reference_metric(display_name, device, heatmap=None, temp_padding="replicate"):
    """A `pycvvdp.cvvdp` object with true-fp32 convolutions (TF32 off), quiet."""
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    pycvvdp = load()
    import logging
    logging.getLogger().setLevel(logging.ERROR)
    return pycvvdp.cvvdp(display_name=display_name, heatmap=heatmap, quiet=True, device=torch.device(device),
                         temp_padding=temp_padding)
