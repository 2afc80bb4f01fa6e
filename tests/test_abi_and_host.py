"""No-GPU checks: the nvcc-built library loads and exports every symbol of include/cvvdp_b200.h, the
package refuses to run without a CUDA device, and the host-side helpers behave like the reference."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import colorvideovdp_b200 as cv
from colorvideovdp_b200 import _native as N

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "cvvdp_b200.h")).read()
    return sorted(set(re.findall(r"\b(cvvdp_b200_[a-z_]+)\s*\(", text)))


def test_header_and_binding_declare_the_same_symbols():
    assert _header_symbols() == sorted(N.SYMBOLS)


def test_cuda_library_loads_and_exports_every_symbol():
    if not os.path.isfile(N.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(N.LIB_PATH)
    for name in _header_symbols():
        assert hasattr(lib, name), name
    assert N.load_library().cvvdp_b200_abi_version() == N.ABI_VERSION


def test_struct_sizes_match_the_header():
    # sizes follow from the header's field lists (all 4-byte fields except the two int64)
    assert ctypes.sizeof(N.Params) == 4 * (2 + 4 + 16 + 4 + 1 + 2 + 1 + 2 + 4 + 1 + 4 + 4 + 1)
    assert ctypes.sizeof(N.CsfLut) == 4 * (32 + 32 + 4 * 32 * 32)
    assert ctypes.sizeof(N.Display) == 4 * (2 + 5 + 9 + 1)
    assert ctypes.sizeof(N.Clip) == 8 + 5 * 8 + 8
    assert ctypes.sizeof(N.Yuv) == 4 * 6
    assert ctypes.sizeof(N.Job) == 4 * 10 + 8 + 4 * 6 + 4 + 4  # ... + features + tail padding (8-byte alignment)
    assert ctypes.sizeof(N.PlanInfo) == 4 * 4 + 4 * 16 * 3 + 4 * 4 * 129 + 8


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cv.cvvdp(display_name="standard_4k")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cv.cvvdp(display_name="standard_4k", device=torch.device("cpu"))
    params, lut = cv.cvvdp_metric._default_native_inputs()
    with pytest.raises(N.NativeError, match="no CUDA device"):
        N.Context(params, lut, 0)  # the native library itself refuses as well


def test_metric_registry_and_names():
    assert cv.vq_metric_dict["cvvdp"] is cv.cvvdp
    assert issubclass(cv.cvvdp, cv.vq_metric)


def test_reshuffle_dims():
    t = torch.arange(2 * 3 * 4).reshape(2, 3, 4)
    out = cv.reshuffle_dims(t, "HWC", "BCFHW")
    assert tuple(out.shape) == (1, 4, 1, 2, 3)
    assert torch.equal(out[0, :, 0], t.permute(2, 0, 1))
    assert tuple(cv.reshuffle_dims(torch.zeros(5, 7), "HW", "BCFHW").shape) == (1, 1, 1, 5, 7)
    assert tuple(cv.reshuffle_dims(torch.zeros(6, 3, 5, 7), "FCHW", "BCFHW").shape) == (1, 3, 6, 5, 7)


def test_display_geometry_ppd():
    from oracle import cvvdp_oracle as O
    for name in ("standard_4k", "standard_fhd", "standard_hmd", "standard_phone", "iphone_12_pro"):
        g = cv.vvdp_display_geometry.load(name)
        assert abs(g.get_ppd() - O.Display(name).ppd) < 1e-9
    assert abs(cv.vvdp_display_geometry.load("standard_4k").get_ppd() - 75.40) < 0.01
    assert abs(cv.vvdp_display_geometry.load("standard_fhd").get_ppd() - 37.84) < 0.01


def test_config_paths_lookup(tmp_path):
    custom = tmp_path / "display_models.json"
    custom.write_text('{"my_disp": {"name": "x", "resolution": [100, 50], "viewing_distance_meters": 1, '
                      '"diagonal_size_inches": 10, "max_luminance": 123}}')
    dm = cv.vvdp_display_photometry.load("my_disp", [str(tmp_path)])
    assert dm.get_peak_luminance() == 123 and dm.contrast == 500
    with pytest.raises(RuntimeError):
        cv.config_files.find("nope.json", [])
    with pytest.raises(RuntimeError):
        cv.config_files.find("display_models.json", "not-a-list")


def test_info_string_format():
    from emu_util import emu_library
    from colorvideovdp_b200 import cvvdp_metric
    cvvdp_metric._set_mock_library_for_tests(emu_library())
    try:
        m = cv.cvvdp(display_name="standard_4k")
        assert m.get_info_string() == ('"ColorVideoVDP v0.5.6, 75.4 [pix/deg], Lpeak=200, Lblack=0.2, '
                                       'Lrefl=0.3979 [cd/m^2], (standard_4k)"')
        assert m.quality_unit() == "JOD" and m.short_name() == "cvvdp"
    finally:
        cvvdp_metric._set_mock_library_for_tests(None)


def test_two_stage_temporal_kernel_keeps_its_indices_in_registers():
    """Guard against a code-generation cliff seen in round 2: one more live value at the end of k_temporal_2s made ptxas
    rematerialise thread / block indices and addresses inside the frame loop (S2R, S2UR, LEA per frame) and cost 15-30 %
    of the kernel.  The healthy 17-tap table variant has a handful of S2R and about 3000 instructions."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump) or not os.path.isfile(N.LIB_PATH):
        pytest.skip("cuobjdump or the built library is not available")
    out = subprocess.run([cuobjdump, "-sass", "-fun", "_ZN5cvvdp13k_temporal_2sILi17ELi1EEEvNS_12TemporalArgsE", N.LIB_PATH],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    ops = [ln.split()[1 if not ln.split()[1].startswith("@") else 2] for ln in out.splitlines()
           if ln.strip().startswith("/*") and len(ln.split()) > 2 and ln.split()[0].endswith("*/")]
    if not ops:
        pytest.skip("kernel not found in the library (development build)")
    s2r = sum(1 for o in ops if o.startswith("S2R") or o.startswith("S2UR"))
    assert s2r <= 30, f"{s2r} S2R/S2UR in k_temporal_2s<17, LUT>: indices are being rematerialised in the frame loop"
    assert len(ops) <= 3200, f"k_temporal_2s<17, LUT> grew to {len(ops)} instructions"
