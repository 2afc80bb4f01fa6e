"""Load the CPU mock-device build of the native library (tests/emu) and inject it into the package."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_LIB = os.path.join(HERE, "emu", "libcvvdp_b200_emu.so")
_state = {"lib": None}


def emu_library():
    if _state["lib"] is None:
        if not os.path.isfile(EMU_LIB):
            if shutil.which("g++") is None:
                pytest.skip("g++ not available to build the mock-device library")
            subprocess.run(["sh", os.path.join(HERE, "emu", "build_emu.sh")], check=True)
        from colorvideovdp_b200 import _native
        _state["lib"] = _native.load_library(EMU_LIB)
    return _state["lib"]


@pytest.fixture
def mock_device():
    """Metric objects created inside the test run on the mock device (CPU tensors)."""
    from colorvideovdp_b200 import cvvdp_metric
    cvvdp_metric._set_mock_library_for_tests(emu_library())
    yield
    cvvdp_metric._set_mock_library_for_tests(None)
