"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every function that
include/orb_b200.h declares; compute entry points fail loudly without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from morb_slam_b200 import capi
from tests.conftest import ROOT, has_cuda


@pytest.fixture(scope="module", autouse=True)
def _build_lib():
    subprocess.run(["make", "-s", "-j4", "-C", os.path.join(ROOT, "morb_slam_b200", "csrc")], check=True)


def _declared():
    src = open(os.path.join(ROOT, "include", "orb_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(orb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(capi.LIB_PATH)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n


def test_host_helpers_need_no_device():
    a = np.zeros(32, np.uint8); b = np.full(32, 255, np.uint8)
    assert capi.descriptor_distance(a, b) == 256
    assert capi.descriptor_distance(a, a) == 0
    assert capi.lib().orb_status_string(-5).decode() == "capacity exceeded"


@pytest.mark.skipif(has_cuda(), reason="only meaningful without a GPU")
def test_no_cpu_fallback_without_device():
    with pytest.raises(capi.OrbError) as e:
        capi.ORBextractor(1000)
    assert e.value.status == -3   # ORB_ERR_CUDA


def test_keypoint_record_layout_is_cv_keypoint():
    assert capi.KP_DTYPE.itemsize == 28
    assert [capi.KP_DTYPE.fields[n][1] for n in ("x", "y", "size", "angle", "response", "octave", "class_id")] == [0, 4, 8, 12, 16, 20, 24]


def test_header_is_plain_c(tmp_path):
    """include/orb_b200.h is a C ABI: it must compile as C99 on its own (cgo / JNI / ctypes generators read it as C)."""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "orb_b200.h"\nint main(void) { orb_grid_params g; orb_bow_out o; orb_bow_keyframes k; (void)g; (void)o; (void)k; return 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I" + os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
