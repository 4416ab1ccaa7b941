"""The C++ drop-in class ORB_SLAM3::ORBextractor (morb_slam_b200/cpp) exercised through a small C++
driver: same constructor / operator() / getters / mvImagePyramid as the reference header, results
compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

from morb_slam_b200 import synth
from oracle import oracle_py as op
from tests.conftest import ROOT, has_cuda

DRIVER = os.path.join(ROOT, "tests", "cpp", "dropin_driver")


def build_driver():
    subprocess.run(["make", "-s", "-j4", "-C", os.path.join(ROOT, "morb_slam_b200", "csrc")], check=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "oracle", "shim"), "-I" + os.path.join(ROOT, "include"),
                    "-I" + os.path.join(ROOT, "morb_slam_b200", "cpp"), os.path.join(ROOT, "tests", "cpp", "dropin_driver.cc"),
                    os.path.join(ROOT, "morb_slam_b200", "cpp", "ORBextractor.cc"), "-L" + os.path.join(ROOT, "morb_slam_b200", "lib"),
                    "-lorb_b200", "-Wl,-rpath," + os.path.join(ROOT, "morb_slam_b200", "lib"), "-o", DRIVER], check=True)


def test_dropin_compiles_against_opencv_style_headers():
    build_driver()
    assert os.path.exists(DRIVER)


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
@pytest.mark.parametrize("cfg,seed", [("euroc", 2000), ("tumvi", 3000)])
def test_dropin_matches_oracle(tmp_path, cfg, seed):
    build_driver()
    w, h, nf, lap, fx, b = synth.CONFIGS[cfg]
    L, R = synth.stereo_pair(seed, w, h)
    L.tofile(tmp_path / "l.raw"); R.tofile(tmp_path / "r.raw")
    mbf, maxD = float(np.float32(fx * b)), float(np.float32(fx))
    out = tmp_path / "out.bin"
    r = subprocess.run([DRIVER, str(w), str(h), str(nf), str(lap[0]), str(lap[1]), str(tmp_path / "l.raw"), str(tmp_path / "r.raw"),
                        str(out), repr(mbf), repr(maxD)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    buf = out.read_bytes()
    mono, n = np.frombuffer(buf, np.int32, 2)
    o = 8
    kps = np.frombuffer(buf, op.KP_DTYPE, n, o); o += 28 * n
    desc = np.frombuffer(buf, np.uint8, 32 * n, o).reshape(n, 32); o += 32 * n
    oL, oR = op.OracleExtractor(nf), op.OracleExtractor(nf)
    mo, ko, do = oL(L, lap)
    assert mono == mo and kps.tobytes() == ko.tobytes() and np.array_equal(desc, do)
    for l in range(8):
        lw, lh = np.frombuffer(buf, np.int32, 2, o); o += 8
        lvl = np.frombuffer(buf, np.uint8, lw * lh, o).reshape(lh, lw); o += lw * lh
        assert np.array_equal(lvl, oL.level(l)), l      # mvImagePyramid
    _, kR, dR = oR(R, lap)
    u = np.frombuffer(buf, np.float32, n, o); o += 4 * n
    d = np.frombuffer(buf, np.float32, n, o)
    u_o, d_o = op.oracle_stereo(oL, oR, ko, do, kR, dR, mbf, maxD)
    assert u.tobytes() == u_o.tobytes() and d.tobytes() == d_o.tobytes()
