"""The C++ drop-in class ORB_SLAM3::ORBextractor (morb_slam_b200/cpp) exercised through a small C++
driver: same constructor / operator() / getters / mvImagePyramid as the reference header, results
compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

from morb_slam_b200 import capi, synth
from oracle import oracle_py as op
from tests.conftest import ROOT, has_cuda

DRIVER = os.path.join(ROOT, "tests", "cpp", "dropin_driver")


def build_driver():
    subprocess.run(["make", "-s", "-j4", "-C", os.path.join(ROOT, "morb_slam_b200", "csrc")], check=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I" + os.path.join(ROOT, "oracle", "shim"), "-I" + os.path.join(ROOT, "include"),
                    "-I" + os.path.join(ROOT, "morb_slam_b200", "cpp"), os.path.join(ROOT, "tests", "cpp", "dropin_driver.cc"),
                    os.path.join(ROOT, "morb_slam_b200", "cpp", "ORBextractor.cc"), "-L" + os.path.join(ROOT, "morb_slam_b200", "lib"),
                    "-lorb_b200", "-Wl,-rpath," + os.path.join(ROOT, "morb_slam_b200", "lib"), "-o", DRIVER], check=True)


def test_dropin_compiles_against_opencv_style_headers():
    build_driver()
    assert os.path.exists(DRIVER)


@pytest.mark.parametrize("nf", [1000, 1200, 1500, 2000])
def test_dropin_tables_before_first_extraction(nf):
    """the reference fills mvScaleFactor / mvInvScaleFactor / mvLevelSigma2 / mvInvLevelSigma2 in its constructor and every Frame
    constructor copies them before ExtractORB (src/Frame.cc:181-187 vs :194): the drop-in's getters must return the same tables
    right after construction, with no device and no extraction."""
    build_driver()
    r = subprocess.run([DRIVER, "--tables", str(nf)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0].split()[0] == "8" and np.float32(lines[0].split()[1]) == np.float32(1.2)
    t = op.OracleExtractor(nf).tables()
    for line, key in zip(lines[1:], ("scale", "inv_scale", "sigma2", "inv_sigma2")):
        got = np.array(line.split(), np.float32)
        assert got.tobytes() == np.asarray(t[key], np.float32).tobytes(), key
    sc, inv, s2, is2, nfl = capi.compute_tables(nf)
    assert np.array_equal(nfl, t["nfeat"]) and sc.tobytes() == np.asarray(t["scale"], np.float32).tobytes()


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


FRAME_DRIVER = os.path.join(ROOT, "tests", "cpp", "frame_driver")


def build_frame_driver():
    subprocess.run(["make", "-s", "-j4", "-C", os.path.join(ROOT, "morb_slam_b200", "csrc")], check=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I" + os.path.join(ROOT, "oracle", "shim"), "-I" + os.path.join(ROOT, "include"),
                    "-I" + os.path.join(ROOT, "morb_slam_b200", "cpp"), os.path.join(ROOT, "tests", "cpp", "frame_driver.cc"),
                    os.path.join(ROOT, "morb_slam_b200", "cpp", "ORBextractor.cc"), "-L" + os.path.join(ROOT, "morb_slam_b200", "lib"),
                    "-lorb_b200", "-Wl,-rpath," + os.path.join(ROOT, "morb_slam_b200", "lib"), "-o", FRAME_DRIVER], check=True)


def test_frame_helpers_compile_against_opencv_style_headers():
    build_frame_driver()
    assert os.path.exists(FRAME_DRIVER)


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
def test_frame_helpers_match_oracle(tmp_path):
    """FrameB200.h (UndistortKeyPointsB200, AssignFeaturesToGridB200, ComputeBoWB200 with std::map types, vocabulary loaded from the
    ORBvoc.txt text format) through a C++ driver against the oracle."""
    import ctypes as C
    from oracle import oracle_bow_py as ob
    build_frame_driver()
    w, h, nf, lap, fx, b = synth.CONFIGS["euroc"]
    img = synth.mono_frame(8100, w, h)
    img.tofile(tmp_path / "i.raw")
    voc = synth.synth_vocabulary(81, 10, 3, 0.03, 0.1)
    synth.write_vocabulary_text(voc, str(tmp_path / "voc.txt"))
    out = tmp_path / "out.bin"
    r = subprocess.run([FRAME_DRIVER, str(w), str(h), str(nf), str(tmp_path / "i.raw"), str(tmp_path / "voc.txt"), "2", str(out)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = out.read_bytes()
    n = int(np.frombuffer(raw, np.int32, 1, 0)[0]); off = 4
    un = np.frombuffer(raw, op.KP_DTYPE, n, off); off += n * 28
    nb = int(np.frombuffer(raw, np.int32, 1, off)[0]); off += 4
    bow = np.frombuffer(raw, np.dtype([("w", "<u4"), ("v", "<f8")]), nb, off); off += nb * 12
    nn = int(np.frombuffer(raw, np.int32, 1, off)[0]); off += 4
    nodes, feats = [], []
    for _ in range(nn):
        node, c = np.frombuffer(raw, np.uint32, 1, off)[0], int(np.frombuffer(raw, np.int32, 1, off + 4)[0]); off += 8
        nodes.append(node); feats.append(np.frombuffer(raw, np.uint32, c, off)); off += 4 * c
    mo, ko, do = op.OracleExtractor(nf)(img, lap)
    assert n == len(ko)
    lib = op.oracle_lib()
    lib.shim_undistort_points.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    K = np.array([458.654, 0, 367.215, 0, 457.296, 248.375, 0, 0, 1], np.float32)
    D = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05], np.float32)
    pts = np.ascontiguousarray(np.stack([ko["x"], ko["y"]], 1), np.float32)
    und = np.zeros_like(pts)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    lib.shim_undistort_points(p(pts), n, p(K), p(D), 4, p(K), p(und))
    want = ko.copy(); want["x"], want["y"] = und[:, 0], und[:, 1]
    assert un.tobytes() == want.tobytes()
    wb = ob.OracleVocabulary(voc).transform(do, 2)
    assert np.array_equal(bow["w"], wb["bow_word"]) and bow["v"].tobytes() == wb["bow_val"].tobytes()
    assert np.array_equal(np.array(nodes, np.uint32), wb["fv_node"])
    assert np.array_equal(np.concatenate(feats) if feats else np.zeros(0, np.uint32), wb["fv_feat"])
    # SearchForInitializationB200: the frame (undistorted keypoints, grid built on them) against itself, windows moved by (3, -2)
    from oracle import oracle_map_py as omap
    from oracle import oracle_match_py as om
    n_ini = int(np.frombuffer(raw, np.int32, 1, off)[0]); off += 4
    m12 = np.frombuffer(raw, np.int32, n, off); off += 4 * n
    prev = np.frombuffer(raw, np.float32, 2 * n, off).reshape(n, 2)
    p0 = np.stack([want["x"] + np.float32(3), want["y"] - np.float32(2)], 1).astype(np.float32)
    onm, om12, oprev = omap.search_for_initialization(want, do, p0, want, do, om.grid_params(w, h), 30, 0.9, True)
    assert n_ini == onm and np.array_equal(m12, om12) and prev.tobytes() == oprev.tobytes() and onm > 100
