"""Input rectification on the GPU (orb_set_rectify_maps + ORB_INPUT_REMAP: System::TrackStereo's cv::remap,
src/System.cc:254-261) against the oracle's restatement of cv::remap (oracle/shim remap_linear_8u, pinned against cv2 by
tests/test_oracle_primitives.py): rectified image bytes and everything extracted from it are bit-exact."""
import ctypes as C

import numpy as np
import pytest

from morb_slam_b200 import capi, synth
from oracle import oracle_py as op

pytestmark = pytest.mark.gpu


def oracle_remap(src, mx, my):
    lib = op.oracle_lib()
    lib.shim_remap.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    out = np.zeros(mx.shape, np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    lib.shim_remap(p(src), src.shape[1], src.shape[0], src.strides[0], p(mx), p(my), mx.shape[1], mx.shape[0], p(out))
    return out


@pytest.mark.parametrize("raw,rect", [((752, 480), (752, 480)), ((800, 520), (752, 480)), ((640, 400), (701, 443)), ((641, 401), (600, 380))])
def test_rectified_extraction_equals_oracle(raw, rect):
    op.build()
    (rw, rh), (w, h) = raw, rect
    nf, lap = 1200, (0, 0)
    B = 10                                     # more than one frame chunk of the kernel
    raws = np.stack([synth.mono_frame(6100 + i, rw, rh) for i in range(B)])
    mx, my = synth.rectify_maps(w, h, rw, rh, seed=3)
    assert (mx < 0).any() or (my < 0).any() or (mx > rw - 1).any() or (my > rh - 1).any()      # some pixels leave the raw image
    ex = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B)
    ex.set_rectify_maps(mx, my)
    n, mono, kps, desc = ex.extract_batch(raws, lap, flags=capi.ORB_INPUT_REMAP)
    o = op.OracleExtractor(nf)
    for i in range(B):
        rect_img = oracle_remap(raws[i], mx, my)
        assert np.array_equal(ex.pyramid_level(0, i), rect_img), i
        mo, ko, do = o(rect_img, lap)
        assert n[i] == len(ko) and mono[i] == mo and kps[i, :n[i]].tobytes() == ko.tobytes() and np.array_equal(desc[i, :n[i]], do), i
    # strided raw frames (a view into a wider buffer) and the plain path on the same handle afterwards
    wide = np.zeros((2, rh, rw + 24), np.uint8)
    wide[:, :, 5:5 + rw] = raws[:2]
    n2, _, kps2, desc2 = ex.extract_batch(wide[:, :, 5:5 + rw], lap, flags=capi.ORB_INPUT_REMAP)
    assert np.array_equal(n2, n[:2]) and kps2[0, :n2[0]].tobytes() == kps[0, :n[0]].tobytes()
    rect0 = oracle_remap(raws[0], mx, my)
    n3, _, kps3, _ = ex.extract_batch(rect0[None], lap)
    assert n3[0] == n[0] and kps3[0, :n3[0]].tobytes() == kps[0, :n[0]].tobytes()


def test_remap_extreme_maps_and_errors():
    op.build()
    w, h = 320, 240
    rng = np.random.default_rng(7)
    raw = rng.integers(0, 256, (1, 200, 300), dtype=np.uint8)
    vals = np.array([-1e6, -1.0, -0.5, 0, 0.5, 298.5, 299, 299.5, 300, 1e6, 40000.3, 1e12, -1e12, np.nan, np.inf, -np.inf, 17.015625], np.float32)
    mx = rng.choice(vals, (h, w)); my = rng.choice(np.array([-1, -0.49, 0, 198.99, 199, 200, 70000.7, 3e9, np.nan, 33.5], np.float32), (h, w))
    ex = capi.ORBextractor(500, max_width=w, max_height=h)
    with pytest.raises(capi.OrbError):
        ex.extract_batch(raw, (0, 0), flags=capi.ORB_INPUT_REMAP)            # no maps yet
    ex.set_rectify_maps(mx, my)
    ex.extract_batch(raw, (0, 0), flags=capi.ORB_INPUT_REMAP)
    assert np.array_equal(ex.pyramid_level(0, 0), oracle_remap(raw[0], mx, my))
    with pytest.raises(capi.OrbError):
        ex.set_rectify_maps(np.zeros((h + 1, w), np.float32), np.zeros((h + 1, w), np.float32))   # larger than the handle
    ex.set_rectify_maps(None, None)
    with pytest.raises(capi.OrbError):
        ex.extract_batch(raw, (0, 0), flags=capi.ORB_INPUT_REMAP)


def oracle_resize(src, dw, dh):
    lib = op.oracle_lib()
    out = np.zeros((dh, dw), np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    lib.shim_resize(p(src), src.shape[1], src.shape[0], src.strides[0], p(out), dw, dh)
    return out


@pytest.mark.parametrize("raw,new", [((1504, 960), (752, 480)), ((1024, 768), (752, 480)), ((400, 300), (640, 480)), ((753, 481), (600, 350))])
def test_resized_input_equals_oracle(raw, new):
    """ORB_INPUT_RESIZE = System::TrackStereo's cv::resize(im, imToFeed, newImSize) (src/System.cc:262-264): exact 2x shrink (OpenCV's
    box filter), a general shrink, an enlargement, odd sizes. Checker: the shim's resize (pinned against cv2.resize)."""
    op.build()
    (rw, rh), (w, h) = raw, new
    B, nf, lap = 3, 1000, (0, 0)
    raws = np.stack([synth.mono_frame(6300 + i, rw, rh) for i in range(B)])
    ex = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B)
    with pytest.raises(capi.OrbError):
        ex.extract_batch(raws, lap, flags=capi.ORB_INPUT_RESIZE)            # no size yet
    ex.set_input_size(w, h)
    n, mono, kps, desc = ex.extract_batch(raws, lap, flags=capi.ORB_INPUT_RESIZE)
    o = op.OracleExtractor(nf)
    for i in range(B):
        small = oracle_resize(raws[i], w, h)
        assert np.array_equal(ex.pyramid_level(0, i), small), i
        mo, ko, do = o(small, lap)
        assert n[i] == len(ko) and mono[i] == mo and kps[i, :n[i]].tobytes() == ko.tobytes() and np.array_equal(desc[i, :n[i]], do), i
    # another raw size on the same handle rebuilds the tables
    raws2 = raws[:1, :rh - 7, :rw - 5].copy()
    n2, _, kps2, _ = ex.extract_batch(raws2, lap, flags=capi.ORB_INPUT_RESIZE)
    mo, ko, do = o(oracle_resize(raws2[0], w, h), lap)
    assert n2[0] == len(ko) and kps2[0, :n2[0]].tobytes() == ko.tobytes()
    with pytest.raises(capi.OrbError):
        ex.extract_batch(raws, lap, flags=capi.ORB_INPUT_RESIZE | capi.ORB_INPUT_REMAP)
