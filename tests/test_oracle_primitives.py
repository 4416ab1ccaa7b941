"""Pins the oracle's OpenCV-primitive restatements (oracle/shim) bit-exact against the cv2 wheel and
its sinf/cosf restatement against this image's libm. CPU only."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle_py as op

cv2 = pytest.importorskip("cv2")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _rand_img(rng, w, h, kind):
    if kind == "noise":
        return rng.integers(0, 256, (h, w), dtype=np.uint8)
    if kind == "smooth":
        a = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2)).astype(np.float32)
        return np.clip(cv2.resize(a, (w, h), interpolation=cv2.INTER_CUBIC), 0, 255).astype(np.uint8)
    if kind == "quant":
        a = rng.integers(0, 256, (h // 4 + 2, w // 4 + 2)).astype(np.float32)
        b = np.clip(cv2.resize(a, (w, h), interpolation=cv2.INTER_LINEAR), 0, 255).astype(np.uint8)
        return ((b // 32) * 32).astype(np.uint8)
    raise ValueError(kind)


@pytest.mark.parametrize("size", [(752, 480), (512, 512), (1241, 376), (97, 71), (640, 480), (333, 257)])
def test_resize_chain_matches_cv2(size):
    lib = op.oracle_lib()
    rng = np.random.default_rng(size[0])
    w, h = size
    src = _rand_img(rng, w, h, "noise")
    inv = 1.0
    for l in range(1, 8):
        inv = np.float32(1.0) / (np.float32(1.2) ** l)
        dw, dh = int(np.rint(np.float32(w) * np.float32(inv))), int(np.rint(np.float32(h) * np.float32(inv)))
        if dw < 8 or dh < 8:
            break
        ref = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
        out = np.zeros((dh, dw), np.uint8)
        lib.shim_resize(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(out), dw, dh)
        assert np.array_equal(out, ref), (size, l)
        src = ref


@pytest.mark.parametrize("pair", [((100, 80), (73, 91)), ((64, 64), (32, 32)), ((50, 40), (120, 90)), ((301, 203), (300, 202))])
def test_resize_arbitrary_ratio_matches_cv2(pair):
    lib = op.oracle_lib()
    (sw, sh), (dw, dh) = pair
    src = _rand_img(np.random.default_rng(7), sw, sh, "noise")
    ref = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
    out = np.zeros((dh, dw), np.uint8)
    lib.shim_resize(_p(src), sw, sh, src.strides[0], _p(out), dw, dh)
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("size", [(752, 480), (627, 400), (210, 134), (143, 143), (346, 105), (9, 8)])
@pytest.mark.parametrize("kind", ["noise", "smooth"])
def test_gauss7_matches_cv2(size, kind):
    lib = op.oracle_lib()
    w, h = size
    src = _rand_img(np.random.default_rng(w + h), w, h, kind)
    ref = cv2.GaussianBlur(src, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
    out = np.zeros_like(src)
    lib.shim_gauss7(_p(src), w, h, src.strides[0], _p(out))
    assert np.array_equal(out, ref)


def test_gauss7_impulse_known_answer():
    lib = op.oracle_lib()
    src = np.zeros((21, 21), np.uint8)
    src[10, 10] = 255
    out = np.zeros_like(src)
    lib.shim_gauss7(_p(src), 21, 21, 21, _p(out))
    k = np.array([18, 34, 48, 56, 48, 34, 18])
    expect = (np.outer(k, k) * 255 + 32768) >> 16
    assert np.array_equal(out[7:14, 7:14], expect)


@pytest.mark.parametrize("kind", ["noise", "smooth", "quant"])
@pytest.mark.parametrize("th", [20, 7])
def test_fast_matches_cv2(kind, th):
    lib = op.oracle_lib()
    det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    for seed in range(10):
        rng = np.random.default_rng(seed * 31 + th)
        w, h = int(rng.integers(7, 90)), int(rng.integers(7, 90))
        img = _rand_img(rng, w, h, kind)
        kps = det.detect(img, None)
        ref = np.array([[int(k.pt[0]), int(k.pt[1]), int(k.response)] for k in kps], np.int32).reshape(-1, 3)
        out = np.zeros((w * h, 3), np.int32)
        n = lib.shim_fast(_p(img), w, h, img.strides[0], th, _p(out), len(out))
        assert n == len(ref), (kind, th, seed, n, len(ref))
        assert np.array_equal(out[:n], ref)


def test_fast_on_roi_with_stride():
    lib = op.oracle_lib()
    det = cv2.FastFeatureDetector_create(threshold=20, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    big = _rand_img(np.random.default_rng(3), 200, 150, "smooth")
    roi = big[30:90, 50:120]
    kps = det.detect(np.ascontiguousarray(roi), None)
    ref = np.array([[int(k.pt[0]), int(k.pt[1]), int(k.response)] for k in kps], np.int32).reshape(-1, 3)
    out = np.zeros((4096, 3), np.int32)
    n = lib.shim_fast(C.c_void_p(big.ctypes.data + 30 * big.strides[0] + 50), 70, 60, big.strides[0], 20, _p(out), 4096)
    assert n == len(ref) and np.array_equal(out[:n], ref)


def test_fastatan2_matches_cv2():
    lib = op.oracle_lib()
    rng = np.random.default_rng(0)
    ys = rng.integers(-2_900_000, 2_900_000, 20000).astype(np.float32)
    xs = rng.integers(-2_900_000, 2_900_000, 20000).astype(np.float32)
    for y, x in list(zip(ys, xs)) + [(0, 0), (0, -5), (-3, 0), (1, 1), (5, 0), (0, 7), (-1, -1)]:
        a = lib.shim_fastatan2(float(y), float(x))
        b = cv2.fastAtan2(float(y), float(x))
        assert np.float32(a).tobytes() == np.float32(b).tobytes(), (y, x, a, b)
    assert abs(lib.shim_fastatan2(1.0, 1.0) - 44.990456) < 1e-5
    assert lib.shim_fastatan2(0.0, 0.0) == 0.0
    assert lib.shim_fastatan2(0.0, -5.0) == 180.0
    assert lib.shim_fastatan2(-3.0, 0.0) == 270.0


def test_border101_matches_cv2():
    lib = op.oracle_lib()
    src = _rand_img(np.random.default_rng(5), 40, 30, "noise")
    ref = cv2.copyMakeBorder(src, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
    out = np.zeros_like(ref)
    lib.shim_border101(_p(src), 40, 30, _p(out), 19)
    assert np.array_equal(out, ref)


def test_sincosf_restatement_matches_libm():
    lib = op.oracle_lib()
    rng = np.random.default_rng(1)
    # every angle the reference can produce is fastAtan2-degrees * (float)(pi/180): sample the range densely
    deg = np.concatenate([rng.uniform(0, 360, 200000), np.arange(0, 360, 0.25)]).astype(np.float32)
    rad = deg * np.float32(np.pi / np.float32(180.0))
    for x in rad[:60000]:
        assert lib.restated_sinf(float(x)) == lib.libm_sinf(float(x))
        assert lib.restated_cosf(float(x)) == lib.libm_cosf(float(x))


def test_sincosf_restatement_negative_angles():
    """KannalaBrandt8::project calls cosf / sinf on psi = atan2f(y, x) in [-pi, pi] (reference src/CameraModels/KannalaBrandt8.cpp:92-93)"""
    lib = op.oracle_lib()
    rng = np.random.default_rng(2)
    for x in rng.uniform(-np.pi, np.pi, 40000).astype(np.float32):
        assert lib.restated_sinf(float(x)) == lib.libm_sinf(float(x))
        assert lib.restated_cosf(float(x)) == lib.libm_cosf(float(x))


def test_tanf_atanf_atan2f_restatements_match_libm():
    """oracle/libm_restate.h (host twin of morb_slam_b200/csrc/orb_libm_glibc.cuh) against this image's glibc: every 61st float of
    tanf's range [0, 3 pi / 4), every 127th positive float for atanf, 300 k random and structured pairs for atan2f. The exhaustive
    form (every float; 0 differences) is tools/probe/libm_check.cc."""
    import ctypes as C
    lib = op.oracle_lib()
    for f in (lib.restated_tanf_mismatches, lib.restated_atanf_mismatches):
        f.restype = C.c_long
        f.argtypes = [C.c_uint, C.c_uint, C.c_uint]
    assert lib.restated_tanf_mismatches(0, 0x4016cbe4, 61) == 0
    assert lib.restated_tanf_mismatches(0x3fc00000, 0x3fd00000, 1) == 0      # every float around pi / 2
    assert lib.restated_atanf_mismatches(0, 0x7f800000, 127) == 0
    for f in (lib.restated_atan2f, lib.libm_atan2f):
        f.restype = C.c_float
        f.argtypes = [C.c_float, C.c_float]
    rng = np.random.default_rng(3)
    ys = np.concatenate([rng.standard_normal(100000) * 10 ** rng.uniform(-6, 6, 100000), rng.uniform(-600, 600, 50000), [0.0, -0.0, 1.0, 0.0, 5.0]])
    xs = np.concatenate([rng.standard_normal(100000) * 10 ** rng.uniform(-6, 6, 100000), rng.uniform(-600, 600, 50000), [1.0, -1.0, 0.0, -3.0, 1.0]])
    for y, x in zip(ys.astype(np.float32)[:60000], xs.astype(np.float32)[:60000]):
        a, b = lib.restated_atan2f(float(y), float(x)), lib.libm_atan2f(float(y), float(x))
        assert np.float32(a).tobytes() == np.float32(b).tobytes(), (y, x)


def test_knn2_matches_bfmatcher_including_ties():
    rng = np.random.default_rng(2)
    q = rng.integers(0, 256, (64, 32), dtype=np.uint8)
    db = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    db[17] = db[3]
    db[200] = q[5]
    db[201] = q[5]          # exact duplicates: lower train index must win
    q[9] = db[40]
    db[41] = db[40]
    idx, dist = op.oracle_knn2(q, db)
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, db, k=2)
    for i, mm in enumerate(m):
        assert [x.trainIdx for x in mm] == list(idx[i]), i
        assert [int(x.distance) for x in mm] == list(dist[i]), i
    # SURVEY.md A.6 known answer
    base = np.zeros((1, 32), np.uint8)
    train = np.zeros((6, 32), np.uint8)
    for j, d in enumerate([3, 1, 1, 2, 1, 1]):
        train[j, 0] = (1 << d) - 1
    idx, dist = op.oracle_knn2(base, train)
    assert list(idx[0]) == [1, 2] and list(dist[0]) == [1, 1]
    idx, dist = op.oracle_knn2(base, train[::-1].copy())
    assert list(idx[0]) == [0, 1]
    # fewer than two train rows
    idx, dist = op.oracle_knn2(base, train[:1])
    assert list(idx[0]) == [0, -1] and list(dist[0]) == [3, -1]


def test_ratio_test_is_evaluated_in_double():
    d = np.array([[7, 10], [6, 10], [14, 20], [21, 30], [0, 0], [5, -1]], np.int32)
    got = op.oracle_ratio_test(d)
    expect = [float(np.float32(a)) < float(np.float32(b)) * 0.7 if b >= 0 else False for a, b in d]
    assert list(got) == expect


def test_sad_norm_l1_matches_cv2():
    rng = np.random.default_rng(4)
    a = rng.integers(0, 256, (11, 11), dtype=np.uint8)
    b = rng.integers(0, 256, (11, 11), dtype=np.uint8)
    assert cv2.norm(a, b, cv2.NORM_L1) == float(np.abs(a.astype(int) - b.astype(int)).sum())


def _remap(lib, src, mx, my):
    import ctypes as C
    lib.shim_remap.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    dh, dw = mx.shape
    out = np.zeros((dh, dw), np.uint8)
    lib.shim_remap(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(mx), _p(my), dw, dh, _p(out))
    return out


@pytest.mark.parametrize("trial", range(16))
def test_remap_linear_matches_cv2(trial):
    """cv::remap(src, dst, M1 CV_32FC1, M2 CV_32FC1, INTER_LINEAR) (System::TrackStereo, src/System.cc:260-261): random
    maps far outside the source, smooth warps, exact integers / 1/64 ties of cvRound, borders, values outside the int
    range, NaN and infinities."""
    lib = op.oracle_lib()
    rng = np.random.default_rng(900 + trial)
    sw, sh = int(rng.integers(8, 300)), int(rng.integers(8, 200))
    dw, dh = int(rng.integers(1, 300)), int(rng.integers(1, 200))
    src = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
    kind = trial % 4
    if kind == 0:
        mx = rng.uniform(-20, sw + 20, (dh, dw)).astype(np.float32)
        my = rng.uniform(-20, sh + 20, (dh, dw)).astype(np.float32)
    elif kind == 1:
        yy, xx = np.mgrid[0:dh, 0:dw].astype(np.float32)
        mx = (xx * sw / dw + 3 * np.sin(yy / 7)).astype(np.float32)
        my = (yy * sh / dh + 2 * np.cos(xx / 5)).astype(np.float32)
    elif kind == 2:
        mx = (rng.integers(-2 * 64, (sw + 2) * 64, (dh, dw)) / 64).astype(np.float32)
        my = (rng.integers(-2 * 64, (sh + 2) * 64, (dh, dw)) / 64).astype(np.float32)
    else:
        mx = rng.choice(np.array([-1e6, -1.0, -0.5, 0, 0.5, sw - 1.5, sw - 1, sw - 0.5, sw, 1e6, 40000.3, 1e12, -1e12, np.nan, np.inf, -np.inf],
                                 np.float32), (dh, dw))
        my = rng.choice(np.array([-1e6, -1, -0.49, 0, sh - 1.01, sh - 1, sh, 70000.7, 3e9, np.nan], np.float32), (dh, dw))
    assert np.array_equal(_remap(lib, src, mx, my), cv2.remap(src, mx, my, cv2.INTER_LINEAR))


def test_remap_euroc_like_rectification_matches_cv2():
    """Maps as cv::initUndistortRectifyMap produces them for a radial-tangential camera (src/Settings.cc:540-545), CV_32FC1"""
    lib = op.oracle_lib()
    w, h = 752, 480
    K = np.array([[458.654, 0, 367.215], [0, 457.296, 248.375], [0, 0, 1]])
    D = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05])
    R = cv2.Rodrigues(np.array([0.01, -0.02, 0.005]))[0]
    P = np.array([[435.2, 0, 367.4, 0], [0, 435.2, 252.2, 0], [0, 0, 1, 0]])
    mx, my = cv2.initUndistortRectifyMap(K, D, R, P, (w, h), cv2.CV_32FC1)
    src = _rand_img(np.random.default_rng(3), w, h, "noise")
    assert np.array_equal(_remap(lib, src, mx, my), cv2.remap(src, mx, my, cv2.INTER_LINEAR))


def _undistort(lib, pts, K, D, P):
    import ctypes as C
    lib.shim_undistort_points.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    pts = np.ascontiguousarray(pts, np.float32); K = np.ascontiguousarray(K, np.float32); P = np.ascontiguousarray(P, np.float32)
    D = np.ascontiguousarray(D, np.float32).ravel()
    out = np.zeros_like(pts)
    lib.shim_undistort_points(_p(pts), len(pts), _p(K), _p(D), len(D), _p(P), _p(out))
    return out


UNDISTORT_CAMERAS = [
    # fx, fy, cx, cy, distortion (float, as Settings stores it), image size
    ((458.654, 457.296, 367.215, 248.375), (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05), (752, 480)),       # EuRoC cam0
    ((517.306408, 516.469215, 318.643040, 255.313989), (0.262383, -0.953104, -0.005358, 0.002628, 1.163314), (640, 480)),  # TUM1, with k3
    ((300.0, 300.0, 320.0, 240.0), (-0.9, 0.5, 0.01, -0.02), (640, 480)),                                             # strong: icdist < 0 far out
    ((190.97847, 190.97330, 254.93170, 256.89744), (0.0034823, 0.0007150, -0.0020532, 0.00020293, 0.0, 0.0, 0.0, 0.0), (512, 512)),  # 8 coefficients
]


@pytest.mark.parametrize("cam", UNDISTORT_CAMERAS)
def test_undistort_points_matches_cv2(cam):
    """cv::undistortPoints(mat, mat, K, mDistCoef, cv::Mat(), mK) (Frame::UndistortKeyPoints, src/Frame.cc:845-846): float
    results bit for bit, incl. points far outside the image and the icdist < 0 bail-out."""
    lib = op.oracle_lib()
    (fx, fy, cx, cy), dist, (w, h) = cam
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32)
    D = np.array(dist, np.float32).reshape(-1, 1)
    rng = np.random.default_rng(int(fx))
    pts = np.concatenate([np.stack([rng.uniform(0, w, 3000), rng.uniform(0, h, 3000)], 1),
                          np.stack([rng.uniform(-3 * w, 4 * w, 500), rng.uniform(-3 * h, 4 * h, 500)], 1),
                          np.array([[0, 0], [w, 0], [0, h], [w, h], [cx, cy]])]).astype(np.float32)   # ComputeImageBounds corners
    for P in (K, np.array([[fx * 0.9, 0, cx + 3], [0, fy * 0.95, cy - 2], [0, 0, 1]], np.float32)):
        ref = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, D, None, P).reshape(-1, 2)
        out = _undistort(lib, pts, K, D, P)
        assert out.tobytes() == ref.tobytes()
