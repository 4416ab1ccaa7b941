"""Fisheye stereo triangulation: the CPU restatement (oracle/orb_oracle_kb8.cc) against the reference's own
lines compiled by line range (oracle/_ref/libmorb_ref_kb8.so: src/CameraModels/KannalaBrandt8.cpp:68-94,111-147,323-395,415-428 and
src/Frame.cc:1244-1273) on the mini Eigen stand-in, plus the checks that stand in for the missing Eigen: the Jacobi SVD against
numpy.linalg.svd (LAPACK) and the float pipeline against a float64 evaluation. CPU only."""
import os

import numpy as np
import pytest

from morb_slam_b200 import synth
from oracle import oracle_kb8_py as ok
from oracle.oracle_py import KP_DTYPE

have_ref = os.path.exists(ok.REF_KB8_SO)
needs_ref = pytest.mark.skipif(not have_ref, reason="oracle/_ref/libmorb_ref_kb8.so not built (no /root/reference)")
RIGS = ("tumvi", "parallel", "toed")


@needs_ref
@pytest.mark.parametrize("kind", RIGS)
def test_restatement_equals_reference_lines(kind):
    o, r = ok.oracle(), ok.reference()
    rig = synth.kb8_rig(kind)
    seen = set()
    for seed in range(4):
        xy1, xy2, s1, s2 = synth.kb8_pairs(100 + seed, rig, 3000)
        ro, po, _ = o.triangulate(rig, xy1, xy2, s1, s2)
        rr, pr, _ = r.triangulate(rig, xy1, xy2, s1, s2)
        assert ro.tobytes() == rr.tobytes() and po.tobytes() == pr.tobytes()
        seen |= set(np.where(ro > 0, 1, ro).astype(int).tolist())
        for cam, prec in ((rig["cam1"], rig["prec1"]), (rig["cam2"], rig["prec2"])):
            assert o.unproject(cam, prec, xy1).tobytes() == r.unproject(cam, prec, xy1).tobytes()
            pts = np.concatenate([po[ro > 0], np.float32([[0, 0, 1], [0, 0, -1], [1, 0, 0], [0, -2, 1e-9]])])
            assert o.project(cam, pts).tobytes() == r.project(cam, pts).tobytes()
    assert {1, -1, -2, -4} <= seen, seen     # every common exit is exercised (-3 / -5 depend on the rig)


def _frame_case(seed, rig, n=900, nr=800):
    """keypoints + kNN lists of one synthetic frame: true pairs, pairs that fail the ratio test, absent neighbours, and several
    queries that share one train keypoint (mvRightToLeftMatch keeps the last accepted one)"""
    rng = np.random.default_rng(seed)
    xy1, xy2, s1, s2 = synth.kb8_pairs(seed, rig, n)
    mono_l, mono_r = 37, 21
    kL = np.zeros(mono_l + n, KP_DTYPE); kR = np.zeros(mono_r + nr, KP_DTYPE)
    kL["x"][mono_l:], kL["y"][mono_l:] = xy1[:, 0], xy1[:, 1]
    kL["octave"] = rng.integers(0, 8, len(kL)); kR["octave"] = rng.integers(0, 8, len(kR))
    train = rng.permutation(n)[:nr]                      # right keypoint j observes pair train[j]
    kR["x"][mono_r:], kR["y"][mono_r:] = xy2[train, 0], xy2[train, 1]
    inv = np.full(n, -1); inv[train] = np.arange(nr)
    idx = np.full((n, 2), -1, np.int32); dist = np.full((n, 2), -1, np.int32)
    for i in range(n):
        j = inv[i] if inv[i] >= 0 and rng.random() < 0.8 else rng.integers(0, nr)
        idx[i] = (j, rng.integers(0, nr))
        dist[i] = (rng.integers(5, 60), rng.integers(40, 120))
    idx[::50, 1] = -1; dist[::50, 1] = -1                # fewer than two neighbours
    idx[7] = idx[5]; idx[9] = idx[5]; dist[5] = dist[7] = dist[9] = (10, 100)   # three queries, one train keypoint
    kL["x"][mono_l + 7], kL["y"][mono_l + 7] = kL["x"][mono_l + 5] + 0.25, kL["y"][mono_l + 5]
    kL["x"][mono_l + 9], kL["y"][mono_l + 9] = kL["x"][mono_l + 5], kL["y"][mono_l + 5] + 0.25
    sigma2 = (np.float32(1.2) ** np.arange(8, dtype=np.float32)) ** 2
    return kL, mono_l, kR, mono_r, sigma2.astype(np.float32), idx, dist


@needs_ref
@pytest.mark.parametrize("kind", RIGS)
def test_acceptance_loop_equals_reference_lines(kind):
    o, r = ok.oracle(), ok.reference()
    rig = synth.kb8_rig(kind)
    for seed in range(3):
        case = _frame_case(300 + seed, rig)
        a = o.fisheye_accept(rig, *case)
        b = r.fisheye_accept(rig, *case)
        for x, y in zip(a[:4], b[:4]):
            assert x.tobytes() == y.tobytes()
        l2r, r2l, depth, p3d, code, _ = a
        assert (code == 1).sum() > 50 and ((l2r >= 0) == (code == 1)).all() and ((depth > 0) == (code == 1)).all()
        # the inverse map holds the last accepted left keypoint of every matched right keypoint
        for j in np.unique(l2r[l2r >= 0]):
            assert r2l[j] == np.nonzero(l2r == j)[0].max()
    # empty inputs
    e = o.fisheye_accept(rig, case[0][:0], 0, case[2][:0], 0, case[4], case[5][:0], case[6][:0])
    assert len(e[0]) == 0 and len(e[1]) == 0


def test_jacobi_svd_against_lapack():
    """the stand-in for Eigen::JacobiSVD<Matrix4f>::matrixV(): singular values and the last column against numpy.linalg.svd"""
    o = ok.oracle()
    rng = np.random.default_rng(0)
    for i in range(1500):
        A = rng.standard_normal((4, 4)).astype(np.float32) * np.float32(10.0 ** rng.integers(-3, 4))
        if i % 3 == 0:
            A[3] = A[2] * np.float32(1.0001) + np.float32(1e-4) * rng.standard_normal(4).astype(np.float32)   # nearly rank 3
        V, sv = o.svd4_v(A)
        _, s, vt = np.linalg.svd(A.astype(np.float64))
        assert np.allclose(sv, s, rtol=1e-10, atol=1e-12 * s[0])
        assert np.allclose(V.T @ V, np.eye(4), atol=1e-12)
        if s[2] - s[3] > 1e-6 * s[0]:
            assert min(np.abs(V[:, 3] - vt[3]).max(), np.abs(V[:, 3] + vt[3]).max()) < 1e-9


def test_float_jacobi_svd_restatement_against_lapack():
    """orb_eigen_jacobi_svd4f (Eigen's two-sided float Jacobi SVD, restated): a valid SVD to float accuracy - singular values,
    orthonormal V, descending order, the null vector of nearly rank-3 matrices - against numpy.linalg.svd in float64."""
    o = ok.oracle()
    rng = np.random.default_rng(1)
    for i in range(1500):
        A = rng.standard_normal((4, 4)).astype(np.float32) * np.float32(10.0 ** rng.integers(-3, 4))
        if i % 3 == 0:
            A[3] = A[2] * np.float32(1.0001) + np.float32(1e-3) * rng.standard_normal(4).astype(np.float32)   # nearly rank 3
        if i % 50 == 7:
            A[:] = 0
        V, sv = o.jacobi_svd4f(A)
        _, s, vt = np.linalg.svd(A.astype(np.float64))
        assert np.all(np.diff(sv) <= 0)
        assert np.allclose(sv, s, rtol=2e-5, atol=2e-6 * max(s[0], 1e-30))
        assert np.allclose(V.astype(np.float64).T @ V.astype(np.float64), np.eye(4), atol=2e-6)
        if s[0] > 0 and s[2] - s[3] > 1e-2 * s[0]:
            assert min(np.abs(V[:, 3] - vt[3]).max(), np.abs(V[:, 3] + vt[3]).max()) < 2e-4


def _truth64(rig, xy1, xy2):
    """the same triangulation evaluated in float64 (Newton to convergence, LAPACK SVD)"""
    def unproj(cam, xy):
        cam = cam.astype(np.float64); pw = (xy.astype(np.float64) - cam[2:4]) / cam[0:2]
        td = np.minimum(np.hypot(pw[:, 0], pw[:, 1]), np.pi / 2); th = td.copy()
        for _ in range(30):
            t2 = th * th
            f = th * (1 + cam[4] * t2 + cam[5] * t2 ** 2 + cam[6] * t2 ** 3 + cam[7] * t2 ** 4) - td
            th = th - f / (1 + 3 * cam[4] * t2 + 5 * cam[5] * t2 ** 2 + 7 * cam[6] * t2 ** 3 + 9 * cam[7] * t2 ** 4)
        return pw * np.where(td > 1e-8, np.tan(th) / np.maximum(td, 1e-300), 1.0)[:, None]
    r1, r2 = unproj(rig["cam1"], xy1), unproj(rig["cam2"], xy2)
    R21 = rig["R12"].astype(np.float64).T
    T2 = np.hstack([R21, (-R21 @ rig["t12"].astype(np.float64))[:, None]])
    out = np.zeros((len(xy1), 3))
    for i in range(len(xy1)):
        A = np.array([[-1, 0, r1[i, 0], 0], [0, -1, r1[i, 1], 0], r2[i, 0] * T2[2] - T2[0], r2[i, 1] * T2[2] - T2[1]])
        v = np.linalg.svd(A)[2][3]
        out[i] = v[:3] / v[3]
    return out


@pytest.mark.parametrize("kind", RIGS)
def test_float_pipeline_against_float64(kind):
    """What the tolerance of this row rests on: with the parallax gate (cos <= 0.9998) the triangulation is well conditioned, the
    float pipeline stays within 1e-4 (relative) of the float64 evaluation on every accepted pair, so any correct float
    implementation - Eigen's included - agrees with the oracle to that order."""
    o = ok.oracle()
    rig = synth.kb8_rig(kind)
    xy1, xy2, s1, s2 = synth.kb8_pairs(5, rig, 3000)
    ret, p3d, _ = o.triangulate(rig, xy1, xy2, s1, s2)
    acc = ret > 0
    X = _truth64(rig, xy1[acc], xy2[acc])
    rel = np.abs(X - p3d[acc]).max(1) / np.abs(X).max(1)
    assert acc.sum() > 1000 and rel.max() < 1e-4, rel.max()


def test_known_answers():
    o = ok.oracle()
    rig = synth.kb8_rig("parallel")        # f = 190, no distortion, baseline 0.6 along x
    # a point on the optical axis of camera 1 at z = 2: pixel (cx, cy) in camera 1, x2 = -0.6 -> theta = atan(0.3) in camera 2
    X2 = np.float64([-0.6, 0.0, 2.0])
    th = np.arctan2(0.6, 2.0)
    xy2 = np.float32([[255.5 - 190.0 * th, 255.5]])
    ret, p3d, q = o.triangulate(rig, np.float32([[255.5, 255.5]]), xy2, [1.0], [1.0])
    assert abs(ret[0] - 2.0) < 1e-4 and np.allclose(p3d[0], [0, 0, 2], atol=1e-4)
    # the same pixel in both cameras: parallel rays -> parallax gate
    ret, _, _ = o.triangulate(rig, np.float32([[300, 200]]), np.float32([[300, 200]]), [1.0], [1.0])
    assert ret[0] == -1
    # the match is on the wrong side (negative disparity): the rays meet behind the cameras
    ret, _, _ = o.triangulate(rig, np.float32([[255.5, 255.5]]), np.float32([[255.5 + 30, 255.5]]), [1.0], [1.0])
    assert ret[0] == -2
    # 3 px of vertical offset at level 0: reprojection gate of the first camera (chi-square 5.991)
    ret, _, _ = o.triangulate(rig, np.float32([[255.5, 255.5]]), xy2 + np.float32([0, 5]), [1.0], [1.0])
    assert ret[0] == -4
    # ... accepted at a coarse level (sigma2 = 1.2^14)
    ret, _, _ = o.triangulate(rig, np.float32([[255.5, 255.5]]), xy2 + np.float32([0, 5]), [12.8], [12.8])
    assert ret[0] > 0
    # unproject / project round trip on the distorted cameras
    for kind in RIGS:
        r = synth.kb8_rig(kind)
        xy = np.random.default_rng(1).uniform(110, 400, (500, 2)).astype(np.float32)   # theta < 1.1 rad: tan(theta) keeps its sign
        back = o.project(r["cam1"], o.unproject(r["cam1"], r["prec1"], xy))
        assert np.abs(back - xy).max() < 2e-3


def test_camera_model_against_opencv_fisheye():
    """KannalaBrandt8::project / unproject are OpenCV's fisheye model (theta_d = theta (1 + k1 theta^2 + ... + k4 theta^8)):
    an independent implementation of the same published model (cv2.fisheye, double precision) agrees with the float restatement."""
    cv2 = pytest.importorskip("cv2")
    o = ok.oracle()
    rng = np.random.default_rng(5)
    for kind in RIGS:
        cam = synth.kb8_rig(kind)["cam1"].astype(np.float64)
        K = np.array([[cam[0], 0, cam[2]], [0, cam[1], cam[3]], [0, 0, 1]])
        D = cam[4:8].reshape(4, 1)
        z = rng.uniform(0.5, 20.0, 400)
        ang = rng.uniform(0, 2 * np.pi, 400); rad = np.tan(rng.uniform(0.0, 1.1, 400))
        P = np.stack([z * rad * np.cos(ang), z * rad * np.sin(ang), z], 1)
        uv_cv, _ = cv2.fisheye.projectPoints(P.reshape(-1, 1, 3), np.zeros(3), np.zeros(3), K, D)
        uv = o.project(cam.astype(np.float32), P.astype(np.float32))
        assert np.abs(uv - uv_cv.reshape(-1, 2)).max() < 2e-3            # float32 evaluation of a ~500 px coordinate
        rays_cv = cv2.fisheye.undistortPoints(uv_cv.astype(np.float64), K, D).reshape(-1, 2)
        rays = o.unproject(cam.astype(np.float32), 1e-6, uv_cv.reshape(-1, 2).astype(np.float32))
        assert np.abs(rays[:, :2] - rays_cv).max() < 2e-4 * max(1.0, np.abs(rays_cv).max())
        assert np.all(rays[:, 2] == 1.0)


@pytest.mark.parametrize("kind", RIGS)
def test_triangulation_against_opencv_dlt(kind):
    """KannalaBrandt8::Triangulate is the linear DLT triangulation: cv2.triangulatePoints builds the same 4 x 4 system
    (x P[2] - P[0], y P[2] - P[1] for both views) and takes the last right singular vector with OpenCV's own SVD. On the rays the
    oracle unprojects, the two agree to float rounding for every accepted pair - an independent stand-in for the Eigen SVD."""
    cv2 = pytest.importorskip("cv2")
    o = ok.oracle()
    rig = synth.kb8_rig(kind)
    xy1, xy2, s1, s2 = synth.kb8_pairs(11, rig, 3000)
    ret, p3d, _ = o.triangulate(rig, xy1, xy2, s1, s2)
    acc = ret > 0
    r1 = o.unproject(rig["cam1"], rig["prec1"], xy1[acc])[:, :2].astype(np.float64)
    r2 = o.unproject(rig["cam2"], rig["prec2"], xy2[acc])[:, :2].astype(np.float64)
    R21 = rig["R12"].astype(np.float64).T
    P1 = np.hstack([np.eye(3), np.zeros((3, 1))])
    P2 = np.hstack([R21, (-R21 @ rig["t12"].astype(np.float64))[:, None]])
    Xh = cv2.triangulatePoints(P1, P2, r1.T.copy(), r2.T.copy())
    X = (Xh[:3] / Xh[3]).T
    rel = np.abs(X - p3d[acc]).max(1) / np.abs(X).max(1)
    assert acc.sum() > 1000 and rel.max() < 1e-4, rel.max()
