"""Windowed matcher on the GPU (include/orb_b200.h: orb_assign_features_to_grid, orb_search_by_projection) against the
CPU oracle restatement (oracle/orb_oracle_match.cc, pinned against the reference's own code by
tests/test_oracle_match.py). Bit-exact: grid lists, match indices, match counts."""
import numpy as np
import pytest

from morb_slam_b200 import capi, synth
from oracle import oracle_py as op
from oracle import oracle_match_py as om

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pair():
    op.build()
    w, h, nf, lap, fx, b = synth.CONFIGS["euroc"]
    B = 6
    Ls = np.stack([synth.stereo_pair(4200 + i, w, h)[0] for i in range(B)])
    Rs = np.stack([synth.stereo_pair(4200 + i, w, h)[1] for i in range(B)])
    exL = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B)
    exR = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B)
    nL, _, kL, dL = exL.extract_batch(Ls, lap)
    nR, _, kR, dR = exR.extract_batch(Rs, lap)
    mbf, mb = float(np.float32(fx * b)), float(np.float32(b))
    uR = np.full((B, exL.kcap), -1, np.float32)
    dp = np.full((B, exL.kcap), -1, np.float32)
    capi.compute_stereo_matches_batch(exL, exR, mbf, float(np.float32(fx)), out=(uR, dp))
    scale = exL.tables()["scale"]
    return dict(w=w, h=h, B=B, exL=exL, exR=exR, nL=nL, kL=kL, dL=dL, nR=nR, kR=kR, dR=dR, uR=uR, scale=scale, mbf=mbf, mb=mb)


def test_grid_equals_oracle(pair):
    p = pair
    gp = capi.grid_params(p["w"], p["h"])
    capi.assign_features_to_grid(p["exL"], gp)
    o = om.oracle()
    for f in range(p["B"]):
        off, idx = capi.get_grid(p["exL"], f)
        oo, oi = o.assign_grid(p["kL"][f, :p["nL"][f]], gp)
        assert np.array_equal(off, oo) and np.array_equal(idx, oi), f


CASES = [
    (7.0, False, 0.0, True, 4.0, 0.8),
    (15.0, True, 0.0, True, 8.0, 0.8),
    (7.0, False, 0.5, True, 4.0, 0.8),
    (7.0, False, -0.5, True, 4.0, 0.8),
    (15.0, False, 0.0, False, 10.0, 0.3),
    (30.0, False, 0.0, True, 2.0, 1.0),
]


@pytest.mark.parametrize("th,mono,tlc,ori,jit,pobs", CASES)
def test_search_by_projection_equals_oracle(pair, th, mono, tlc, ori, jit, pobs):
    """Current frames = the left images' device-resident keypoints / descriptors / uRight; last frames = the right
    images' keypoints turned into projected map points (jittered, with invalid / behind-camera / outside / unlocked /
    duplicate queries)."""
    p = pair
    B, kcap = p["B"], p["exL"].kcap
    gp = capi.grid_params(p["w"], p["h"])
    capi.assign_features_to_grid(p["exL"], gp)
    qcap = p["exR"].kcap
    Q = np.zeros((B, qcap), capi.Q_DTYPE)
    QD = np.zeros((B, qcap, 32), np.uint8)
    for f in range(B):
        n = p["nR"][f]
        q, qd = om.synth_queries(100 * f + 7, p["kR"][f, :n], p["dR"][f, :n], None, None, p["w"], p["h"], p_obs=pobs, jitter=jit)
        Q[f, :n] = q
        QD[f, :n] = qd
    tz = np.full(B, tlc, np.float32)
    nm, match = capi.search_by_projection(p["exL"], Q, QD, p["nR"], th, mono, tz, p["mb"], p["mbf"], ori)
    o = om.oracle()
    for f in range(B):
        nC, n = p["nL"][f], p["nR"][f]
        no, mo = o.search_by_projection(p["kL"][f, :nC], p["dL"][f, :nC], p["uR"][f, :nC], p["scale"], gp, p["mb"], p["mbf"],
                                        Q[f, :n], QD[f, :n], th, mono, tlc, ori)
        assert nm[f] == no, (f, nm[f], no)
        assert np.array_equal(match[f, :nC], mo), f
        assert np.all(match[f, nC:] == -1)
        assert no > 50


def test_search_by_projection_lock_pressure(pair):
    """All queries of a frame aim at a handful of keypoints with identical descriptors: every stored candidate of the
    later queries is locked, which drives the resolver's exact re-scan path."""
    p = pair
    B = p["B"]
    gp = capi.grid_params(p["w"], p["h"])
    capi.assign_features_to_grid(p["exL"], gp)
    qcap = p["exR"].kcap
    Q = np.zeros((B, qcap), capi.Q_DTYPE)
    QD = np.zeros((B, qcap, 32), np.uint8)
    nq = np.zeros(B, np.int32)
    for f in range(B):
        nC = p["nL"][f]
        k = p["kL"][f, :nC]
        # 400 queries, all at the densest spot of the frame, descriptor = that of the keypoint next to it
        c = int(np.argmin(np.abs(k["x"] - np.median(k["x"])) + np.abs(k["y"] - np.median(k["y"]))))
        m = 400
        Q[f, :m]["u"] = k["x"][c]; Q[f, :m]["v"] = k["y"][c]; Q[f, :m]["z"] = 2.0
        Q[f, :m]["angle"] = k["angle"][c]; Q[f, :m]["octave"] = 3; Q[f, :m]["flags"] = 3
        QD[f, :m] = p["dL"][f, c]
        nq[f] = m
    tz = np.zeros(B, np.float32)
    o = om.oracle()
    deepest = 0
    for th in (40.0, 120.0, 400.0):
        nm, match = capi.search_by_projection(p["exL"], Q, QD, nq, th, False, tz, p["mb"], p["mbf"], False)
        deepest = max(deepest, int(nm.max()))
        for f in range(B):
            nC = p["nL"][f]
            no, mo = o.search_by_projection(p["kL"][f, :nC], p["dL"][f, :nC], p["uR"][f, :nC], p["scale"], gp, p["mb"], p["mbf"],
                                            Q[f, :nq[f]], QD[f, :nq[f]], th, False, 0.0, False)
            assert nm[f] == no and np.array_equal(match[f, :nC], mo), (th, f)
    # identical locked queries take one keypoint each: more matches than stored candidates (SP_K = 4 in orb_match.cu)
    # means the later ones were found by the re-scan
    assert deepest > 4, deepest


LOCAL_CASES = [
    # th, nnratio, jitter, p_obs, share of keypoints locked before the call, n_extra
    (1.0, 0.8, 2.0, 0.9, 0.3, 0.5),
    (3.0, 0.8, 3.0, 0.9, 0.3, 0.5),
    (5.0, 0.8, 6.0, 0.9, 0.0, 1.0),
    (15.0, 0.9, 10.0, 0.3, 0.5, 0.5),
    (3.0, 0.6, 1.0, 1.0, 0.0, 0.5),
    (40.0, 0.8, 10.0, 1.0, 0.0, 2.0),    # very wide windows: many candidates per map point, deep lock chains
]


@pytest.mark.parametrize("th,ratio,jit,pobs,plock,extra", LOCAL_CASES)
def test_search_local_points_equals_oracle(pair, th, ratio, jit, pobs, plock, extra):
    """orb_search_local_points (ORBmatcher::SearchByProjection(F, vpMapPoints, th), src/ORBmatcher.cc:42-209): F = the
    left images' device-resident keypoints / descriptors / uRight; local map = the frame's own keypoints turned into
    tracked map points (jitter, bit flips, wrong associations, out-of-view / unobserved / duplicate points)."""
    p = pair
    B, kcap = p["B"], p["exL"].kcap
    gp = capi.grid_params(p["w"], p["h"])
    capi.assign_features_to_grid(p["exL"], gp)
    qs = []
    for f in range(B):
        nC = p["nL"][f]
        qs.append(synth.synth_track_queries(300 * f + 11, p["kL"][f, :nC], p["dL"][f, :nC], p["uR"][f, :nC], p["w"], p["h"], n_extra=extra,
                                            p_obs=pobs, jitter=jit, mbf=p["mbf"]))
    qcap = max(len(q) for q, _ in qs) + 3
    Q = np.zeros((B, qcap), capi.TQ_DTYPE)
    QD = np.zeros((B, qcap, 32), np.uint8)
    nq = np.zeros(B, np.int32)
    for f, (q, qd) in enumerate(qs):
        Q[f, :len(q)] = q; QD[f, :len(q)] = qd; nq[f] = len(q)
    rng = np.random.default_rng(int(th * 10))
    locked0 = (rng.random((B, kcap)) < plock).astype(np.uint8)
    nm, match = capi.search_local_points(p["exL"], Q, QD, nq, locked0 if plock > 0 else None, th, ratio)
    o = om.oracle()
    for f in range(B):
        nC, n = p["nL"][f], nq[f]
        no, mo = o.search_local_points(p["kL"][f, :nC], p["dL"][f, :nC], p["uR"][f, :nC], locked0[f, :nC] if plock > 0 else np.zeros(nC, np.uint8),
                                       p["scale"], gp, Q[f, :n], QD[f, :n], th, ratio)
        assert nm[f] == no, (f, nm[f], no)
        assert np.array_equal(match[f, :nC], mo), f
        assert np.all(match[f, nC:] == -1)
        assert no > 100


def test_search_local_points_lock_pressure(pair):
    """Many identical map points aimed at one spot: each takes the best keypoint still free, so the stored candidates of
    the later ones are all locked and the resolver's exact re-scan decides (more matches than SL_K = 4 stored ones)."""
    p = pair
    B, kcap = p["B"], p["exL"].kcap
    gp = capi.grid_params(p["w"], p["h"])
    capi.assign_features_to_grid(p["exL"], gp)
    m = 300
    Q = np.zeros((B, m), capi.TQ_DTYPE)
    QD = np.zeros((B, m, 32), np.uint8)
    nq = np.full(B, m, np.int32)
    for f in range(B):
        nC = p["nL"][f]
        k = p["kL"][f, :nC]
        c = int(np.argmin(np.abs(k["x"] - np.median(k["x"])) + np.abs(k["y"] - np.median(k["y"]))))
        Q[f]["proj_x"] = k["x"][c]; Q[f]["proj_y"] = k["y"][c]; Q[f]["proj_xr"] = k["x"][c] - 5
        Q[f]["view_cos"] = 0.5; Q[f]["level"] = 1; Q[f]["flags"] = 3
        QD[f] = p["dL"][f, c]
    o = om.oracle()
    deepest = 0
    for th, ratio in ((30.0, 1.0), (100.0, 1.0), (100.0, 0.9)):
        nm, match = capi.search_local_points(p["exL"], Q, QD, nq, None, th, ratio)
        deepest = max(deepest, int(nm.max()))
        for f in range(B):
            nC = p["nL"][f]
            no, mo = o.search_local_points(p["kL"][f, :nC], p["dL"][f, :nC], p["uR"][f, :nC], np.zeros(nC, np.uint8), p["scale"], gp,
                                           Q[f], QD[f], th, ratio)
            assert nm[f] == no and np.array_equal(match[f, :nC], mo), (th, ratio, f)
    assert deepest > 4, deepest


def _shim_undistort(pts, K, D, P):
    import ctypes as C
    lib = op.oracle_lib()
    lib.shim_undistort_points.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    pts = np.ascontiguousarray(pts, np.float32)
    out = np.zeros_like(pts)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    lib.shim_undistort_points(p(pts), len(pts), p(K), p(D), len(D), p(P), p(out))
    return out


@pytest.mark.parametrize("dist", [(-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05), (0.262383, -0.953104, -0.005358, 0.002628, 1.163314),
                                  (0.0, 0.1, 0.0, 0.0)])
def test_undistort_keypoints_feed_grid_and_search(pair, dist):
    """orb_undistort_keypoints (Frame::UndistortKeyPoints, src/Frame.cc:829-857): mvKeysUn bit-exact against the restatement of
    cv::undistortPoints (pinned against cv2), and the grid / search then work on mvKeysUn like the reference (EuRoC monocular)."""
    p = pair
    B, kcap = p["B"], p["exL"].kcap
    K = np.array([[458.654, 0, 367.215], [0, 457.296, 248.375], [0, 0, 1]], np.float32)
    D = np.array(dist, np.float32)
    un = capi.undistort_keypoints(p["exL"], K, D)
    gp = capi.grid_params(p["w"], p["h"])
    capi.assign_features_to_grid(p["exL"], gp)
    o = om.oracle()
    qcap = p["exR"].kcap
    Q = np.zeros((B, qcap), capi.Q_DTYPE)
    QD = np.zeros((B, qcap, 32), np.uint8)
    want_un = []
    for f in range(B):
        nC = p["nL"][f]
        k = p["kL"][f, :nC].copy()
        if dist[0] != 0.0:
            xy = _shim_undistort(np.stack([k["x"], k["y"]], 1), K, D, K)
            k["x"], k["y"] = xy[:, 0], xy[:, 1]
        assert un[f, :nC].tobytes() == k.tobytes(), f
        assert (dist[0] == 0.0) == (k.tobytes() == p["kL"][f, :nC].tobytes())
        want_un.append(k)
        off, idx = capi.get_grid(p["exL"], f)
        oo, oi = o.assign_grid(k, gp)
        assert np.array_equal(off, oo) and np.array_equal(idx, oi), f
        n = p["nR"][f]
        q, qd = om.synth_queries(100 * f + 9, p["kR"][f, :n], p["dR"][f, :n], None, None, p["w"], p["h"])
        Q[f, :n], QD[f, :n] = q, qd
    tz = np.zeros(B, np.float32)
    nm, match = capi.search_by_projection(p["exL"], Q, QD, p["nR"], 7.0, True, tz, p["mb"], p["mbf"], True)
    for f in range(B):
        nC, n = p["nL"][f], p["nR"][f]
        no, mo = o.search_by_projection(want_un[f], p["dL"][f, :nC], p["uR"][f, :nC], p["scale"], gp, p["mb"], p["mbf"], Q[f, :n], QD[f, :n],
                                        7.0, True, 0.0, True)
        assert nm[f] == no and np.array_equal(match[f, :nC], mo), f
    # the next extraction forgets mvKeysUn
    capi.undistort_keypoints(p["exL"], K, np.zeros(4, np.float32))
    capi.assign_features_to_grid(p["exL"], gp)
    off, idx = capi.get_grid(p["exL"], 0)
    oo, oi = o.assign_grid(p["kL"][0, :p["nL"][0]], gp)
    assert np.array_equal(off, oo) and np.array_equal(idx, oi)
