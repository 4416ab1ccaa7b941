"""LocalMapping / Relocalization matchers on the GPU (include/orb_b200.h: orb_load_frames, orb_fuse_search,
orb_search_by_projection_kf, orb_search_for_triangulation, orb_distinctive_descriptors) against the numpy restatements of
oracle/oracle_map_py.py (equal to the reference's own lines: tests/test_oracle_map.py) and, where oracle/_ref/libmorb_ref_map.so
exists, against those lines directly. Index / integer work: everything must be EQUAL."""
import numpy as np
import pytest

from morb_slam_b200 import synth
from tests.conftest import has_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")]


@pytest.fixture(scope="module")
def env():
    from morb_slam_b200 import capi
    from oracle import oracle_py as op
    from oracle import oracle_match_py as om
    op.build()
    w, h, nf, lap, fx, b = synth.CONFIGS["euroc"]
    ex = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=4)
    exR = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=4)
    frames = []
    for s in (6100, 6101, 6102):
        L, R = synth.stereo_pair(s, w, h)
        _, kL, dL = ex(L, lap)
        _, kR, dR = exR(R, lap)
        uR, _ = capi.compute_stereo_matches(ex, exR, kL, dL, kR, dR, fx * b, fx)
        frames.append(dict(k=kL.copy(), d=dL.copy(), ur=uR.copy()))
    frames.append(dict(k=frames[0]["k"][:0], d=frames[0]["d"][:0], ur=frames[0]["ur"][:0]))     # an empty keyframe
    t = ex.tables()
    return dict(capi=capi, ex=ex, w=w, h=h, frames=frames, scale=t["scale"], sigma2=t["sigma2"], inv_sigma2=t["inv_sigma2"],
                bf=float(np.float32(fx * b)), gp=om.grid_params(w, h), gp_c=capi.grid_params(w, h))


def _load(env, stereo=True):
    capi, ex, fr = env["capi"], env["ex"], env["frames"]
    capi.load_frames(ex, [f["k"] for f in fr], [f["d"] for f in fr], [f["ur"] for f in fr] if stereo else None)
    capi.assign_features_to_grid(ex, env["gp_c"])


@pytest.mark.parametrize("mode,stereo,th", [(0, True, 3.0), (0, False, 3.0), (1, True, 4.0), (0, True, 7.5)])
def test_fuse_search(env, mode, stereo, th):
    from oracle import oracle_map_py as omap
    capi, ex, fr = env["capi"], env["ex"], env["frames"]
    _load(env, stereo)
    sets = [synth.synth_fuse_points(40 + i, f["k"], f["d"], env["w"], env["h"], nulls=mode == 0) for i, f in enumerate(fr)]
    qs = [omap.fuse_queries(s[0], env["bf"]) for s in sets]
    qcap = max(len(q) for q in qs) + 3
    Q = np.zeros((len(fr), qcap), capi.FQ_DTYPE); QD = np.zeros((len(fr), qcap, 32), np.uint8); nq = np.zeros(len(fr), np.int32)
    for i, q in enumerate(qs):
        nq[i] = len(q); Q[i, :len(q)] = q; QD[i, :len(q)] = sets[i][1]
    bi, bd = capi.fuse_search(ex, Q, QD, nq, th, mode)
    ref = omap.reference() if omap.have_reference() else None
    for i, f in enumerate(fr):
        ur = f["ur"] if stereo else None
        obi, obd = omap.fuse_search(f["k"], f["d"], ur, env["scale"], env["inv_sigma2"], env["gp"], qs[i], sets[i][1], th, mode)
        n = len(qs[i])
        assert np.array_equal(bi[i, :n], obi) and np.array_equal(bd[i, :n], obd), i
        assert np.all(bi[i, n:] == -1) and np.all(bd[i, n:] == 256)
        if len(f["k"]):
            assert (obi >= 0).sum() > 300
        if ref is not None:
            # the library's search + the replay of the map surgery == ORBmatcher::Fuse itself
            pts, pdesc, kf_nobs, kf_bad = sets[i]
            out_r = ref.fuse(f["k"], f["d"], ur, env["gp"], env["scale"], env["sigma2"], env["bf"], kf_nobs, kf_bad, pts, pdesc, th, mode == 1)
            out = omap.fuse_replay(pts, qs[i], bi[i, :n], bd[i, :n], kf_nobs, kf_bad, ur, env["gp"], mode == 1)
            assert out[0] == out_r[0] and out[1] == out_r[1]
            for a, b in zip(out[2:], out_r[2:]):
                assert np.array_equal(a, b)


@pytest.mark.parametrize("th,orb_dist,ori,lock", [(10.0, 100, True, 0.2), (3.0, 64, True, 0.0), (10.0, 100, False, 0.5), (25.0, 255, True, 0.1)])
def test_search_by_projection_keyframe(env, th, orb_dist, ori, lock):
    from oracle import oracle_map_py as omap
    capi, ex, fr = env["capi"], env["ex"], env["frames"]
    _load(env)
    rng = np.random.default_rng(int(th * 10) + orb_dist)
    qs = []
    for i, f in enumerate(fr):
        src = fr[(i + 1) % 3]       # the keyframe's map points were seen in another frame: ragged query counts
        q, qd = synth.synth_queries(50 + i, f["k"] if len(f["k"]) else src["k"], f["d"] if len(f["k"]) else src["d"], None, None, env["w"], env["h"], jitter=2.0)
        qd = synth.flip_bits(rng, qd, 60)
        perm = rng.permutation(len(q))[:1000 - 100 * i]
        q, qd = q[perm], qd[perm]
        q["flags"] = (q["flags"] & 1) & (rng.random(len(q)) > 0.1)
        qs.append((q, qd))
    qcap = max(len(q[0]) for q in qs)
    Q = np.zeros((len(fr), qcap), capi.Q_DTYPE); QD = np.zeros((len(fr), qcap, 32), np.uint8); nq = np.zeros(len(fr), np.int32)
    for i, (q, qd) in enumerate(qs):
        nq[i] = len(q); Q[i, :len(q)] = q; QD[i, :len(q)] = qd
    locked0 = (rng.random((len(fr), ex.kcap)) < lock).astype(np.uint8) if lock else None
    nm, match = capi.search_by_projection_kf(ex, Q, QD, nq, locked0, th, orb_dist, ori)
    ref = omap.reference() if omap.have_reference() else None
    for i, f in enumerate(fr):
        n = len(f["k"])
        lk = None if locked0 is None else locked0[i, :n]
        onm, om_ = omap.search_by_projection_kf(f["k"], f["d"], lk, env["scale"], env["gp"], qs[i][0], qs[i][1], th, orb_dist, ori)
        assert nm[i] == onm and np.array_equal(match[i, :n], om_), i
        assert np.all(match[i, n:] == -1)
        if n:
            assert onm > 100
        if ref is not None:
            rnm, rm = ref.search_by_projection_kf(f["k"], f["d"], lk, env["scale"], env["gp"], qs[i][0], qs[i][1], th, orb_dist, ori)
            assert nm[i] == rnm and np.array_equal(match[i, :n], rm)


@pytest.mark.parametrize("only_stereo,coarse,ori", [(False, False, True), (False, True, True), (True, False, True), (False, False, False)])
def test_search_for_triangulation(env, only_stereo, coarse, ori):
    from oracle import oracle_map_py as omap
    capi, ex, fr = env["capi"], env["ex"], env["frames"]
    kfs, pairs, Fs, eps = [], [], [], []
    for i in range(3):
        ur = fr[i]["ur"] if i != 1 else None
        k1, k2 = synth.synth_triangulation_pair(70 + i, fr[i]["k"], fr[i]["d"], ur, env["w"], env["h"])
        if ur is None:
            k1["uright"] = None; k2["uright"] = None
        kfs += [k1, k2]
        pairs.append((2 * i, 2 * i + 1))
        Fs.append(synth.synth_fundamental(70 + i)); eps.append((380.0, 240.0) if i == 2 else (5000.0, 240.0))
    pairs.append((1, 0)); Fs.append(synth.synth_fundamental(99)); eps.append((100.0, 100.0))   # reversed pair: ragged sizes, few matches
    empty = dict(kps=fr[3]["k"], desc=fr[3]["d"], uright=None, has_mp=np.zeros(0, np.uint8),
                 fv=dict(fv_node=np.zeros(0, np.uint32), fv_off=np.zeros(1, np.int32), fv_feat=np.zeros(0, np.uint32)))
    kfs.append(empty); pairs.append((6, 0)); Fs.append(Fs[0]); eps.append(eps[0])
    pairs.append((0, 6)); Fs.append(Fs[0]); eps.append(eps[0])
    # the set shares one mvuRight array: keyframes without one carry -1 everywhere
    for k in kfs:
        if k.get("uright") is None:
            k["uright"] = np.full(len(k["kps"]), -1, np.float32)
    nm, m12 = capi.search_for_triangulation(ex, kfs, pairs, np.array(Fs), np.array(eps, np.float32), only_stereo, coarse, ori)
    ref = omap.reference() if omap.have_reference() else None
    for p, (a, b) in enumerate(pairs):
        onm, om_ = omap.search_for_triangulation(kfs[a], kfs[b], env["scale"], env["sigma2"], Fs[p], eps[p], only_stereo, coarse, ori)
        n1 = len(kfs[a]["kps"])
        assert nm[p] == onm and np.array_equal(m12[p, :n1], om_), p
        assert np.all(m12[p, n1:] == -1)
        if p < 3 and not only_stereo:
            assert onm > 100
        if ref is not None:
            rnm, rm = ref.search_for_triangulation(kfs[a], kfs[b], env["gp"], env["scale"], env["sigma2"], Fs[p], eps[p], only_stereo, coarse, ori)
            assert nm[p] == rnm and np.array_equal(m12[p, :n1], rm)


def test_distinctive_descriptors(env):
    from oracle import oracle_map_py as omap
    capi, ex = env["capi"], env["ex"]
    obs = synth.synth_observations(5, 300)
    obs.append(synth.flip_bits(np.random.default_rng(1), np.zeros((700, 32), np.uint8), 120))     # a map point seen 700 times
    best, med = capi.distinctive_descriptors(ex, obs)
    ref = omap.reference() if omap.have_reference() else None
    for p, d in enumerate(obs):
        ob, omed = omap.distinctive(d)
        assert best[p] == ob and med[p] == omed, p
        if ref is not None and len(d) and len(d) <= 100:
            assert np.array_equal(d[best[p]], d[ref.distinctive(d)])


def test_loaded_frames_have_no_pyramid(env):
    capi, ex = env["capi"], env["ex"]
    _load(env)
    with pytest.raises(capi.OrbError) as e:
        ex.pyramid_level(0, 0)
    assert e.value.status == capi.ORB_ERR_STATE
    # the next extraction makes the handle an extractor again
    L, _ = synth.stereo_pair(6100, env["w"], env["h"])
    _, k, d = ex(L, (0, 0))
    assert k.tobytes() == env["frames"][0]["k"].tobytes() and np.array_equal(d, env["frames"][0]["d"])
    assert ex.pyramid_level(0, 0).shape == (env["h"], env["w"])


@pytest.mark.parametrize("th,ratio", [(8, 1.0), (4, 1.5), (30, 0.7)])
def test_search_by_projection_sim3(env, th, ratio):
    """ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) (src/ORBmatcher.cc:397-494)"""
    from oracle import oracle_map_py as omap
    from tests.test_oracle_map import _sim3_queries
    capi, ex, fr = env["capi"], env["ex"], env["frames"]
    _load(env)
    cases = []
    for i, f in enumerate(fr):
        src = f if len(f["k"]) else fr[0]
        cases.append(_sim3_queries(60 + i, dict(kL=src["k"], dL=src["d"], w=env["w"], h=env["h"])))
    qcap = max(len(c[0]) for c in cases)
    Q = np.zeros((len(fr), qcap), capi.Q_DTYPE); QD = np.zeros((len(fr), qcap, 32), np.uint8); nq = np.zeros(len(fr), np.int32)
    M0 = np.zeros((len(fr), ex.kcap), np.uint8)
    for i, c in enumerate(cases):
        nq[i] = len(c[1]); Q[i, :nq[i]] = c[1]; QD[i, :nq[i]] = c[2]
        n = len(fr[i]["k"])
        M0[i, :n] = c[5][:n]
    nm, match = capi.search_by_projection_sim3(ex, Q, QD, nq, M0, th, ratio)
    ref = omap.reference() if omap.have_reference() else None
    for i, f in enumerate(fr):
        n = len(f["k"])
        q, qdev, qd, found_slot, matched0, m0dev = cases[i]
        onm, om_ = omap.search_by_projection_sim3(f["k"], f["d"], M0[i, :n], env["scale"], env["gp"], qdev, qd, th, ratio)
        assert nm[i] == onm and np.array_equal(match[i, :n], om_), i
        if n:
            assert onm > 100
            if ref is not None:
                rnm, rm = ref.search_by_projection_sim3(f["k"], f["d"], matched0, env["gp"], env["scale"], env["sigma2"], q, qd, found_slot, th, ratio)
                assert nm[i] == rnm and np.array_equal(match[i, :n], rm)


@pytest.mark.parametrize("th", [7.5, 3.0])
def test_search_by_sim3(env, th):
    """ORBmatcher::SearchBySim3 (src/ORBmatcher.cc:1323-1519) = orb_fuse_search(mode 1) in both directions (one call: the two keyframes
    are the two resident frames) + the agreement test on the host"""
    from oracle import oracle_map_py as omap
    from tests.test_oracle_map import _sim3_case
    capi, ex = env["capi"], env["ex"]
    f0 = env["frames"][0]
    k1, d1, p1, pd1, k2, d2, p2, pd2, init12 = _sim3_case(5, dict(kL=f0["k"], dL=f0["d"], w=env["w"], h=env["h"]))
    capi.load_frames(ex, [k2, k1], [d2, d1])            # frame 0 = pKF2 (searched with pKF1's map points), frame 1 = pKF1
    capi.assign_features_to_grid(ex, env["gp_c"])
    gp = env["gp"]
    already1 = init12 >= 0
    already2 = np.zeros(len(k2), bool); already2[init12[already1]] = True
    q12, q21 = omap.sim3_queries(p1, already1), omap.sim3_queries(p2, already2)
    in_img = lambda q: (q["u"] >= gp[0]) & (q["u"] < gp[2]) & (q["v"] >= gp[1]) & (q["v"] < gp[3])
    q12["flags"] &= in_img(q12); q21["flags"] &= in_img(q21)
    qcap = max(len(q12), len(q21))
    Q = np.zeros((2, qcap), capi.FQ_DTYPE); QD = np.zeros((2, qcap, 32), np.uint8)
    Q[0, :len(q12)] = q12; QD[0, :len(q12)] = pd1; Q[1, :len(q21)] = q21; QD[1, :len(q21)] = pd2
    bi, bd = capi.fuse_search(ex, Q, QD, np.array([len(q12), len(q21)], np.int32), th, 1)
    nf, m = omap.search_by_sim3_compose(bi[0, :len(q12)], bd[0, :len(q12)], bi[1, :len(q21)], bd[1, :len(q21)], init12, None, None)
    assert nf > 100
    if omap.have_reference():
        nf_r, m_r = omap.reference().search_by_sim3(k1, d1, p1, pd1, k2, d2, p2, pd2, gp, env["scale"], env["sigma2"], init12, th)
        assert nf == nf_r and np.array_equal(m, m_r)
    b12, e12 = omap.fuse_search(k2, d2, None, env["scale"], env["inv_sigma2"], gp, q12, pd1, th, mode=1)
    assert np.array_equal(bi[0, :len(q12)], b12) and np.array_equal(bd[0, :len(q12)], e12)


@pytest.mark.parametrize("ratio,ori", [(0.75, True), (0.9, True), (0.6, False)])
def test_search_by_bow_keyframes(env, ratio, ori):
    """ORBmatcher::SearchByBoW(pKF1, pKF2, vpMatches12) (src/ORBmatcher.cc:702-819)"""
    from oracle import oracle_map_py as omap
    from tests.test_oracle_map import _bow_kf_pair
    capi, ex, fr = env["capi"], env["ex"], env["frames"]
    kfs, pairs = [], []
    for i in range(3):
        k1, k2 = _bow_kf_pair(80 + i, dict(kL=fr[i]["k"], dL=fr[i]["d"], w=env["w"], h=env["h"]))
        kfs += [k1, k2]
        pairs.append((2 * i, 2 * i + 1))
    pairs.append((1, 0)); pairs.append((0, 3))       # reversed, and two unrelated keyframes (few or no matches)
    nm, m12 = capi.search_by_bow_kf(ex, kfs, pairs, ratio, ori)
    ref = omap.reference() if omap.have_reference() else None
    for p, (a, b) in enumerate(pairs):
        onm, om_ = omap.search_by_bow_kf(kfs[a], kfs[b], ratio, ori)
        n1 = len(kfs[a]["kps"])
        assert nm[p] == onm and np.array_equal(m12[p, :n1], om_), p
        assert np.all(m12[p, n1:] == -1)
        if p < 3:
            assert onm > 100
        if ref is not None:
            rnm, rm = ref.search_by_bow_kf(kfs[a], kfs[b], env["gp"], ratio, ori)
            assert nm[p] == rnm and np.array_equal(m12[p, :n1], rm)


@pytest.mark.parametrize("window,ratio,ori,jit", [(100, 0.9, True, 0.0), (100, 0.9, False, 0.0), (30, 0.7, True, 4.0), (400, 0.9, True, 0.0),
                                                  (8, 1.0, True, 1.0)])
def test_search_for_initialization(env, window, ratio, ori, jit):
    """ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:603-700): F2 = the resident frames, F1 = queries; vnMatches12, the return
    value and the updated vbPrevMatched equal the restatement and the reference's own lines (window 400 holds more than SFI_CAP
    candidates per query: the on-the-fly path)"""
    from oracle import oracle_map_py as omap
    capi, ex, fr = env["capi"], env["ex"], env["frames"]
    # F2 of frame i = keyframe i; F1 = keyframe (i + 1) % 3 with duplicates (the same scene family: enough descriptors agree)
    sets = []
    for i in range(3):
        a, b = fr[(i + 1) % 3], fr[i]
        sets.append(synth.synth_init_frames(90 + i, a["k"], a["d"], b["k"], b["d"], prev_jitter=jit))
    sets.append((sets[0][0][:5], sets[0][1][:5], sets[0][2][:5], fr[3]["k"], fr[3]["d"]))     # empty F2
    capi.load_frames(ex, [s[3] for s in sets], [s[4] for s in sets])
    capi.assign_features_to_grid(ex, env["gp_c"])
    qcap = max(len(s[0]) for s in sets) + 2
    Q = np.zeros((len(sets), qcap), capi.IQ_DTYPE); QD = np.zeros((len(sets), qcap, 32), np.uint8); nq = np.zeros(len(sets), np.int32)
    for i, s in enumerate(sets):
        n = len(s[0]); nq[i] = n
        Q[i, :n]["x"] = s[2][:, 0]; Q[i, :n]["y"] = s[2][:, 1]; Q[i, :n]["angle"] = s[0]["angle"]; Q[i, :n]["octave"] = s[0]["octave"]
        QD[i, :n] = s[1]
    nm, m12, prev = capi.search_for_initialization(ex, Q, QD, nq, window, ratio, ori)
    for i, s in enumerate(sets):
        n = len(s[0])
        onm, om12, oprev = omap.search_for_initialization(s[0], s[1], s[2], s[3], s[4], env["gp"], window, ratio, ori)
        assert nm[i] == onm and np.array_equal(m12[i, :n], om12) and prev[i, :n].tobytes() == oprev.tobytes(), i
        assert np.all(m12[i, n:] == -1)
        if i < 3 and window >= 30:
            assert onm > 30
        if omap.have_reference():
            rnm, rm12, rprev = omap.ref_search_for_initialization(s[0], s[1], s[2], s[3], s[4], env["gp"], window, ratio, ori)
            assert nm[i] == rnm and np.array_equal(m12[i, :n], rm12) and prev[i, :n].tobytes() == rprev.tobytes(), i


def test_fuse_right_camera(env):
    """Fuse(pKF, vpMapPoints, th, bRight = true) of a two-camera keyframe = orb_fuse_search on the handle that holds the RIGHT camera's
    keypoints (no mvuRight: the 5.99 gate only) + NLeft on the host (src/ORBmatcher.cc:1173); the replay equals the reference's lines"""
    from oracle import oracle_map_py as omap
    capi, ex, fr = env["capi"], env["ex"], env["frames"]
    kL, dL = fr[0]["k"], fr[0]["d"]          # left camera of the keyframe
    kR, dR = fr[1]["k"], fr[1]["d"]          # right camera (any other keypoint set)
    nL = len(kL)
    capi.load_frames(ex, [kR], [dR])
    capi.assign_features_to_grid(ex, env["gp_c"])
    pts, pdesc, nobs_r, bad_r = synth.synth_fuse_points(77, kR, dR, env["w"], env["h"])
    q = omap.fuse_queries(pts, env["bf"])
    Q = np.zeros((1, len(q)), capi.FQ_DTYPE); Q[0] = q
    bi, bd = capi.fuse_search(ex, Q, pdesc[None], np.array([len(q)], np.int32), 3.0, 0)
    obi, obd = omap.fuse_search(kR, dR, None, env["scale"], env["inv_sigma2"], env["gp"], q, pdesc, 3.0, 0)
    assert np.array_equal(bi[0], obi) and np.array_equal(bd[0], obd) and (obi >= 0).sum() > 300
    if omap.have_reference():
        rng = np.random.default_rng(3)
        nobs = np.concatenate([np.where(rng.random(nL) < 0.5, rng.integers(1, 7, nL), -1).astype(np.int32), nobs_r])
        bad = np.concatenate([np.zeros(nL, np.uint8), bad_r])
        out_r = omap.reference().fuse_right(kL, dL, kR, dR, env["gp"], env["scale"], env["sigma2"], env["bf"], nobs, bad, pts, pdesc, 3.0)
        out = omap.fuse_replay(pts, q, np.where(bi[0] >= 0, bi[0] + nL, -1), bd[0], nobs, bad, None, env["gp"], False)
        assert out[0] == out_r[0] and out[1] == out_r[1]
        for a, b in zip(out[2:], out_r[2:]):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("kw", [{}, dict(coarse=True), dict(check_orientation=False), dict(only_stereo=True)])
def test_search_for_triangulation_two_camera(env, kw):
    """ORBmatcher::SearchForTriangulation between two-camera keyframes (mpCamera2 branch, KannalaBrandt8::epipolarConstrain =
    TriangulateMatches > 0.0001f per candidate with the cameras and relative pose of the left / right combination): equal to the
    restatement and to the reference's own lines; three pairs in one call, incl. a reversed one"""
    from oracle import oracle_map_py as omap
    capi, ex = env["capi"], env["ex"]
    sets = [synth.synth_two_camera_keyframes(20 + i) for i in range(2)]
    kfs = [sets[0][0], sets[0][1], sets[1][0], sets[1][1]]
    pairs = [(0, 1), (2, 3), (1, 0)]
    inv = sets[0][2].copy()          # the reversed pair needs the inverse relative poses: x2 = R^T x1 - R^T t, combos transposed (lr <-> rl)
    for c, src in enumerate((0, 2, 1, 3)):
        R = sets[0][2][src]["R12"].reshape(3, 3).astype(np.float64); t = sets[0][2][src]["t12"].astype(np.float64)
        inv[c]["R12"] = R.T.astype(np.float32).reshape(9); inv[c]["t12"] = (-R.T @ t).astype(np.float32)
        inv[c]["cam1"], inv[c]["cam2"] = sets[0][2][src]["cam2"], sets[0][2][src]["cam1"]
        inv[c]["prec1"], inv[c]["prec2"] = sets[0][2][src]["prec2"], sets[0][2][src]["prec1"]
    rigs = np.stack([sets[0][2], sets[1][2], inv])
    nm, m12 = capi.search_for_triangulation_fisheye(ex, kfs, pairs, rigs, **kw)
    for p, (a, b) in enumerate(pairs):
        onm, om12 = omap.search_for_triangulation_fisheye(kfs[a], kfs[b], env["sigma2"], rigs[p], **kw)
        n1 = len(kfs[a]["kps"])
        assert nm[p] == onm and np.array_equal(m12[p, :n1], om12), p
        assert np.all(m12[p, n1:] == -1)
        if not kw.get("only_stereo"):
            assert onm > 100
        if omap.have_reference_sft2():
            rnm, rm12 = omap.ref_search_for_triangulation_fisheye(kfs[a], kfs[b], env["scale"], env["sigma2"], rigs[p], **kw)
            assert nm[p] == rnm and np.array_equal(m12[p, :n1], rm12), p
