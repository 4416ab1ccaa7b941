"""LocalMapping / Relocalization matchers (widening beyond SURVEY.md 8): the numpy restatements of oracle/oracle_map_py.py against the
reference's own code compiled by line range (oracle/_ref/libmorb_ref_map.so: src/ORBmatcher.cc:821-1042, 1044-1322, 1735-1842,
src/KeyFrame.cc:729-778, src/CameraModels/Pinhole.cpp:125-138, src/MapPoint.cc:367-435). CPU only; skipped where /root/reference was
never mounted."""
import numpy as np
import pytest

from morb_slam_b200 import synth
from oracle import oracle_py as op
from oracle import oracle_match_py as om
from oracle import oracle_map_py as omap

pytestmark = pytest.mark.skipif(not omap.have_reference(), reason="oracle/_ref/libmorb_ref_map.so not built (no /root/reference)")


@pytest.fixture(scope="module")
def frames():
    op.build()
    w, h, nf, lap, fx, b = synth.CONFIGS["euroc"]
    L, R = synth.stereo_pair(5200, w, h)
    oL, oR = op.OracleExtractor(nf), op.OracleExtractor(nf)
    _, kL, dL = oL(L, lap)
    _, kR, dR = oR(R, lap)
    uR, _ = op.oracle_stereo(oL, oR, kL, dL, kR, dR, float(np.float32(fx * b)), float(np.float32(fx)))
    t = oL.tables()
    return dict(w=w, h=h, kL=kL, dL=dL, kR=kR, dR=dR, uR=uR, scale=t["scale"], sigma2=t["sigma2"], inv_sigma2=t["inv_sigma2"],
                bf=float(np.float32(fx * b)))


def test_keyframe_features_in_area_is_the_frame_one_without_levels(frames):
    """KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:729-774) == Frame::GetFeaturesInArea(x, y, r, -1, -1) on the same grid"""
    gp = om.grid_params(frames["w"], frames["h"])
    o, r = om.oracle(), omap.reference()
    rng = np.random.default_rng(11)
    for kps in (frames["kL"], frames["kL"][:1], frames["kL"][:0]):
        for _ in range(150):
            x, y = rng.uniform(-30, frames["w"] + 30), rng.uniform(-30, frames["h"] + 30)
            rad = rng.choice([0.5, 3, 7, 15, 40, 120, 2000])
            assert np.array_equal(o.features_in_area(kps, gp, x, y, rad, -1, -1), r.features_in_area(kps, gp, x, y, rad))


@pytest.mark.parametrize("seed,th,sim3,stereo", [(1, 3.0, False, True), (2, 3.0, False, False), (3, 6.0, False, True), (4, 4.0, True, True),
                                                 (5, 3.0, True, False)])
def test_fuse_search_and_replay_equal_reference(frames, seed, th, sim3, stereo):
    """device-shaped search (restated) + the sequential replay of the map surgery == ORBmatcher::Fuse run on the map model"""
    gp = om.grid_params(frames["w"], frames["h"])
    kps, desc = frames["kL"], frames["dL"]
    ur = frames["uR"] if stereo else None
    pts, pdesc, kf_nobs, kf_bad = synth.synth_fuse_points(seed, kps, desc, frames["w"], frames["h"], nulls=not sim3)
    nf_r, ev_r, repl_r, kf_r, cb_r, cn_r = omap.reference().fuse(kps, desc, ur, gp, frames["scale"], frames["sigma2"], frames["bf"], kf_nobs,
                                                                 kf_bad, pts, pdesc, th, sim3)
    q = omap.fuse_queries(pts, frames["bf"])
    bi, bd = omap.fuse_search(kps, desc, ur, frames["scale"], frames["inv_sigma2"], gp, q, pdesc, th, mode=1 if sim3 else 0)
    nf, ev, repl, kf, cb, cn = omap.fuse_replay(pts, q, bi, bd, kf_nobs, kf_bad, ur, gp, sim3)
    assert nf == nf_r and nf > 50
    assert ev == ev_r
    assert np.array_equal(repl, repl_r) and np.array_equal(kf, kf_r) and np.array_equal(cb, cb_r) and np.array_equal(cn, cn_r)
    if not sim3:
        kinds = [e[0] for e in ev]
        assert kinds.count(1) > 10 and kinds.count(2) > 10            # both new observations and replacements happen
        assert any(e[0] == 2 and e[1] >= 0 for e in ev) and any(e[0] == 2 and e[1] <= -2 for e in ev)   # in both directions
    else:
        assert (repl >= 0).sum() > 10


def test_fuse_on_an_empty_keyframe_and_without_candidates(frames):
    gp = om.grid_params(frames["w"], frames["h"])
    r = omap.reference()
    pts, pdesc, _, _ = synth.synth_fuse_points(9, frames["kL"], frames["dL"], frames["w"], frames["h"])
    e = np.zeros(0, np.int32)
    nf, ev, *_ = r.fuse(frames["kL"][:0], frames["dL"][:0], None, gp, frames["scale"], frames["sigma2"], 40.0, e, e.astype(np.uint8), pts, pdesc, 3.0)
    q = omap.fuse_queries(pts, 40.0)
    bi, bd = omap.fuse_search(frames["kL"][:0], frames["dL"][:0], None, frames["scale"], frames["inv_sigma2"], gp, q, pdesc, 3.0)
    assert nf == 0 and not ev and np.all(bi == -1) and np.all(bd == 256)
    kf_nobs = np.full(len(frames["kL"]), -1, np.int32)
    nf, ev, *_ = r.fuse(frames["kL"], frames["dL"], None, gp, frames["scale"], frames["sigma2"], 40.0, kf_nobs, np.zeros(len(kf_nobs), np.uint8),
                        pts[:0], pdesc[:0], 3.0)
    assert nf == 0 and not ev


def _kf_queries(seed, frames, p_bad=0.05, p_found=0.1):
    q, qd = synth.synth_queries(seed, frames["kL"], frames["dL"], None, None, frames["w"], frames["h"], jitter=2.0)
    rng = np.random.default_rng(seed + 77)
    qd = synth.flip_bits(rng, qd, 60)
    perm = rng.permutation(len(q))[:900]          # the keyframe's map points come in its own keypoint order, not the frame's
    q, qd = q[perm], qd[perm]
    n = len(q)
    q["flags"] = (q["flags"] & 1) | np.where(rng.random(n) < p_bad, 4, 0) | np.where(rng.random(n) < p_found, 8, 0)
    qdev = q.copy()
    qdev["flags"] = ((q["flags"] & 1) != 0) & ((q["flags"] & 12) == 0)
    return q, qdev, qd


@pytest.mark.parametrize("seed,th,orb_dist,ori,lock", [(1, 10.0, 100, True, 0.2), (2, 3.0, 64, True, 0.0), (3, 10.0, 100, False, 0.5),
                                                       (4, 25.0, 255, True, 0.1)])
def test_search_by_projection_keyframe_equals_reference(frames, seed, th, orb_dist, ori, lock):
    gp = om.grid_params(frames["w"], frames["h"])
    kps, desc = frames["kL"], frames["dL"]
    q, qdev, qd = _kf_queries(seed, frames)
    locked0 = (np.random.default_rng(seed).random(len(kps)) < lock).astype(np.uint8) if lock else None
    nm_r, m_r = omap.reference().search_by_projection_kf(kps, desc, locked0, frames["scale"], gp, q, qd, th, orb_dist, ori)
    nm, m = omap.search_by_projection_kf(kps, desc, locked0, frames["scale"], gp, qdev, qd, th, orb_dist, ori)
    assert nm == nm_r and np.array_equal(m, m_r) and nm > 30


@pytest.mark.parametrize("seed,only_stereo,coarse,ori,ep", [(1, False, False, True, (5000.0, 240.0)), (2, False, True, True, (300.0, 200.0)),
                                                            (3, True, False, True, (300.0, 200.0)), (4, False, False, False, (380.0, 240.0)),
                                                            (5, False, False, True, (380.0, 240.0))])
def test_search_for_triangulation_equals_reference(frames, seed, only_stereo, coarse, ori, ep):
    gp = om.grid_params(frames["w"], frames["h"])
    ur = frames["uR"] if seed != 4 else None
    k1, k2 = synth.synth_triangulation_pair(seed, frames["kL"], frames["dL"], ur, frames["w"], frames["h"])
    F = synth.synth_fundamental(seed)
    nm_r, m_r = omap.reference().search_for_triangulation(k1, k2, gp, frames["scale"], frames["sigma2"], F, ep, only_stereo, coarse, ori)
    nm, m = omap.search_for_triangulation(k1, k2, frames["scale"], frames["sigma2"], F, ep, only_stereo, coarse, ori)
    assert nm == nm_r and np.array_equal(m, m_r)
    assert nm > (5 if only_stereo else 40)


def test_distinctive_descriptors_equal_reference():
    r = omap.reference()
    for p, d in enumerate(synth.synth_observations(3, 60)):
        best, med = omap.distinctive(d)
        ref = r.distinctive(d)
        if len(d) == 0:
            assert best == -1 and ref == -1
        else:
            # the reference reports the descriptor it kept: equal content (duplicates share it)
            assert np.array_equal(d[best], d[ref]), (p, best, ref)


def _sim3_queries(seed, frames, p_bad=0.05, p_found=0.08):
    """candidates of SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratio): (queries for the driver, for the library, descriptors,
    found_slot, matched0)"""
    kps, desc = frames["kL"], frames["dL"]
    rng = np.random.default_rng(seed)
    q, qd = synth.synth_queries(seed, kps, desc, None, None, frames["w"], frames["h"], jitter=2.0)
    qd = synth.flip_bits(rng, qd, 50)
    perm = rng.permutation(len(q))[:900]
    q, qd = q[perm], qd[perm]
    n = len(q)
    q["z"] = np.abs(q["z"]) + 1
    q["octave"] = np.clip(q["octave"] + rng.choice([0, 0, 1], n), 0, 7)
    q["flags"] = 1 | np.where(rng.random(n) < p_bad, 4, 0)
    matched0 = (rng.random(len(kps)) < 0.25).astype(np.uint8)
    found_slot = np.full(n, -1, np.int32)
    free = np.nonzero(matched0 == 0)[0]
    take = rng.permutation(len(free))[:int(n * p_found)]
    found_slot[rng.permutation(n)[:len(take)]] = free[take]
    qdev = q.copy()
    qdev["flags"] = ((q["flags"] & 4) == 0) & (found_slot < 0)
    m0dev = matched0.copy()
    m0dev[found_slot[found_slot >= 0]] = 1
    return q, qdev, qd, found_slot, matched0, m0dev


@pytest.mark.parametrize("seed,th,ratio", [(1, 8, 1.0), (2, 4, 1.5), (3, 30, 0.7), (4, 8, 2.0)])
def test_search_by_projection_sim3_equals_reference(frames, seed, th, ratio):
    gp = om.grid_params(frames["w"], frames["h"])
    q, qdev, qd, found_slot, matched0, m0dev = _sim3_queries(seed, frames)
    nm_r, m_r = omap.reference().search_by_projection_sim3(frames["kL"], frames["dL"], matched0, gp, frames["scale"], frames["sigma2"], q, qd,
                                                           found_slot, th, ratio)
    nm, m = omap.search_by_projection_sim3(frames["kL"], frames["dL"], m0dev, frames["scale"], gp, qdev, qd, th, ratio)
    assert nm == nm_r and np.array_equal(m, m_r) and nm > 100


def _sim3_points(seed, k_from, d_from, w, h, p_mp=0.8, jitter=2.0):
    rng = np.random.default_rng(seed)
    n = len(k_from)
    p = np.zeros(n, omap.S3_DTYPE)
    p["u"] = k_from["x"] + rng.normal(0, jitter, n).astype(np.float32)
    p["v"] = k_from["y"] + rng.normal(0, jitter, n).astype(np.float32)
    p["level"] = np.clip(k_from["octave"] + rng.choice([0, 0, 1], n), 0, 7)
    p["flags"] = (rng.random(n) < p_mp).astype(np.int32) | np.where(rng.random(n) < 0.05, 2, 0)
    outside = rng.random(n) < 0.03
    p["u"][outside] = np.float32(w + 9)
    return p, synth.flip_bits(rng, d_from, 40)


def _sim3_case(seed, frames):
    """two keyframes looking at the same points: pKF2 = pKF1's keypoints shuffled and moved by a few pixels"""
    rng = np.random.default_rng(seed)
    k1, d1 = frames["kL"], frames["dL"]
    perm = rng.permutation(len(k1))
    k2 = np.array(k1[perm], copy=True)
    k2["x"] += rng.normal(0, 1.0, len(k2)).astype(np.float32); k2["y"] += rng.normal(0, 1.0, len(k2)).astype(np.float32)
    d2 = synth.flip_bits(rng, d1[perm], 25)
    p1, pd1 = _sim3_points(seed + 1, k1, d1, frames["w"], frames["h"])      # map points of pKF1, projected into pKF2 (same place)
    p2, pd2 = _sim3_points(seed + 2, k2, d2, frames["w"], frames["h"])
    init12 = np.full(len(k1), -1, np.int32)
    inv = np.argsort(perm)
    pick = rng.random(len(k1)) < 0.1
    init12[pick] = inv[pick]
    init12[(p2["flags"][np.maximum(init12, 0)] & 1) == 0] = -1             # an initial match needs a map point in pKF2
    return k1, d1, p1, pd1, k2, d2, p2, pd2, init12


@pytest.mark.parametrize("seed,th", [(1, 7.5), (2, 3.0)])
def test_search_by_sim3_is_two_fuse_searches_and_an_agreement_test(frames, seed, th):
    gp = om.grid_params(frames["w"], frames["h"])
    k1, d1, p1, pd1, k2, d2, p2, pd2, init12 = _sim3_case(seed, frames)
    nf_r, m_r = omap.reference().search_by_sim3(k1, d1, p1, pd1, k2, d2, p2, pd2, gp, frames["scale"], frames["sigma2"], init12, th)
    already1 = init12 >= 0
    already2 = np.zeros(len(k2), bool); already2[init12[already1]] = True
    q12, q21 = omap.sim3_queries(p1, already1), omap.sim3_queries(p2, already2)
    in_img = lambda q: (q["u"] >= gp[0]) & (q["u"] < gp[2]) & (q["v"] >= gp[1]) & (q["v"] < gp[3])
    q12["flags"] &= in_img(q12); q21["flags"] &= in_img(q21)
    b12, e12 = omap.fuse_search(k2, d2, None, frames["scale"], frames["inv_sigma2"], gp, q12, pd1, th, mode=1)
    b21, e21 = omap.fuse_search(k1, d1, None, frames["scale"], frames["inv_sigma2"], gp, q21, pd2, th, mode=1)
    nf, m = omap.search_by_sim3_compose(b12, e12, b21, e21, init12, None, None)
    assert nf == nf_r and np.array_equal(m, m_r) and nf > 100


def _bow_kf_pair(seed, frames):
    k1, k2 = synth.synth_triangulation_pair(seed, frames["kL"], frames["dL"], None, frames["w"], frames["h"], p_mp=0.75)
    rng = np.random.default_rng(seed + 5)
    for k in (k1, k2):
        st = k["has_mp"].astype(np.uint8)
        st[(st == 1) & (rng.random(len(st)) < 0.06)] = 2      # a few bad map points
        k["mp_state"] = st
        k["has_mp"] = (st == 1).astype(np.uint8)
    return k1, k2


@pytest.mark.parametrize("seed,ratio,ori", [(1, 0.75, True), (2, 0.9, True), (3, 0.6, False)])
def test_search_by_bow_keyframes_equals_reference(frames, seed, ratio, ori):
    gp = om.grid_params(frames["w"], frames["h"])
    k1, k2 = _bow_kf_pair(seed, frames)
    nm_r, m_r = omap.reference().search_by_bow_kf(k1, k2, gp, ratio, ori)
    nm, m = omap.search_by_bow_kf(k1, k2, ratio, ori)
    assert nm == nm_r and np.array_equal(m, m_r) and nm > 100


@pytest.mark.parametrize("seed,window,ratio,ori,jit", [(1, 100, 0.9, True, 0.0), (2, 100, 0.9, False, 0.0), (3, 30, 0.7, True, 4.0),
                                                       (4, 400, 0.9, True, 0.0), (5, 8, 1.0, True, 1.0)])
def test_search_for_initialization_equals_reference(frames, seed, window, ratio, ori, jit):
    """restatement of ORBmatcher::SearchForInitialization == the reference's own lines (src/ORBmatcher.cc:603-700,
    oracle/_ref/libmorb_ref_match.so) incl. take-overs, `<=` skips on ties and the records of stolen matches in the histogram"""
    gp = om.grid_params(frames["w"], frames["h"])
    k1, d1, prev, k2, d2 = synth.synth_init_frames(seed, frames["kL"], frames["dL"], frames["kR"], frames["dR"], prev_jitter=jit)
    nm_r, m_r, p_r = omap.ref_search_for_initialization(k1, d1, prev, k2, d2, gp, window, ratio, ori)
    nm, m, p = omap.search_for_initialization(k1, d1, prev, k2, d2, gp, window, ratio, ori)
    assert nm == nm_r and np.array_equal(m, m_r) and p.tobytes() == p_r.tobytes()
    assert nm == int((m >= 0).sum())
    if window >= 30:
        assert nm > 40
    # empty frames
    for a, b in ((0, len(k2)), (len(k1), 0)):
        r1 = omap.ref_search_for_initialization(k1[:a], d1[:a], prev[:a], k2[:b], d2[:b], gp, window, ratio, ori)
        r2 = omap.search_for_initialization(k1[:a], d1[:a], prev[:a], k2[:b], d2[:b], gp, window, ratio, ori)
        assert r1[0] == r2[0] == 0 and np.array_equal(r1[1], r2[1]) and r1[2].tobytes() == r2[2].tobytes()


@pytest.mark.parametrize("seed,th", [(11, 3.0), (12, 6.0)])
def test_fuse_right_camera_equals_reference(frames, seed, th):
    """Fuse(pKF, vpMapPoints, th, bRight = true) on a two-camera keyframe (src/ORBmatcher.cc:1050-1053, :1134, :1145, :1173) == the
    single-camera search on the RIGHT keypoints (mvuRight = -1 everywhere: only the 5.99 gate) with the match moved to NLeft + idx,
    followed by the same replay"""
    gp = om.grid_params(frames["w"], frames["h"])
    kL, dL, kR, dR = frames["kL"], frames["dL"], frames["kR"], frames["dR"]
    nL = len(kL)
    pts, pdesc, nobs_r, bad_r = synth.synth_fuse_points(seed, kR, dR, frames["w"], frames["h"])
    rng = np.random.default_rng(seed)
    nobs = np.concatenate([np.where(rng.random(nL) < 0.5, rng.integers(1, 7, nL), -1).astype(np.int32), nobs_r])
    bad = np.concatenate([np.zeros(nL, np.uint8), bad_r])
    out_r = omap.reference().fuse_right(kL, dL, kR, dR, gp, frames["scale"], frames["sigma2"], frames["bf"], nobs, bad, pts, pdesc, th)
    q = omap.fuse_queries(pts, frames["bf"])
    bi, bd = omap.fuse_search(kR, dR, None, frames["scale"], frames["inv_sigma2"], gp, q, pdesc, th, mode=0)
    out = omap.fuse_replay(pts, q, np.where(bi >= 0, bi + nL, -1), bd, nobs, bad, None, gp, False)
    assert out[0] == out_r[0] and out[0] > 50 and out[1] == out_r[1]
    for a, b in zip(out[2:], out_r[2:]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("seed,kw", [(1, {}), (2, {}), (3, dict(coarse=True)), (4, dict(check_orientation=False)), (5, dict(only_stereo=True))])
def test_search_for_triangulation_two_camera_equals_reference(seed, kw):
    """the mpCamera2 branch of ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:891-903, :935-981) with
    KannalaBrandt8::epipolarConstrain (src/CameraModels/KannalaBrandt8.cpp:229-236): restatement == the reference's own lines
    (oracle/_ref/libmorb_ref_sft2.so) on two fisheye keyframes whose four camera combinations all carry matches"""
    if not omap.have_reference_sft2():
        pytest.skip("oracle/_ref/libmorb_ref_sft2.so not built")
    op.build()
    t = op.OracleExtractor(1500).tables()
    k1, k2, rigs = synth.synth_two_camera_keyframes(seed)
    nm, m = omap.search_for_triangulation_fisheye(k1, k2, t["sigma2"], rigs, **kw)
    nr, mr = omap.ref_search_for_triangulation_fisheye(k1, k2, t["scale"], t["sigma2"], rigs, **kw)
    assert nm == nr and np.array_equal(m, mr)
    if kw.get("only_stereo"):
        assert nm == 0            # bStereo1 is false with a second camera: the reference matches nothing
        return
    assert nm > 150
    hit = np.nonzero(m >= 0)[0]
    combos = np.bincount(2 * (hit >= k1["nleft"]) + (m[hit] >= k2["nleft"]), minlength=4)
    assert combos.min() > 20      # left-left, left-right, right-left, right-right
    if not kw:
        nc, _ = omap.search_for_triangulation_fisheye(k1, k2, t["sigma2"], rigs, coarse=True)
        assert nc > nm            # the triangulation gate rejects some candidates
