"""Windowed matcher: the CPU restatement (oracle/orb_oracle_match.cc) against the reference's
own code compiled by line range (oracle/_ref/libmorb_ref_match.so: src/Frame.cc:501-528,742-820 and
src/ORBmatcher.cc:1521-1733,1844-1876). CPU only; skipped where /root/reference was never mounted."""
import numpy as np
import pytest

from morb_slam_b200 import synth
from oracle import oracle_py as op
from oracle import oracle_match_py as om

pytestmark = pytest.mark.skipif(not om.have_reference(), reason="oracle/_ref/libmorb_ref_match.so not built (no /root/reference)")


@pytest.fixture(scope="module")
def frames():
    op.build()
    w, h, nf, lap, fx, b = synth.CONFIGS["euroc"]
    L, R = synth.stereo_pair(4100, w, h)
    oL, oR = op.OracleExtractor(nf), op.OracleExtractor(nf)
    _, kL, dL = oL(L, lap)
    _, kR, dR = oR(R, lap)
    uR, _ = op.oracle_stereo(oL, oR, kL, dL, kR, dR, float(np.float32(fx * b)), float(np.float32(fx)))
    scale = oL.tables()["scale"]
    return dict(w=w, h=h, kL=kL, dL=dL, kR=kR, dR=dR, uR=uR, scale=scale, mbf=float(np.float32(fx * b)), mb=float(np.float32(b)))


def test_grid_equals_reference(frames):
    gp = om.grid_params(frames["w"], frames["h"])
    o, r = om.oracle(), om.reference()
    for kps in (frames["kL"], frames["kR"], frames["kL"][:1], frames["kL"][:0]):
        oo, oi = o.assign_grid(kps, gp)
        ro, ri = r.assign_grid(kps, gp)
        assert np.array_equal(oo, ro) and np.array_equal(oi, ri)
    # keypoints outside the image bounds (undistorted coordinates may leave the image, :814-818)
    k = frames["kL"].copy()
    k["x"][::7] -= 40
    k["y"][::5] += 300
    oo, oi = o.assign_grid(k, gp)
    ro, ri = r.assign_grid(k, gp)
    assert np.array_equal(oo, ro) and np.array_equal(oi, ri) and len(oi) < len(k)


def test_features_in_area_equals_reference(frames):
    gp = om.grid_params(frames["w"], frames["h"])
    o, r = om.oracle(), om.reference()
    rng = np.random.default_rng(5)
    kps = frames["kL"]
    for _ in range(300):
        x, y = rng.uniform(-30, frames["w"] + 30), rng.uniform(-30, frames["h"] + 30)
        rad = rng.choice([0.5, 3, 7, 15, 40, 120, 2000])
        lo, hi = rng.choice([-1, 0, 1, 3]), rng.choice([-1, 0, 2, 7])
        assert np.array_equal(o.features_in_area(kps, gp, x, y, rad, lo, hi), r.features_in_area(kps, gp, x, y, rad, lo, hi))


CASES = [
    # th, mono, tlc_z, check_orientation, jitter, p_obs
    (7.0, False, 0.0, True, 4.0, 0.8),      # stereo, neither forward nor backward
    (15.0, True, 0.0, True, 8.0, 0.8),      # monocular
    (7.0, False, 0.5, True, 4.0, 0.8),      # forward: tlc_z > mb
    (7.0, False, -0.5, True, 4.0, 0.8),     # backward
    (15.0, False, 0.0, False, 10.0, 0.3),   # no orientation check, mostly unlocked points (overwrites)
    (30.0, False, 0.0, True, 2.0, 1.0),     # large windows, every point locks
]


@pytest.mark.parametrize("th,mono,tlc,ori,jit,pobs", CASES)
def test_search_by_projection_equals_reference(frames, th, mono, tlc, ori, jit, pobs):
    f = frames
    gp = om.grid_params(f["w"], f["h"])
    o, r = om.oracle(), om.reference()
    for seed in range(3):
        # last frame = left image keypoints, current frame = right image keypoints (same scene, shifted)
        q, qd = om.synth_queries(seed, f["kL"], f["dL"], f["kR"], f["dR"], f["w"], f["h"], p_obs=pobs, jitter=jit)
        ur = f["uR"] if seed != 1 else np.full(len(f["kR"]), -1, np.float32)
        ur = np.resize(ur, len(f["kR"])).astype(np.float32)
        no, mo = o.search_by_projection(f["kR"], f["dR"], ur, f["scale"], gp, f["mb"], f["mbf"], q, qd, th, mono, tlc, ori)
        nr, mr = r.search_by_projection(f["kR"], f["dR"], ur, f["scale"], gp, f["mb"], f["mbf"], q, qd, th, mono, tlc, ori)
        assert no == nr and np.array_equal(mo, mr)
        assert nr > 50  # the case really matches something


def test_search_by_projection_degenerate(frames):
    f = frames
    gp = om.grid_params(f["w"], f["h"])
    o, r = om.oracle(), om.reference()
    ur = np.full(len(f["kR"]), -1, np.float32)
    # identical frames: every query sits on its own keypoint with distance 0; all-duplicate descriptors (ties)
    q, qd = om.synth_queries(9, f["kR"], f["dR"], f["kR"], f["dR"], f["w"], f["h"], p_valid=1.0, p_obs=1.0, jitter=0.0, p_dup=0.0)
    for qdesc in (qd, np.zeros_like(qd)):
        for dC in (f["dR"], np.zeros_like(f["dR"])):
            no, mo = o.search_by_projection(f["kR"], dC, ur, f["scale"], gp, f["mb"], f["mbf"], q, qdesc, 7.0)
            nr, mr = r.search_by_projection(f["kR"], dC, ur, f["scale"], gp, f["mb"], f["mbf"], q, qdesc, 7.0)
            assert no == nr and np.array_equal(mo, mr)
    # no valid query, no current keypoint
    q0 = q.copy(); q0["flags"] = 0
    assert o.search_by_projection(f["kR"], f["dR"], ur, f["scale"], gp, f["mb"], f["mbf"], q0, qd, 7.0)[0] == 0
    assert r.search_by_projection(f["kR"], f["dR"], ur, f["scale"], gp, f["mb"], f["mbf"], q0, qd, 7.0)[0] == 0
    no, mo = o.search_by_projection(f["kR"][:0], f["dR"][:0], ur[:0], f["scale"], gp, f["mb"], f["mbf"], q, qd, 7.0)
    nr, mr = r.search_by_projection(f["kR"][:0], f["dR"][:0], ur[:0], f["scale"], gp, f["mb"], f["mbf"], q, qd, 7.0)
    assert no == nr == 0


LOCAL_CASES = [
    # th, nnratio, jitter, p_obs, share of keypoints locked before the call
    (1.0, 0.8, 2.0, 0.9, 0.3),     # bFactor false (th == 1)
    (3.0, 0.8, 3.0, 0.9, 0.3),     # Tracking::SearchLocalPoints default
    (5.0, 0.8, 6.0, 0.9, 0.0),     # RGBD / recently relocalised
    (15.0, 0.9, 10.0, 0.3, 0.5),   # coarse search, mostly unobserved points (overwrites)
    (3.0, 0.6, 1.0, 1.0, 0.0),     # strict ratio
]


@pytest.mark.parametrize("th,ratio,jit,pobs,plock", LOCAL_CASES)
def test_search_local_points_equals_reference(frames, th, ratio, jit, pobs, plock):
    f = frames
    gp = om.grid_params(f["w"], f["h"])
    o, r = om.oracle(), om.reference()
    for seed in range(3):
        # searched frame = right image; its uRight = the left image's stereo result resized (values only need to be plausible)
        ur = np.resize(f["uR"], len(f["kR"])).astype(np.float32) if seed != 1 else np.full(len(f["kR"]), -1, np.float32)
        q, qd = om.synth_track_queries(seed, f["kR"], f["dR"], ur, f["w"], f["h"], p_obs=pobs, jitter=jit, mbf=f["mbf"])
        rng = np.random.default_rng(100 + seed)
        locked0 = (rng.random(len(f["kR"])) < plock).astype(np.uint8)
        no, mo = o.search_local_points(f["kR"], f["dR"], ur, locked0, f["scale"], gp, q, qd, th, ratio)
        nr, mr = r.search_local_points(f["kR"], f["dR"], ur, locked0, f["scale"], gp, q, qd, th, ratio)
        assert no == nr and np.array_equal(mo, mr)
        assert nr > 100


def test_search_local_points_degenerate(frames):
    f = frames
    gp = om.grid_params(f["w"], f["h"])
    o, r = om.oracle(), om.reference()
    ur = np.full(len(f["kR"]), -1, np.float32)
    none = np.zeros(len(f["kR"]), np.uint8)
    q, qd = om.synth_track_queries(3, f["kR"], f["dR"], ur, f["w"], f["h"], n_extra=1.0, p_view=1.0, p_obs=1.0, jitter=0.0, p_dup=0.0)
    # identical descriptors everywhere: best == second, the ratio test rejects same-level pairs, ties keep the first
    for qdesc in (qd, np.zeros_like(qd)):
        for dC in (f["dR"], np.zeros_like(f["dR"])):
            for locked0 in (none, np.ones_like(none)):
                no, mo = o.search_local_points(f["kR"], dC, ur, locked0, f["scale"], gp, q, qdesc, 3.0)
                nr, mr = r.search_local_points(f["kR"], dC, ur, locked0, f["scale"], gp, q, qdesc, 3.0)
                assert no == nr and np.array_equal(mo, mr)
    no, mo = o.search_local_points(f["kR"][:0], f["dR"][:0], ur[:0], none[:0], f["scale"], gp, q, qd, 3.0)
    nr, mr = r.search_local_points(f["kR"][:0], f["dR"][:0], ur[:0], none[:0], f["scale"], gp, q, qd, 3.0)
    assert no == nr == 0


BOW_CASES = [
    # vocabulary (k, L), levelsup, nnratio, check orientation, share of keyframe keypoints with a map point
    ((10, 4), 2, 0.7, True, 0.8),      # TrackReferenceKeyFrame: ORBmatcher(0.7, true)
    ((10, 4), 3, 0.75, True, 1.0),     # Relocalization: ORBmatcher(0.75, true); coarser nodes -> larger groups
    ((6, 3), 3, 0.9, False, 0.5),      # everything in one node (levelsup = L): all pairs
    ((10, 4), 0, 0.7, True, 0.8),      # nodes = words: tiny groups
]


@pytest.mark.parametrize("kl,levelsup,ratio,ori,pmp", BOW_CASES)
def test_search_by_bow_equals_reference(frames, kl, levelsup, ratio, ori, pmp):
    """Keyframe = right image, frame = left image of the same scene; feature vectors from the bag-of-words oracle."""
    from oracle import oracle_bow_py as ob
    f = frames
    o, r = om.oracle(), om.reference()
    voc = synth.synth_vocabulary(61, kl[0], kl[1])
    ov = ob.OracleVocabulary(voc)
    fvF = ov.transform(f["dL"], levelsup)
    for seed in range(3):
        rng = np.random.default_rng(seed)
        # the keyframe's descriptors: the frame's own (shuffled, a few bit flips) so that close pairs exist, plus the right image's
        perm = rng.permutation(len(f["dL"]))[:800]
        dK = f["dL"][perm].copy()
        bits = np.unpackbits(dK, axis=1)
        flip = rng.random(bits.shape) < 0.02
        dK = np.concatenate([np.packbits(bits ^ flip.astype(np.uint8), axis=1), f["dR"][:400]])
        aK = np.concatenate([f["kL"]["angle"][perm] + rng.normal(0, 3, len(perm)).astype(np.float32), f["kR"]["angle"][:400]]).astype(np.float32)
        flags = (rng.random(len(dK)) < pmp).astype(np.uint8)
        fvK = ov.transform(dK, levelsup)
        no, mo = o.search_by_bow(dK, aK, flags, fvK, f["dL"], f["kL"]["angle"], fvF, ratio, ori)
        nr, mr = r.search_by_bow(dK, aK, flags, fvK, f["dL"], f["kL"]["angle"], fvF, ratio, ori)
        assert no == nr and np.array_equal(mo, mr)
        assert nr > 100
    # degenerate: empty keyframe / empty frame / no common node
    e = dict(fv_node=np.zeros(0, np.uint32), fv_off=np.zeros(1, np.int32), fv_feat=np.zeros(0, np.uint32))
    for args in ((dK[:0], aK[:0], flags[:0], e, f["dL"], f["kL"]["angle"], fvF), (dK, aK, flags, fvK, f["dL"][:0], f["kL"]["angle"][:0], e)):
        no, mo = o.search_by_bow(*args, ratio, ori)
        nr, mr = r.search_by_bow(*args, ratio, ori)
        assert no == nr == 0 and np.array_equal(mo, mr)
