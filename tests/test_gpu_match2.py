"""Two-camera (Nleft != -1) halves of the windowed matcher on the GPU (orb_search_by_projection_stereo,
orb_search_local_points_stereo; right grid = orb_assign_features_to_grid on the right handle) against the Python
restatement of the reference's loops (oracle/oracle_match2_py.py, equal to the reference's own code: tests/test_oracle_match2.py)
and, when the built reference library travelled to the box, against the reference itself. Bit-exact, TUM-VI-shape frames."""
import numpy as np
import pytest

from morb_slam_b200 import capi, synth
from oracle import oracle_match2_py as o2
from oracle import oracle_match_py as om
from tests.conftest import has_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")]
W, H, NF, LAP = synth.CONFIGS["tumvi"][:4]
TRL = (-14.25, 0.75)
B = 3


@pytest.fixture(scope="module")
def rig():
    pairs = [synth.stereo_pair(3000 + i, W, H) for i in range(B)]
    L = np.stack([p[0] for p in pairs]); R = np.stack([p[1] for p in pairs])
    exL = capi.ORBextractor(NF, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
    exR = capi.ORBextractor(NF, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
    nL, _, kL, dL = exL.extract_batch(L, LAP)
    nR, _, kR, dR = exR.extract_batch(R, LAP)
    gp = capi.grid_params(W, H)
    capi.assign_features_to_grid(exL, gp)
    capi.assign_features_to_grid(exR, gp)     # = mGridRight
    fr = [(kL[f, :nL[f]].copy(), dL[f, :nL[f]].copy(), kR[f, :nR[f]].copy(), dR[f, :nR[f]].copy()) for f in range(B)]
    return exL, exR, fr, exL.tables()["scale"], om.grid_params(W, H)


def _pad(rows, dtype, cap, tail=()):
    out = np.zeros((len(rows), cap) + tuple(tail), dtype)
    for f, r in enumerate(rows):
        out[f, :len(r)] = r
    return out


@pytest.mark.parametrize("th,mono,tlc,ori,jit,pobs", [(7.0, False, 0.0, True, 4.0, 0.8), (15.0, True, 0.0, True, 8.0, 0.8),
                                                       (7.0, False, 0.5, True, 4.0, 0.8), (7.0, False, -0.5, False, 4.0, 0.8),
                                                       (15.0, False, 0.0, False, 10.0, 0.3)])
def test_search_by_projection_stereo(rig, th, mono, tlc, ori, jit, pobs):
    exL, exR, fr, scale, gp = rig
    qs = [synth.synth_queries2(500 + f, *fr[f], W, H, TRL, p_obs=pobs, jitter=jit) for f in range(B)]
    qcap = max(len(q[1]) for q in qs) + 5
    Q = _pad([q[1] for q in qs], capi.Q2_DTYPE, qcap)
    QD = _pad([q[2] for q in qs], np.uint8, qcap, (32,))
    nq = np.array([len(q[1]) for q in qs], np.int32)
    nm, mL, mR = capi.search_by_projection_stereo(exL, exR, Q, QD, nq, th, mono, np.full(B, tlc, np.float32), 0.1, ori)
    for f in range(B):
        kL, dL, kR, dR = fr[f]
        nm_o, m_o = o2.search_by_projection2(kL, dL, kR, dR, scale, gp, 0.1, qs[f][1], qs[f][2], th, mono, tlc, ori)
        assert nm[f] == nm_o, f
        assert np.array_equal(mL[f, :len(kL)], m_o[:len(kL)]) and np.array_equal(mR[f, :len(kR)], m_o[len(kL):]), f
        assert (mL[f, len(kL):] == -1).all() and (mR[f, len(kR):] == -1).all()
        if o2.have_reference():
            nm_r, m_r = o2.ref_search_by_projection2(kL, dL, kR, dR, scale, gp, 0.1, TRL, qs[f][0], qs[f][2], th, mono, tlc, ori)
            assert nm[f] == nm_r and np.array_equal(np.concatenate([mL[f, :len(kL)], mR[f, :len(kR)]]), m_r), f


@pytest.mark.parametrize("th,ratio,jit,pobs,plock", [(3.0, 0.8, 3.0, 0.9, 0.3), (5.0, 0.8, 6.0, 0.9, 0.0), (1.0, 0.9, 3.0, 0.5, 0.6),
                                                      (15.0, 0.9, 10.0, 0.3, 0.85)])
def test_search_local_points_stereo(rig, th, ratio, jit, pobs, plock):
    exL, exR, fr, scale, gp = rig
    pr = [synth.synth_stereo_pairing(600 + f, len(fr[f][0]), len(fr[f][2])) for f in range(B)]
    qs = [synth.synth_track_queries2(700 + f, *fr[f], pr[f][0], W, H, p_obs=pobs, jitter=jit) for f in range(B)]
    lk = [(np.random.default_rng(800 + f).random(len(fr[f][0]) + len(fr[f][2])) < plock).astype(np.uint8) for f in range(B)]
    qcap = max(len(q[0]) for q in qs) + 3
    Q = _pad([q[0] for q in qs], capi.TQ2_DTYPE, qcap)
    QD = _pad([q[1] for q in qs], np.uint8, qcap, (32,))
    nq = np.array([len(q[0]) for q in qs], np.int32)
    lkL = _pad([lk[f][:len(fr[f][0])] for f in range(B)], np.uint8, exL.kcap)
    lkR = _pad([lk[f][len(fr[f][0]):] for f in range(B)], np.uint8, exR.kcap)
    l2r = np.full((B, exL.kcap), -1, np.int32); r2l = np.full((B, exR.kcap), -1, np.int32)
    for f in range(B):
        l2r[f, :len(pr[f][0])] = pr[f][0]; r2l[f, :len(pr[f][1])] = pr[f][1]
    nm, mL, mR = capi.search_local_points_stereo(exL, exR, Q, QD, nq, lkL, lkR, l2r, r2l, th, ratio)
    for f in range(B):
        kL, dL, kR, dR = fr[f]
        nm_o, m_o = o2.search_local_points2(kL, dL, kR, dR, lk[f], pr[f][0], pr[f][1], scale, gp, qs[f][0], qs[f][1], th, ratio)
        assert nm[f] == nm_o, f
        assert np.array_equal(mL[f, :len(kL)], m_o[:len(kL)]) and np.array_equal(mR[f, :len(kR)], m_o[len(kL):]), f
        if o2.have_reference():
            nm_r, m_r = o2.ref_search_local_points2(kL, dL, kR, dR, lk[f], pr[f][0], pr[f][1], scale, gp, qs[f][0], qs[f][1], th, ratio)
            assert nm[f] == nm_r and np.array_equal(np.concatenate([mL[f, :len(kL)], mR[f, :len(kR)]]), m_r), f


def test_local_points_stereo_uses_resident_pairing(rig):
    """left_to_right / right_to_left = None: the pairing left on the device by the fisheye triangulation (mvLeftToRightMatch /
    mvRightToLeftMatch of ComputeStereoFishEyeMatches) - config 3 end to end on the device"""
    exL, exR, fr, scale, gp = rig
    capi.compute_stereo_fisheye_matches_batch(exL, exR)
    tri = capi.compute_stereo_fisheye_triangulation_batch(exL, exR, synth.kb8_rig("tumvi"))
    l2r_all, r2l_all = tri[0], tri[1]
    qs = [synth.synth_track_queries2(900 + f, *fr[f], l2r_all[f, :len(fr[f][0])], W, H) for f in range(B)]
    qcap = max(len(q[0]) for q in qs)
    Q = _pad([q[0] for q in qs], capi.TQ2_DTYPE, qcap)
    QD = _pad([q[1] for q in qs], np.uint8, qcap, (32,))
    nq = np.array([len(q[0]) for q in qs], np.int32)
    nm, mL, mR = capi.search_local_points_stereo(exL, exR, Q, QD, nq, None, None, None, None, 3.0, 0.8)
    for f in range(B):
        kL, dL, kR, dR = fr[f]
        lk = np.zeros(len(kL) + len(kR), np.uint8)
        nm_o, m_o = o2.search_local_points2(kL, dL, kR, dR, lk, l2r_all[f, :len(kL)], r2l_all[f, :len(kR)], scale, gp, qs[f][0], qs[f][1], 3.0, 0.8)
        assert nm[f] == nm_o and np.array_equal(mL[f, :len(kL)], m_o[:len(kL)]) and np.array_equal(mR[f, :len(kR)], m_o[len(kL):]), f
    assert (l2r_all >= 0).sum() > 100


def test_two_camera_argument_errors(rig):
    exL, exR, fr, scale, gp = rig
    Q = np.zeros((B, 4), capi.Q2_DTYPE); QD = np.zeros((B, 4, 32), np.uint8); nq = np.zeros(B, np.int32)
    with pytest.raises(capi.OrbError):
        capi.search_by_projection_stereo(exL, exL, Q, QD, nq, 7.0, False, np.zeros(B, np.float32), 0.1)      # one handle twice
    ex3 = capi.ORBextractor(NF, 1.2, 8, 20, 7, max_width=W, max_height=H)
    ex3(synth.mono_frame(1, W, H), LAP)
    with pytest.raises(capi.OrbError):
        capi.search_by_projection_stereo(exL, ex3, Q, QD, nq, 7.0, False, np.zeros(B, np.float32), 0.1)      # no grid / other batch


@pytest.mark.parametrize("kl,levelsup,ratio,ori,pmp", [((10, 4), 2, 0.7, True, 0.8), ((10, 4), 3, 0.75, True, 1.0), ((6, 3), 3, 0.9, False, 0.5),
                                                        ((10, 4), 0, 0.7, True, 0.8)])
def test_bow_stereo(rig, kl, levelsup, ratio, ori, pmp):
    """orb_compute_bow_stereo = ComputeBoW on vconcat(left, right) descriptors: BowVector doubles and FeatureVector equal to the
    DBoW2 restatement on the concatenation (bit for bit); orb_search_by_bow_stereo = SearchByBoW with F.Nleft != -1."""
    from oracle import oracle_bow_py as ob
    exL, exR, fr, scale, gp = rig
    voc = synth.synth_vocabulary(71, kl[0], kl[1])
    gv = capi.ORBVocabulary(voc)
    ov = ob.OracleVocabulary(voc)
    got = capi.compute_bow_stereo(exL, exR, gv, levelsup)
    kfs = []
    for f in range(B):
        kL, dL, kR, dR = fr[f]
        want = ov.transform(np.concatenate([dL, dR]), levelsup)
        for k in ("bow_word", "fv_node", "fv_off", "fv_feat"):
            assert np.array_equal(got[f][k], want[k]), (f, k)
        assert got[f]["bow_val"].tobytes() == np.asarray(want["bow_val"], np.float64).tobytes(), f
        dK, aK, fl = synth.synth_bow_keyframe(50 + f, kL, dL, kR, dR, pmp)
        if f == 1:
            dK, aK, fl = dK[:0], aK[:0], fl[:0]                      # one frame gets an empty keyframe
        kfs.append(dict(desc=dK, angle=aK, flags=fl, fv=ov.transform(dK, levelsup)))
    nm, mL, mR = capi.search_by_bow_stereo(exL, exR, kfs, ratio, ori)
    for f in range(B):
        kL, dL, kR, dR = fr[f]
        k = kfs[f]
        nm_o, m_o = o2.search_by_bow2(k["desc"], k["angle"], k["flags"], k["fv"], dL, kL["angle"], dR, kR["angle"], got[f], ratio, ori)
        assert nm[f] == nm_o and np.array_equal(mL[f, :len(kL)], m_o[:len(kL)]) and np.array_equal(mR[f, :len(kR)], m_o[len(kL):]), f
        if o2.have_reference() and len(k["desc"]):
            nm_r, m_r = o2.ref_search_by_bow2(k["desc"], k["angle"], k["flags"], k["fv"], dL, kL["angle"], dR, kR["angle"], got[f], ratio, ori)
            assert nm[f] == nm_r and np.array_equal(np.concatenate([mL[f, :len(kL)], mR[f, :len(kR)]]), m_r), f
    assert (mR >= 0).sum() > 30 and (mL >= 0).sum() > 30
