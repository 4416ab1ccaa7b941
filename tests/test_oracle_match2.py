"""Two-camera (Nleft != -1) halves of the windowed matcher: the Python restatement (oracle/oracle_match2_py.py) against the
reference's own code compiled by line range (oracle/_ref/libmorb_ref_match.so), on TUM-VI-shape frames. CPU only."""
import numpy as np
import pytest

from morb_slam_b200 import synth
from oracle import oracle_match2_py as o2
from oracle import oracle_match_py as om
from oracle import oracle_py as op

pytestmark = pytest.mark.skipif(not o2.have_reference(), reason="oracle/_ref was not built (needs the reference mount)")
W, H, NF, LAP = synth.CONFIGS["tumvi"][:4]


@pytest.fixture(scope="module")
def rig():
    L, R = synth.stereo_pair(3000, W, H)
    eL, eR = op.OracleExtractor(NF), op.OracleExtractor(NF)
    _, kL, dL = eL(L, LAP)
    _, kR, dR = eR(R, LAP)
    return kL, dL, kR, dR, eL.tables()["scale"], om.grid_params(W, H)


def test_features_in_area_right_grid(rig):
    kL, dL, kR, dR, scale, gp = rig
    rng = np.random.default_rng(5)
    gl, gr = o2.Grid(kL, gp), o2.Grid(kR, gp)
    for t in range(300):
        x, y = rng.uniform(-30, W + 30), rng.uniform(-30, H + 30)
        r = float(rng.choice([3.0, 7.5, 15.0, 40.0, 120.0]))
        lo = int(rng.integers(-1, 6)); hi = int(rng.choice([-1, lo, lo + 1, lo + 2]))
        right = bool(t % 2)
        want = o2.ref_features_in_area2(kL, kR, gp, x, y, r, lo, hi, right)
        got = (gr if right else gl).features_in_area(x, y, r, lo, hi)
        assert got == want, (t, x, y, r, lo, hi, right)


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("th,mono,tlc,ori,jit,pobs", [(7.0, False, 0.0, True, 4.0, 0.8), (15.0, True, 0.0, True, 8.0, 0.8),
                                                       (7.0, False, 0.5, True, 4.0, 0.8), (7.0, False, -0.5, False, 4.0, 0.8),
                                                       (15.0, False, 0.0, False, 10.0, 0.3)])
def test_search_by_projection2_restatement_equals_reference(rig, seed, th, mono, tlc, ori, jit, pobs):
    kL, dL, kR, dR, scale, gp = rig
    trl = (-14.25, 0.75)
    q, q2, qd = synth.synth_queries2(100 + seed, kL, dL, kR, dR, W, H, trl, p_obs=pobs, jitter=jit)
    nm_r, m_r = o2.ref_search_by_projection2(kL, dL, kR, dR, scale, gp, 0.1, trl, q, qd, th, mono, tlc, ori)
    nm_o, m_o = o2.search_by_projection2(kL, dL, kR, dR, scale, gp, 0.1, q2, qd, th, mono, tlc, ori)
    assert nm_o == nm_r and np.array_equal(m_o, m_r)
    assert (m_r[len(kL):] >= 0).sum() > 50 and (m_r[:len(kL)] >= 0).sum() > 50


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("th,ratio,jit,pobs,plock", [(3.0, 0.8, 3.0, 0.9, 0.3), (5.0, 0.8, 6.0, 0.9, 0.0), (1.0, 0.9, 3.0, 0.5, 0.6),
                                                      (15.0, 0.9, 10.0, 0.3, 0.5)])
def test_search_local_points2_restatement_equals_reference(rig, seed, th, ratio, jit, pobs, plock):
    kL, dL, kR, dR, scale, gp = rig
    l2r, r2l = synth.synth_stereo_pairing(200 + seed, len(kL), len(kR))
    q, qd = synth.synth_track_queries2(300 + seed, kL, dL, kR, dR, l2r, W, H, p_obs=pobs, jitter=jit)
    locked0 = (np.random.default_rng(400 + seed).random(len(kL) + len(kR)) < plock).astype(np.uint8)
    nm_r, m_r = o2.ref_search_local_points2(kL, dL, kR, dR, locked0, l2r, r2l, scale, gp, q, qd, th, ratio)
    nm_o, m_o = o2.search_local_points2(kL, dL, kR, dR, locked0, l2r, r2l, scale, gp, q, qd, th, ratio)
    assert nm_o == nm_r and np.array_equal(m_o, m_r)
    assert (m_r[len(kL):] >= 0).sum() > 30 and (m_r[:len(kL)] >= 0).sum() > 30


def test_degenerate_two_camera_frames(rig):
    kL, dL, kR, dR, scale, gp = rig
    trl = (-14.25, 0.75)
    for a, b in (((kL, dL), (kR[:0], dR[:0])), ((kL[:0], dL[:0]), (kR, dR)), ((kL[:0], dL[:0]), (kR[:0], dR[:0])), ((kL[:3], dL[:3]), (kR[:2], dR[:2]))):
        q, q2, qd = synth.synth_queries2(7, kL, dL, kR, dR, W, H, trl)
        nm_r, m_r = o2.ref_search_by_projection2(a[0], a[1], b[0], b[1], scale, gp, 0.1, trl, q, qd, 7.0)
        nm_o, m_o = o2.search_by_projection2(a[0], a[1], b[0], b[1], scale, gp, 0.1, q2, qd, 7.0)
        assert nm_o == nm_r and np.array_equal(m_o, m_r)
        l2r, r2l = synth.synth_stereo_pairing(8, len(a[0]), len(b[0]))
        tq, tqd = synth.synth_track_queries2(9, kL, dL, kR, dR, np.full(len(kL), -1, np.int32), W, H)
        lk = np.zeros(len(a[0]) + len(b[0]), np.uint8)
        nm_r, m_r = o2.ref_search_local_points2(a[0], a[1], b[0], b[1], lk, l2r, r2l, scale, gp, tq, tqd, 3.0)
        nm_o, m_o = o2.search_local_points2(a[0], a[1], b[0], b[1], lk, l2r, r2l, scale, gp, tq, tqd, 3.0)
        assert nm_o == nm_r and np.array_equal(m_o, m_r)


@pytest.mark.parametrize("kl,levelsup,ratio,ori,pmp", [((10, 4), 2, 0.7, True, 0.8), ((10, 4), 3, 0.75, True, 1.0), ((6, 3), 3, 0.9, False, 0.5),
                                                        ((10, 4), 0, 0.7, True, 0.8)])
def test_search_by_bow2_restatement_equals_reference(rig, kl, levelsup, ratio, ori, pmp):
    from oracle import oracle_bow_py as ob
    kL, dL, kR, dR, scale, gp = rig
    voc = synth.synth_vocabulary(71, kl[0], kl[1])
    ov = ob.OracleVocabulary(voc)
    fvF = ov.transform(np.concatenate([dL, dR]), levelsup)          # ComputeBoW on vconcat(left, right)
    dK, aK, fl = synth.synth_bow_keyframe(50, kL, dL, kR, dR, pmp)
    fvK = ov.transform(dK, levelsup)
    nm_r, m_r = o2.ref_search_by_bow2(dK, aK, fl, fvK, dL, kL["angle"], dR, kR["angle"], fvF, ratio, ori)
    nm_o, m_o = o2.search_by_bow2(dK, aK, fl, fvK, dL, kL["angle"], dR, kR["angle"], fvF, ratio, ori)
    assert nm_o == nm_r and np.array_equal(m_o, m_r)
    assert (m_r[:len(kL)] >= 0).sum() > 20 and (m_r[len(kL):] >= 0).sum() > 20
