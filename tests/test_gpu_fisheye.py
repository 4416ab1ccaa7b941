"""Fisheye stereo triangulation on the GPU (include/orb_b200.h: orb_kb8_triangulate_matches, orb_stereo_fisheye_triangulate_batch =
KannalaBrandt8::TriangulateMatches + the acceptance loop of Frame::ComputeStereoFishEyeMatches, src/Frame.cc:1244-1273) against the
CPU oracle (oracle/orb_oracle_kb8.cc; equal to the reference's own lines on the Eigen stand-in, tests/test_oracle_kb8.py).
Floating-point row, bit-exact since round 2: the kernel runs glibc's tanf / atan2f / sinf / cosf restated for the device (pinned
exhaustively against the image's libm) and Eigen's two-sided float Jacobi SVD restated from its published algorithm, the same
restatement the oracle's Eigen stand-in runs - accept / reject codes, depths and 3-D points must be EQUAL, on 3 rigs x 3 x 20 000
pairs and on whole batches. (north_star's tolerance for this row would be 1e-3 px; it is not needed.)"""
import numpy as np
import pytest

from morb_slam_b200 import capi, synth
from oracle import oracle_kb8_py as ok

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4


def _codes(ret):
    return np.where(ret > 0, 1, ret).astype(np.int64)


@pytest.mark.parametrize("kind", ["tumvi", "parallel", "toed"])
def test_triangulate_matches_pairs(kind):
    ex = capi.ORBextractor(1500, 1.2, 8, 20, 7, max_width=512, max_height=512)
    o = ok.oracle()
    rig = synth.kb8_rig(kind)
    for seed in range(3):
        xy1, xy2, s1, s2 = synth.kb8_pairs(700 + seed, rig, 20000)
        ret, p3d = capi.kb8_triangulate_matches(ex, rig, xy1, xy2, s1, s2)
        ro, po, q = o.triangulate(rig, xy1, xy2, s1, s2)
        cd, co = _codes(ret), _codes(ro)
        assert np.array_equal(cd, co), np.nonzero(cd != co)[0][:5]          # every accept / reject code
        assert ret.tobytes() == ro.tobytes() and p3d.tobytes() == po.tobytes()   # depths and 3-D points, bit for bit
        assert (cd == 1).sum() > 5000 and np.all(p3d[cd != 1] == 0)
        assert len(set(cd.tolist()) & {-1, -2, -3, -4, -5}) >= 4             # the exits of TriangulateMatches are reached (all five over the rigs)
    r0, p0 = capi.kb8_triangulate_matches(ex, rig, xy1[:0], xy2[:0], s1[:0], s2[:0])
    assert len(r0) == 0


@pytest.mark.parametrize("kind,lap", [("parallel", (0, 511)), ("tumvi", (0, 511)), ("parallel", (150, 400)), ("parallel", (600, 700))])
def test_fisheye_stereo_triangulation_batch(kind, lap):
    """the whole Frame::ComputeStereoFishEyeMatches on a TUM-VI-shape batch: extraction with a lapping area, knnMatch + ratio on the
    device, then the triangulation of the passing matches. synth.stereo_pair shifts the scene horizontally (2 .. 60 px), which is what
    the "parallel" rig sees; under the TUM-VI calibration (4.7 % tilt between the cameras) most of these matches fail the gates."""
    w, h, nf = synth.CONFIGS["tumvi"][:3]
    B = 4
    Ls = np.stack([synth.stereo_pair(7300 + i, w, h)[0] for i in range(B)])
    Rs = np.stack([synth.stereo_pair(7300 + i, w, h)[1] for i in range(B)])
    exL = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B)
    exR = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B)
    rig = synth.kb8_rig(kind)
    nL, mL, kL, dL = exL.extract_batch(Ls, lap)
    nR, mR, kR, dR = exR.extract_batch(Rs, lap)
    with pytest.raises(capi.OrbError):       # call order: the kNN has not run on this batch
        capi.compute_stereo_fisheye_triangulation_batch(exL, exR, rig)
    idx, dist, passed = capi.compute_stereo_fisheye_matches_batch(exL, exR)
    l2r, r2l, depth, p3d, code = capi.compute_stereo_fisheye_triangulation_batch(exL, exR, rig)
    o = ok.oracle()
    sigma2 = exL.tables()["sigma2"]
    n_acc = 0
    for f in range(B):
        nq = nL[f] - mL[f]
        lo, ro_, do, po, co, q = o.fisheye_accept(rig, kL[f, :nL[f]], mL[f], kR[f, :nR[f]], mR[f], sigma2, idx[f, :nq], dist[f, :nq])
        cd = code[f, :nL[f]]
        assert np.array_equal(cd, co), (f, np.nonzero(cd != co)[0][:5])
        assert np.array_equal(l2r[f, :nL[f]], lo) and np.array_equal(r2l[f, :nR[f]], ro_)
        assert depth[f, :nL[f]].tobytes() == do.tobytes() and p3d[f, :nL[f]].tobytes() == po.tobytes()
        acc = co == 1
        n_acc += int(acc.sum())
        rej = ~acc
        assert np.all(depth[f, :nL[f]][rej] == -1) and np.all(p3d[f, :nL[f]][rej] == 0) and np.all(l2r[f, :nL[f]][rej] == -1)
        # the ratio test gates everything
        assert np.all(cd[mL[f]:][passed[f, :nq] == 0] == 0) and np.all(cd[:mL[f]] == 0)
        # beyond the frame's keypoints: defaults
        assert np.all(l2r[f, nL[f]:] == -1) and np.all(depth[f, nL[f]:] == -1) and np.all(code[f, nL[f]:] == 0)
    if kind == "parallel" and lap == (0, 511):
        assert n_acc > 200
    if lap == (600, 700):
        assert n_acc == 0 and np.all(code == 0)


def test_full_batch_properties():
    """BASELINE.json configs[2] at batch size (64 frames per launch here, 4 distinct pairs tiled): size-independent properties -
    a frame's result does not depend on its position in the batch (bit-identical for the repeated frames), the inverse map holds the
    largest left index of every matched right keypoint, accepted matches have positive depth = z of the 3-D point."""
    w, h, nf, lap = synth.CONFIGS["tumvi"][:4]
    B, D = 64, 4
    pairs = [synth.stereo_pair(7400 + i, w, h) for i in range(D)]
    Ls = np.stack([pairs[i % D][0] for i in range(B)]); Rs = np.stack([pairs[i % D][1] for i in range(B)])
    exL = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B)
    exR = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B)
    nL, mL, kL, dL = exL.extract_batch(Ls, lap)
    nR, mR, kR, dR = exR.extract_batch(Rs, lap)
    capi.compute_stereo_fisheye_matches_batch(exL, exR, want=False)
    l2r, r2l, depth, p3d, code = capi.compute_stereo_fisheye_triangulation_batch(exL, exR, synth.kb8_rig("parallel"))
    for f in range(D, B):
        g = f % D
        assert np.array_equal(l2r[f], l2r[g]) and np.array_equal(r2l[f], r2l[g]) and np.array_equal(code[f], code[g])
        assert depth[f].tobytes() == depth[g].tobytes() and p3d[f].tobytes() == p3d[g].tobytes()
    total = 0
    for f in range(D):
        acc = l2r[f] >= 0
        total += int(acc.sum())
        assert np.array_equal(acc, code[f] == 1) and np.all(depth[f][acc] > 1e-4) and np.array_equal(depth[f][acc], p3d[f][acc][:, 2])
        for j in np.unique(l2r[f][acc]):
            assert r2l[f, j] == np.nonzero(l2r[f] == j)[0].max()
        assert np.all(r2l[f][np.setdiff1d(np.arange(r2l.shape[1]), l2r[f][acc])] == -1)
    assert total > 400


def test_small_batch_knn_equals_large_batch_knn():
    """batches of at most 8 frames split the right keypoints into chunks over blockIdx.z and merge the chunks' top-2 lists; larger
    batches scan all rows in one block: the same pairs give the same kNN lists, ratio flags and triangulation either way"""
    w, h, nf, lap = synth.CONFIGS["tumvi"][:4]
    D, B = 3, 12
    pairs = [synth.stereo_pair(7450 + i, w, h) for i in range(D)]
    outs = []
    for n in (D, B, 1):
        Ls = np.stack([pairs[i % D][0] for i in range(n)]); Rs = np.stack([pairs[i % D][1] for i in range(n)])
        exL = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=n)
        exR = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=n)
        exL.extract_batch(Ls, lap); exR.extract_batch(Rs, lap)
        idx, dist, passed = capi.compute_stereo_fisheye_matches_batch(exL, exR)
        tri = capi.compute_stereo_fisheye_triangulation_batch(exL, exR, synth.kb8_rig("tumvi"))
        if n <= 8:   # page-locked result buffers (written by the kernel itself for a few frames): the same values
            tri_p = capi.compute_stereo_fisheye_triangulation_batch(exL, exR, synth.kb8_rig("tumvi"), pinned=True)
            for a, b in zip(tri, tri_p):
                assert a.tobytes() == b.tobytes()
        outs.append((idx, dist, passed) + tuple(tri))
    small, large, single = outs
    assert (small[2] != 0).sum() > 100
    for f in range(B):
        for a, b in zip(small, large):
            assert a[f % D].tobytes() == b[f].tobytes(), f
    for a, b in zip(small, single):
        assert a[0].tobytes() == b[0].tobytes()
