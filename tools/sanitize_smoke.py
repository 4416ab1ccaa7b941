"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck / initcheck runs; the
logs are committed under profiles/): two extractions on two handles (both quad-tree kernels via ORB_B200_OCTREE), the stereo
matcher, a batch of 3 with the captured pipeline, a kNN scan + sharded exchange between two handles of this process, the
fisheye matcher + triangulation, the windowed matcher, the bag of words and the LocalMapping-side matchers on loaded keyframes. Results are checked against the oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from morb_slam_b200 import capi, synth  # noqa: E402
from oracle import oracle_py as op  # noqa: E402


def main():
    w, h, nf, lap, fx, b = synth.CONFIGS["euroc"]
    L, R = synth.stereo_pair(2000, w, h)
    exL = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=3)
    exR = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=3)
    mL, kL, dL = exL(L, lap)
    mR, kR, dR = exR(R, lap)
    mbf, maxD = float(np.float32(fx * b)), float(np.float32(fx))
    uR, dp = capi.compute_stereo_matches(exL, exR, kL, dL, kR, dR, mbf, maxD)
    oL, oR = op.OracleExtractor(nf), op.OracleExtractor(nf)
    _, koL, doL = oL(L, lap)
    _, koR, doR = oR(R, lap)
    uo, do = op.oracle_stereo(oL, oR, koL, doL, koR, doR, mbf, maxD)
    assert kL.tobytes() == koL.tobytes() and np.array_equal(dL, doL) and uR.tobytes() == uo.tobytes() and dp.tobytes() == do.tobytes()
    # batch of 3, three times: plain launches, capture, replay
    imgs = np.stack([synth.mono_frame(4300 + i, w, h) for i in range(3)])
    for _ in range(3):
        n, mono, kps, desc = exL.extract_batch(imgs, lap)
    o = op.OracleExtractor(nf)
    _, ko, do_ = o(imgs[2], lap)
    assert kps[2, :n[2]].tobytes() == ko.tobytes() and np.array_equal(desc[2, :n[2]], do_)
    # kNN scan, and the sharded exchange between two handles of this process
    q = synth.random_descriptors(0, 300); db = synth.clustered_descriptors(2, q, 20000)
    idx, dist = capi.hamming_knn2(exL, q, db)
    io, dio = op.oracle_knn2(q, db)
    assert np.array_equal(idx, io) and np.array_equal(dist, dio)
    xs = [capi.KnnExchange(ex, r, 2, 512) for r, ex in enumerate((exL, exR))]
    for x in xs:
        x.connect_local(xs)
    tq = torch.from_numpy(q).cuda(); tdb = torch.from_numpy(db).cuda()
    outs = [torch.empty((2, 300, 2), dtype=torch.int32, device="cuda") for _ in range(2)]
    torch.cuda.synchronize()
    for r, x in enumerate(xs):
        x.search(tq.data_ptr(), 300, tdb[r * 10000:(r + 1) * 10000].data_ptr(), 10000, r * 10000, outs[r][0].data_ptr(), outs[r][1].data_ptr(),
                 flags=capi.ORB_ASYNC)
    exL.sync(); exR.sync(); torch.cuda.synchronize()
    assert np.array_equal(outs[0][0].cpu().numpy(), io) and np.array_equal(outs[1][1].cpu().numpy(), dio)
    for x in xs:
        x.close()
    # fisheye pair: kNN + ratio + triangulation
    wf, hf, nff, lapf, _, _ = synth.CONFIGS["tumvi"]
    Lf, Rf = synth.stereo_pair(3000, wf, hf)
    fL = capi.ORBextractor(nff, 1.2, 8, 20, 7, max_width=wf, max_height=hf)
    fR = capi.ORBextractor(nff, 1.2, 8, 20, 7, max_width=wf, max_height=hf)
    fL.extract_batch(Lf[None], lapf); fR.extract_batch(Rf[None], lapf)
    capi.compute_stereo_fisheye_matches_batch(fL, fR)
    capi.compute_stereo_fisheye_triangulation_batch(fL, fR, synth.kb8_rig("tumvi"))
    # two-camera searches on the fisheye pair (right grid, shared histogram, partner assignments, combined bag of words)
    from oracle import oracle_match_py as om
    gpf = capi.grid_params(wf, hf)
    capi.assign_features_to_grid(fL, gpf); capi.assign_features_to_grid(fR, gpf)
    nl, _, kfl, dfl = fL.extract_batch(Lf[None], lapf, flags=0)
    nr, _, kfr, dfr = fR.extract_batch(Rf[None], lapf, flags=0)
    capi.assign_features_to_grid(fL, gpf); capi.assign_features_to_grid(fR, gpf)
    a, b_, c, d = kfl[0, :nl[0]], dfl[0, :nl[0]], kfr[0, :nr[0]], dfr[0, :nr[0]]
    _, q2, qd2 = synth.synth_queries2(500, a, b_, c, d, wf, hf)
    capi.search_by_projection_stereo(fL, fR, q2[None], qd2[None], np.array([len(q2)], np.int32), 7.0, False, np.zeros(1, np.float32), 0.1)
    l2r, r2l = synth.synth_stereo_pairing(600, len(a), len(c))
    tq, tqd = synth.synth_track_queries2(700, a, b_, c, d, l2r, wf, hf)
    L2R = np.full((1, fL.kcap), -1, np.int32); R2L = np.full((1, fR.kcap), -1, np.int32)
    L2R[0, :len(a)] = l2r; R2L[0, :len(c)] = r2l
    capi.search_local_points_stereo(fL, fR, tq[None], tqd[None], np.array([len(tq)], np.int32), None, None, L2R, R2L, 3.0)
    vocf = synth.synth_vocabulary(71, 10, 4)
    gvf = capi.ORBVocabulary(vocf)
    fv2 = capi.compute_bow_stereo(fL, fR, gvf, 2)
    dK, aK, fl = synth.synth_bow_keyframe(50, a, b_, c, d)
    from oracle import oracle_bow_py as ob
    capi.search_by_bow_stereo(fL, fR, [dict(desc=dK, angle=aK, flags=fl, fv=ob.OracleVocabulary(vocf).transform(dK, 2))])
    # windowed matcher + bag of words on the device-resident left frame
    gp = capi.grid_params(w, h)
    exL.extract_batch(L[None], lap)
    capi.assign_features_to_grid(exL, gp)
    qq, qd = synth.synth_queries(5, kR, dR, None, None, w, h)
    Q = np.zeros((1, len(qq)), capi.Q_DTYPE); Q[0] = qq
    capi.search_by_projection(exL, Q, qd[None], np.array([len(qq)], np.int32), 7.0, False, np.zeros(1, np.float32), float(np.float32(b)), mbf)
    voc = synth.synth_vocabulary(41, 10, 4, 0.0, 0.0)
    gv = capi.ORBVocabulary(voc)
    capi.compute_bow(exL, gv, 2)
    tq1, tqd1 = synth.synth_track_queries(78, kL, dL, uR, w, h)
    TQ = np.zeros((1, len(tq1)), capi.TQ_DTYPE); TQ[0] = tq1
    capi.search_local_points(exL, TQ, tqd1[None], np.array([len(tq1)], np.int32), None, 3.0)
    # LocalMapping-side matchers: keyframes loaded into the handle, Fuse search, SearchByProjection(Frame, KeyFrame), SearchForTriangulation,
    # ComputeDistinctiveDescriptors
    from oracle import oracle_map_py as omap
    capi.load_frames(exL, [kL, kR], [dL, dR], [uR, np.full(len(kR), -1, np.float32)])
    capi.assign_features_to_grid(exL, gp)
    pts, pdesc, _, _ = synth.synth_fuse_points(11, kL, dL, w, h)
    fq = omap.fuse_queries(pts, mbf)
    FQ = np.zeros((2, len(fq)), capi.FQ_DTYPE); FQ[0] = fq; FQ[1] = fq
    bi, bd = capi.fuse_search(exL, FQ, np.stack([pdesc, pdesc]), np.array([len(fq), len(fq) // 2], np.int32), 3.0, 0)
    obi, obd = omap.fuse_search(kL, dL, uR, oL.tables()["scale"], oL.tables()["inv_sigma2"], gp, fq, pdesc, 3.0, 0)
    assert np.array_equal(bi[0], obi) and np.array_equal(bd[0], obd)
    qq2, qd3 = synth.synth_queries(6, kL, dL, None, None, w, h, jitter=2.0)
    qq2["flags"] &= 1
    KQ = np.zeros((2, len(qq2)), capi.Q_DTYPE); KQ[0] = qq2; KQ[1] = qq2
    nmk, mk = capi.search_by_projection_kf(exL, KQ, np.stack([qd3, qd3]), np.array([len(qq2), 0], np.int32), None, 10.0, 100, True)
    onm, omk = omap.search_by_projection_kf(kL, dL, None, oL.tables()["scale"], gp, qq2, qd3, 10.0, 100, True)
    assert nmk[0] == onm and np.array_equal(mk[0, :len(kL)], omk) and nmk[1] == 0
    k1, k2 = synth.synth_triangulation_pair(12, kL, dL, uR, w, h)
    nmt, m12 = capi.search_for_triangulation(exL, [k1, k2], [(0, 1), (1, 0)], np.stack([synth.synth_fundamental(12)] * 2),
                                             np.array([[5000.0, 240.0], [300.0, 200.0]], np.float32))
    onmt, om12 = omap.search_for_triangulation(k1, k2, oL.tables()["scale"], oL.tables()["sigma2"], synth.synth_fundamental(12), (5000.0, 240.0))
    assert nmt[0] == onmt and np.array_equal(m12[0, :len(kL)], om12)
    # SearchForInitialization: F2 = the resident right-image keypoints, F1 = the left ones with near-duplicates (take-overs, ties);
    # window 100 (the lists) and 400 (the on-the-fly path)
    i1k, i1d, iprev, i2k, i2d = synth.synth_init_frames(13, kL, dL, kR, dR)
    capi.load_frames(exL, [i2k], [i2d])
    capi.assign_features_to_grid(exL, gp)
    IQ = np.zeros((1, len(i1k)), capi.IQ_DTYPE)
    IQ[0]["x"] = iprev[:, 0]; IQ[0]["y"] = iprev[:, 1]; IQ[0]["angle"] = i1k["angle"]; IQ[0]["octave"] = i1k["octave"]
    for win in (100, 400):
        inm, im12, ipv = capi.search_for_initialization(exL, IQ, i1d[None], np.array([len(i1k)], np.int32), win, 0.9, True)
        onmi, om12i, opvi = omap.search_for_initialization(i1k, i1d, iprev, i2k, i2d, gp, win, 0.9, True)
        assert inm[0] == onmi and np.array_equal(im12[0], om12i) and ipv[0].tobytes() == opvi.tobytes()
    # SearchForTriangulation between two-camera keyframes (KannalaBrandt8::epipolarConstrain per candidate)
    f1, f2, frigs = synth.synth_two_camera_keyframes(31, npts=300)
    fnm, fm12 = capi.search_for_triangulation_fisheye(exL, [f1, f2], [(0, 1)], frigs[None])
    ofnm, ofm12 = omap.search_for_triangulation_fisheye(f1, f2, oL.tables()["sigma2"], frigs)
    assert fnm[0] == ofnm and np.array_equal(fm12[0, :len(f1["kps"])], ofm12)
    obs = synth.synth_observations(4, 40)
    bb, mm = capi.distinctive_descriptors(exL, obs)
    assert all(bb[p] == omap.distinctive(d)[0] for p, d in enumerate(obs))
    print("sanitize_smoke ok: K = %d / %d, %d stereo matches, octree kernel %s" % (len(kL), len(kR), int((uR >= 0).sum()), os.environ.get("ORB_B200_OCTREE", "passes")))


if __name__ == "__main__":
    main()
