#!/usr/bin/env python3
"""Generate the committed golden vectors under tests/golden/ from the REFERENCE ITSELF
(oracle/_ref: /root/reference/src/ORBextractor.cc, src/Frame.cc:889-1047, src/ORBmatcher.cc:1880-1894
compiled unmodified; windowed matcher: src/Frame.cc:501-528,742-820 and src/ORBmatcher.cc:1521-1733,1844-1876) on the deterministic synthetic frames of morb_slam_b200/synth.py, and the kNN
vector from cv2.BFMatcher. Run in the build container (needs /root/reference and cv2):

    python tools/gen_golden.py
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from morb_slam_b200 import synth  # noqa: E402
from oracle import oracle_py as op  # noqa: E402
from oracle import oracle_match_py as om  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
CASES = [("euroc_mono", 1000), ("euroc", 2000), ("tumvi", 3000), ("kitti", 4000)]
STEREO = [("euroc", 2000), ("kitti", 4000)]
# name, th, bMono, tlc_z, check orientation, query jitter (px), share of map points with observations
MATCH = [("stereo_th7", 7.0, False, 0.0, True, 4.0, 0.8), ("mono_th15", 15.0, True, 0.0, True, 8.0, 0.8),
         ("forward_th7", 7.0, False, 0.5, True, 4.0, 0.8), ("unlocked_th15", 15.0, False, 0.0, False, 10.0, 0.3)]


# local-map search: name, th, nnratio, jitter, share of observed points, share of keypoints locked before the call
LOCAL = [("th3", 3.0, 0.8, 3.0, 0.9, 0.3), ("th5_wide", 5.0, 0.8, 6.0, 0.9, 0.0), ("th15_unobserved", 15.0, 0.9, 10.0, 0.3, 0.5)]
# bag of words: name, vocabulary seed, k, L, p_early_leaf, p_short, scoring, weighting, levelsup
BOW = [("orbvoc_like", 41, 10, 4, 0.0, 0.0, 0, 0, 2), ("ragged", 42, 8, 4, 0.05, 0.15, 0, 0, 2), ("l2_idf", 43, 5, 5, 0.0, 0.0, 1, 2, 3)]


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


SWEEP = [("euroc_mono", 1000), ("euroc", 2000), ("tumvi", 3000), ("kitti", 4000)]   # seeds base + 0..7 per configuration


def seed_sweep():
    """seeds 0..7 x the four image configurations through the REFERENCE (oracle/_ref), as CRC-32 records - equal
    CRCs mean bit-equal keypoints, descriptors, pyramid levels and stereo results. mono frames for every configuration, stereo
    pairs (extraction of both images + ComputeStereoMatches) for the two rectified-stereo configurations."""
    rec = {}
    for cfg, base in SWEEP:
        w, h, nf, lap, fx, b = synth.CONFIGS[cfg]
        rows = []
        for i in range(8):
            img = synth.mono_frame(base + i, w, h)
            r = op.RefExtractor(nf)
            mono, kps, desc = r(img, lap)
            rows.append([crc(img), mono & 0xffffffff, len(kps), crc(kps), crc(desc)] + [crc(r.level(l)) for l in range(8)])
        rec["mono_" + cfg] = np.array(rows, np.uint64)
        print("sweep", cfg, "K", [int(x[2]) for x in rows])
    for cfg, base in STEREO:
        w, h, nf, lap, fx, b = synth.CONFIGS[cfg]
        rows = []
        for i in range(8):
            L, R = synth.stereo_pair(base + 100 + i, w, h)
            rL, rR = op.RefExtractor(nf), op.RefExtractor(nf)
            _, kL, dL = rL(L, lap)
            _, kR, dR = rR(R, lap)
            mbf = np.float32(fx * b)
            mb = np.float32(mbf / np.float32(fx))
            u, d = op.ref_stereo(rL, rR, kL, dL, kR, dR, float(mbf), float(mb))
            rows.append([crc(L), crc(R), crc(kL), crc(dL), crc(kR), crc(dR), crc(u), crc(d), int((u >= 0).sum())])
        rec["stereo_" + cfg] = np.array(rows, np.uint64)
        print("sweep stereo", cfg, "matches", [int(x[8]) for x in rows])
    np.savez_compressed(os.path.join(OUT, "seed_sweep_crc.npz"), **rec)


def two_camera():
    """two-camera (Nleft != -1) searches of the reference itself (oracle/_ref: src/ORBmatcher.cc:42-209 and :1521-1733 compiled by
    line range) on a TUM-VI-shape pair extracted by the reference extractor"""
    from oracle import oracle_match2_py as o2
    w, h, nf, lap, fx, b = synth.CONFIGS["tumvi"]
    L, R = synth.stereo_pair(3000, w, h)
    rL, rR = op.RefExtractor(nf), op.RefExtractor(nf)
    _, kL, dL = rL(L, lap)
    _, kR, dR = rR(R, lap)
    gp = om.grid_params(w, h)
    scale = rL.tables()["scale"]
    trl = (-14.25, 0.75)
    out = dict(kps_crc=np.array([crc(kL), crc(kR)], np.uint64))
    for name, th, mono, tlc, ori, jit, pobs in MATCH:
        q, q2, qd = synth.synth_queries2(500, kL, dL, kR, dR, w, h, trl, p_obs=pobs, jitter=jit)
        nm, m = o2.ref_search_by_projection2(kL, dL, kR, dR, scale, gp, 0.1, trl, q, qd, th, mono, tlc, ori)
        out["sbp_" + name] = np.concatenate([[nm], m]).astype(np.int32)
        print("two-camera sbp", name, nm)
    l2r, r2l = synth.synth_stereo_pairing(600, len(kL), len(kR))
    for name, th, ratio, jit, pobs, plock in LOCAL:
        q, qd = synth.synth_track_queries2(700, kL, dL, kR, dR, l2r, w, h, p_obs=pobs, jitter=jit)
        lk = (np.random.default_rng(800).random(len(kL) + len(kR)) < plock).astype(np.uint8)
        nm, m = o2.ref_search_local_points2(kL, dL, kR, dR, lk, l2r, r2l, scale, gp, q, qd, th, ratio)
        out["local_" + name] = np.concatenate([[nm], m]).astype(np.int32)
        print("two-camera local", name, nm)
    np.savez_compressed(os.path.join(OUT, "two_camera_match.npz"), **out)


def main():
    op.build()
    assert op.ref_available(), "needs oracle/_ref (the reference mount)"
    os.makedirs(OUT, exist_ok=True)
    seed_sweep()
    two_camera()
    if len(sys.argv) > 1 and sys.argv[1] == "--sweep-only":
        return
    for cfg, seed in CASES:
        w, h, nf, lap, fx, b = synth.CONFIGS[cfg]
        img = synth.mono_frame(seed, w, h)
        r = op.RefExtractor(nf)
        mono, kps, desc = r(img, lap)
        lv = np.array([crc(r.level(l)) for l in range(8)], np.uint64)
        np.savez_compressed(os.path.join(OUT, "extract_%s_%d.npz" % (cfg, seed)), image_crc=np.uint64(crc(img)),
                            mono=np.int32(mono), kps=kps, desc=desc, level_crc=lv)
        print(cfg, seed, "K", len(kps), "mono", mono)
    for cfg, seed in STEREO:
        w, h, nf, lap, fx, b = synth.CONFIGS[cfg]
        L, R = synth.stereo_pair(seed, w, h)
        rL, rR = op.RefExtractor(nf), op.RefExtractor(nf)
        _, kL, dL = rL(L, lap)
        _, kR, dR = rR(R, lap)
        mbf = np.float32(fx * b)
        mb = np.float32(mbf / np.float32(fx))
        u, d = op.ref_stereo(rL, rR, kL, dL, kR, dR, float(mbf), float(mb))
        np.savez_compressed(os.path.join(OUT, "stereo_%s_%d.npz" % (cfg, seed)), image_crc=np.array([crc(L), crc(R)], np.uint64),
                            kps_left_crc=np.uint64(crc(kL)), kps_right_crc=np.uint64(crc(kR)), uright=u, depth=d)
        print("stereo", cfg, seed, "matches", int((u >= 0).sum()), "of", len(u))
    # windowed matcher: ORBmatcher::SearchByProjection (src/ORBmatcher.cc:1521-1733) of the reference itself, current frame =
    # left image of the EuRoC pair (keypoints, descriptors, uRight from the reference), last frame = right image
    for name, th, mono, tlc, ori, jit, pobs in MATCH:
        w, h, nf, lap, fx, b = synth.CONFIGS["euroc"]
        L, R = synth.stereo_pair(2000, w, h)
        rL, rR = op.RefExtractor(nf), op.RefExtractor(nf)
        _, kL, dL = rL(L, lap)
        _, kR, dR = rR(R, lap)
        mbf = np.float32(fx * b)
        mb = np.float32(mbf / np.float32(fx))
        u, _ = op.ref_stereo(rL, rR, kL, dL, kR, dR, float(mbf), float(mb))
        gp = om.grid_params(w, h)
        q, qd = om.synth_queries(77, kR, dR, None, None, w, h, p_obs=pobs, jitter=jit)
        nm, match = om.reference().search_by_projection(kL, dL, u, rL.tables()["scale"], gp, float(np.float32(b)), float(mbf), q, qd, th,
                                                        mono, tlc, ori)
        np.savez_compressed(os.path.join(OUT, "match_%s.npz" % name), q_crc=np.uint64(crc(q)), qd_crc=np.uint64(crc(qd)),
                            kps_left_crc=np.uint64(crc(kL)), nmatches=np.int32(nm), match=match)
        print("match", name, "nmatches", nm)
    # local-map search: ORBmatcher::SearchByProjection(F, vpMapPoints, th) (src/ORBmatcher.cc:42-209) of the reference itself,
    # F = left image of the EuRoC pair, map points derived from its own keypoints
    for name, th, ratio, jit, pobs, plock in LOCAL:
        w, h, nf, lap, fx, b = synth.CONFIGS["euroc"]
        L, R = synth.stereo_pair(2000, w, h)
        rL, rR = op.RefExtractor(nf), op.RefExtractor(nf)
        _, kL, dL = rL(L, lap)
        _, kR, dR = rR(R, lap)
        mbf = np.float32(fx * b)
        mb = np.float32(mbf / np.float32(fx))
        u, _ = op.ref_stereo(rL, rR, kL, dL, kR, dR, float(mbf), float(mb))
        gp = om.grid_params(w, h)
        q, qd = synth.synth_track_queries(78, kL, dL, u, w, h, p_obs=pobs, jitter=jit, mbf=float(mbf))
        locked0 = (np.random.default_rng(79).random(len(kL)) < plock).astype(np.uint8)
        nm, match = om.reference().search_local_points(kL, dL, u, locked0, rL.tables()["scale"], gp, q, qd, th, ratio)
        np.savez_compressed(os.path.join(OUT, "local_%s.npz" % name), q_crc=np.uint64(crc(q)), qd_crc=np.uint64(crc(qd)),
                            kps_left_crc=np.uint64(crc(kL)), nmatches=np.int32(nm), match=match)
        print("local", name, "nmatches", nm, "of", len(q))
    # bag of words: the reference's own DBoW2 (text loader + transform) on the descriptors of the EuRoC left image
    import tempfile
    from oracle import oracle_bow_py as ob
    w, h, nf, lap, fx, b = synth.CONFIGS["euroc"]
    _, kL, dL = op.RefExtractor(nf)(synth.stereo_pair(2000, w, h)[0], lap)
    for name, seed, k, Lv, pe, ps, scoring, weighting, levelsup in BOW:
        voc = synth.synth_vocabulary(seed, k, Lv, pe, ps, scoring=scoring, weighting=weighting)
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "voc.txt")
            synth.write_vocabulary_text(voc, path)
            r = ob.ReferenceVocabulary(path).transform(dL, levelsup)
        np.savez_compressed(os.path.join(OUT, "bow_%s.npz" % name), desc_crc=np.uint64(crc(dL)), voc_crc=np.uint64(crc(voc["desc"]) ^ crc(voc["weight"])),
                            **r)
        print("bow", name, "words", len(r["bow_word"]), "nodes", len(r["fv_node"]))
    import cv2
    q = synth.random_descriptors(0, 1200)
    db = synth.clustered_descriptors(2, q, 100000, max_flips=80)
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, db, k=2)
    idx = np.array([[x.trainIdx for x in mm] for mm in m], np.int32)
    dist = np.array([[int(x.distance) for x in mm] for mm in m], np.int32)
    np.savez_compressed(os.path.join(OUT, "knn2_1200x100k.npz"), q_crc=np.uint64(crc(q)), db_crc=np.uint64(crc(db)), idx=idx, dist=dist)
    print("knn ties:", int((dist[:, 0] == dist[:, 1]).sum()))


if __name__ == "__main__":
    main()
