"""Parity at BASELINE.json's full sizes (the shapes bench.py times), through properties that do not need the oracle to
run at that size: a 256-pair batch built from 8 distinct pairs in shuffled order must reproduce, slot by slot, the
oracle's result of the corresponding pair (extraction, stereo, grid search, bag of words); the 1.25 M-row kNN shard of
configs[4] is checked with planted neighbours, a numpy brute force on a query subsample and the shard-merge identity."""
import numpy as np
import pytest

from morb_slam_b200 import capi, synth
from oracle import oracle_py as op

pytestmark = pytest.mark.gpu


def test_batch_256_stereo_pairs_equal_oracle_per_slot():
    op.build()
    w, h, nf, lap, fx, b = synth.CONFIGS["euroc"]
    B, D = 256, 8
    pairs = [synth.stereo_pair(5200 + i, w, h) for i in range(D)]
    order = np.random.default_rng(5).permutation(np.arange(B) % D)
    Ls = np.stack([pairs[i][0] for i in order])
    Rs = np.stack([pairs[i][1] for i in order])
    exL = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B)
    exR = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B)
    nL, mL, kL, dL = exL.extract_batch(Ls, lap)
    nR, mR, kR, dR = exR.extract_batch(Rs, lap)
    mbf, maxd = float(np.float32(fx * b)), float(np.float32(fx))
    uR = np.full((B, exL.kcap), -1, np.float32)
    dp = np.full((B, exL.kcap), -1, np.float32)
    capi.compute_stereo_matches_batch(exL, exR, mbf, maxd, out=(uR, dp))
    # bag of words on the resident left descriptors
    voc = synth.synth_vocabulary(51, 10, 3)
    bows = capi.compute_bow(exL, capi.ORBVocabulary(voc), 2)
    from oracle import oracle_bow_py as ob
    ov = ob.OracleVocabulary(voc)
    want = []
    for i in range(D):
        oL, oR = op.OracleExtractor(nf), op.OracleExtractor(nf)
        moL, koL, doL = oL(pairs[i][0], lap)
        moR, koR, doR = oR(pairs[i][1], lap)
        u, d = op.oracle_stereo(oL, oR, koL, doL, koR, doR, mbf, maxd)
        want.append((moL, koL, doL, moR, koR, doR, u, d, ov.transform(doL, 2)))
    for s in range(B):
        moL, koL, doL, moR, koR, doR, u, d, bw = want[order[s]]
        assert nL[s] == len(koL) and mL[s] == moL and nR[s] == len(koR) and mR[s] == moR, s
        assert kL[s, :nL[s]].tobytes() == koL.tobytes() and np.array_equal(dL[s, :nL[s]], doL), s
        assert kR[s, :nR[s]].tobytes() == koR.tobytes() and np.array_equal(dR[s, :nR[s]], doR), s
        assert uR[s, :nL[s]].tobytes() == u.tobytes() and dp[s, :nL[s]].tobytes() == d.tobytes(), s
        for k in ("bow_word", "bow_val", "fv_node", "fv_off", "fv_feat"):
            assert bows[s][k].tobytes() == bw[k].tobytes(), (s, k)


def _popcount_rows(x):
    return np.unpackbits(x, axis=-1).sum(axis=-1, dtype=np.int32)


def test_knn_full_shard_properties():
    """1200 queries x 1 250 000 rows (one GPU's shard of configs[4])."""
    ex = capi.ORBextractor(1000, max_width=752, max_height=480)
    nq, ndb = 1200, 1_250_000
    q = synth.random_descriptors(10, nq)
    db = synth.random_descriptors(11, ndb)
    rng = np.random.default_rng(12)
    # planted neighbours: query i has an exact copy at pos1[i] and a copy with 3 flipped bits at pos2[i]
    pos = rng.choice(ndb, 2 * nq, replace=False)
    pos1, pos2 = pos[:nq], pos[nq:]
    db[pos1] = q
    near = q.copy()
    near[:, 0] ^= 0b00010101
    db[pos2] = near
    idx, dist = capi.hamming_knn2(ex, q, db)
    assert np.array_equal(idx[:, 0], pos1) and np.all(dist[:, 0] == 0)
    assert np.array_equal(idx[:, 1], pos2) and np.all(dist[:, 1] == 3)
    # without the plants: brute force in numpy on a query subsample
    db[pos1] = synth.random_descriptors(13, nq)
    db[pos2] = synth.random_descriptors(14, nq)
    idx, dist = capi.hamming_knn2(ex, q, db)
    for i in rng.choice(nq, 6, replace=False):
        d = _popcount_rows(db ^ q[i])
        o = np.lexsort((np.arange(ndb), d))[:2]          # (distance, index) order = the BFMatcher tie rule
        assert list(idx[i]) == list(o) and list(dist[i]) == [int(d[o[0]]), int(d[o[1]])], i
    # shard identity at full size: two half scans + merge == one scan
    half = ndb // 2
    i0, d0 = capi.hamming_knn2(ex, q, db[:half], index_base=0)
    i1, d1 = capi.hamming_knn2(ex, q, db[half:], index_base=half)
    mi, md = capi.knn2_merge(ex, np.stack([i0, i1]), np.stack([d0, d1]))
    assert np.array_equal(mi, idx) and np.array_equal(md, dist)
    # sortedness of every answer
    assert np.all(dist[:, 0] <= dist[:, 1])
