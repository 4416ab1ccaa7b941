"""Bag of words on the GPU (include/orb_b200.h: orb_vocab_create / orb_vocab_load_text / orb_compute_bow) against the CPU
oracle restatement (oracle/orb_oracle_bow.cc, pinned bit for bit against the reference's own DBoW2 by
tests/test_oracle_bow.py). Bit-exact: word ids, node ids, feature lists and the normalised double values."""
import numpy as np
import pytest

from morb_slam_b200 import capi, synth
from oracle import oracle_py as op
from oracle import oracle_bow_py as ob

pytestmark = pytest.mark.gpu

KEYS = ("bow_word", "bow_val", "fv_node", "fv_off", "fv_feat")


@pytest.fixture(scope="module")
def frames():
    op.build()
    w, h, nf, lap, fx, b = synth.CONFIGS["euroc"]
    B = 5
    imgs = np.stack([synth.stereo_pair(4300 + i, w, h)[0] for i in range(B)])
    imgs[4, :, :] = 128                    # a frame without keypoints
    ex = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B)
    n, _, kps, desc = ex.extract_batch(imgs, lap)
    assert n[4] == 0 and n[0] > 1000
    return dict(ex=ex, B=B, n=n, desc=desc)


def check(frames, voc, gv, levelsup):
    orc = ob.OracleVocabulary(voc)
    got = capi.compute_bow(frames["ex"], gv, levelsup)
    for f in range(frames["B"]):
        n = int(frames["n"][f])
        want = orc.transform(frames["desc"][f, :n], levelsup)
        for k in KEYS:
            assert got[f][k].tobytes() == want[k].tobytes(), (f, k)
        assert np.array_equal(got[f]["feat_word"][:n], want["feat_word"]) and np.array_equal(got[f]["feat_node"][:n], want["feat_node"]), f
    return got


VOCABS = [
    # seed, k, L, p_early_leaf, p_short, scoring, weighting, levelsups
    (11, 10, 4, 0.0, 0.0, 0, 0, (4, 2, 0, 7)),     # ORBvoc settings: L1_NORM, TF_IDF, levelsup = 4 as Frame::ComputeBoW
    (12, 10, 3, 0.05, 0.1, 0, 0, (1, 2)),          # ragged tree
    (13, 6, 5, 0.0, 0.2, 1, 0, (4,)),              # L2_NORM
    (14, 9, 3, 0.0, 0.0, 5, 1, (2,)),              # DOT_PRODUCT, TF
    (15, 4, 6, 0.0, 0.0, 0, 2, (4,)),              # IDF
    (16, 3, 4, 0.0, 0.3, 2, 3, (4,)),              # CHI_SQUARE, BINARY
    (17, 20, 2, 0.0, 0.0, 0, 0, (1,)),             # widest tree the text format accepts: 32-lane groups
    (18, 2, 9, 0.0, 0.0, 0, 0, (4,)),              # deepest / narrowest: 4-lane groups, few words (heavy repetition)
]


@pytest.mark.parametrize("seed,k,L,pe,ps,scoring,weighting,levelsups", VOCABS)
def test_compute_bow_equals_oracle(frames, seed, k, L, pe, ps, scoring, weighting, levelsups):
    voc = synth.synth_vocabulary(seed, k, L, pe, ps, scoring=scoring, weighting=weighting)
    gv = capi.ORBVocabulary(voc)
    info = gv.info()
    assert info["nodes"] == len(voc["parent"]) and info["words"] == int(voc["is_leaf"].sum())
    for levelsup in levelsups:
        got = check(frames, voc, gv, levelsup)
    assert len(got[0]["bow_word"]) > 20 and len(got[4]["bow_word"]) == 0 and len(got[4]["fv_node"]) == 0
    gv.close()


def test_vocabulary_text_loader(frames, tmp_path):
    """orb_vocab_load_text (replaces TemplatedVocabulary::loadFromTextFile) on the ORBvoc.txt format: same transform as
    the vocabulary built from the arrays; a header the reference rejects is an error."""
    voc = synth.synth_vocabulary(21, 10, 3, 0.03, 0.1)
    path = str(tmp_path / "voc.txt")
    synth.write_vocabulary_text(voc, path)
    gv = capi.ORBVocabulary(path=path)
    info = gv.info()
    assert (info["k"], info["L"], info["scoring"], info["weighting"]) == (10, 3, 0, 0)
    assert info["nodes"] == len(voc["parent"]) and info["words"] == int(voc["is_leaf"].sum())
    check(frames, voc, gv, 2)
    with open(path, "a") as f:
        f.write("\n")                      # trailing newline (skipped here; the reference would add a garbage node)
    gv2 = capi.ORBVocabulary(path=path)
    assert gv2.info() == info
    bad = str(tmp_path / "bad.txt")
    with open(bad, "w") as f:
        f.write("30 3 0 0\n0 1 " + " ".join(["0"] * 32) + " 1.0")
    with pytest.raises(capi.OrbError):
        capi.ORBVocabulary(path=bad)
    with pytest.raises(capi.OrbError):
        capi.ORBVocabulary(path=str(tmp_path / "missing.txt"))
    # a truncated node line must not borrow tokens from the next line (the reference parses line by line): rejected
    lines = open(path).read().splitlines()
    short = str(tmp_path / "short.txt")
    with open(short, "w") as f:
        f.write("\n".join(lines[:3] + [" ".join(lines[3].split()[:20])] + lines[4:]) + "\n")
    with pytest.raises(capi.OrbError):
        capi.ORBVocabulary(path=short)
    with pytest.raises(capi.OrbError):
        capi.ORBVocabulary(path="/dev/stdin" if False else "/proc/self/environ/none")   # unreadable path


def test_empty_vocabulary_and_stopped_words(frames):
    z = np.zeros((1, 32), np.uint8)
    empty = dict(k=10, L=3, scoring=0, weighting=0, parent=np.zeros(1, np.int32), is_leaf=np.zeros(1, np.uint8), desc=z,
                 weight=np.zeros(1))
    gv = capi.ORBVocabulary(empty)
    got = capi.compute_bow(frames["ex"], gv, 4)
    assert all(len(g["bow_word"]) == 0 and len(g["fv_node"]) == 0 for g in got)
    # every word stopped: nothing survives
    voc = synth.synth_vocabulary(31, 5, 2, p_stop=1.0)
    got = check(frames, voc, capi.ORBVocabulary(voc), 1)
    assert all(len(g["bow_word"]) == 0 for g in got)


SBOW_CASES = [
    # vocabulary (k, L), levelsup, nnratio, check orientation, share of keyframe keypoints with a map point
    ((10, 4), 2, 0.7, True, 0.8),
    ((10, 4), 3, 0.75, True, 1.0),
    ((6, 3), 3, 0.9, False, 0.5),      # one node holds everything: groups of > 1000 keypoints, lanes loop
    ((10, 4), 0, 0.7, True, 0.8),
]


@pytest.mark.parametrize("kl,levelsup,ratio,ori,pmp", SBOW_CASES)
def test_search_by_bow_equals_oracle(frames, kl, levelsup, ratio, ori, pmp):
    """orb_search_by_bow (ORBmatcher::SearchByBoW, src/ORBmatcher.cc:218-395): frames = the resident batch with the FeatureVectors
    orb_compute_bow left on the device; keyframes = their own descriptors shuffled with a few bit flips plus foreign ones."""
    from oracle import oracle_match_py as om
    ex, B = frames["ex"], frames["B"]
    voc = synth.synth_vocabulary(71, kl[0], kl[1])
    gv = capi.ORBVocabulary(voc)
    ov = ob.OracleVocabulary(voc)
    n, _, kps, desc = ex.extract_batch(np.stack([synth.stereo_pair(4300 + i, 752, 480)[0] for i in range(B)]), (0, 0))
    fvs = capi.compute_bow(ex, gv, levelsup)
    kfs = []
    for f in range(B):
        rng = np.random.default_rng(50 + f)
        m = int(n[f])
        perm = rng.permutation(m)[:min(m, 800)]
        dK = desc[f, perm].copy()
        bits = np.unpackbits(dK, axis=1)
        dK = np.packbits(bits ^ (rng.random(bits.shape) < 0.02).astype(np.uint8), axis=1)
        other = desc[(f + 1) % B, :min(int(n[(f + 1) % B]), 400)]
        dK = np.concatenate([dK, other]) if f != 2 else dK[:0]            # one frame gets an empty keyframe
        aK = np.concatenate([kps[f, perm]["angle"] + rng.normal(0, 3, len(perm)).astype(np.float32),
                             kps[(f + 1) % B, :len(other)]["angle"]]).astype(np.float32)[:len(dK)]
        kfs.append(dict(desc=dK, angle=aK, flags=(rng.random(len(dK)) < pmp).astype(np.uint8), fv=ov.transform(dK, levelsup)))
    nm, match = capi.search_by_bow(ex, kfs, ratio, ori)
    o = om.oracle()
    for f in range(B):
        m = int(n[f])
        no, mo = o.search_by_bow(kfs[f]["desc"], kfs[f]["angle"], kfs[f]["flags"], kfs[f]["fv"], desc[f, :m], kps[f, :m]["angle"], fvs[f], ratio, ori)
        assert nm[f] == no, (f, nm[f], no)
        assert np.array_equal(match[f, :m], mo) and np.all(match[f, m:] == -1), f
    assert nm[0] > 100 and nm[2] == 0
    # needs the frames' FeatureVectors: a fresh extraction without orb_compute_bow is a state error
    ex.extract_batch(np.stack([synth.stereo_pair(4300, 752, 480)[0]]), (0, 0))
    with pytest.raises(capi.OrbError):
        capi.search_by_bow(ex, kfs[:1], ratio, ori)
