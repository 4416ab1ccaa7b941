"""Bag of words (Frame::ComputeBoW src/Frame.cc:822-827): the CPU restatement
(oracle/orb_oracle_bow.cc) against the reference's own Thirdparty/DBoW2 compiled unmodified
(oracle/_ref/libmorb_ref_bow.so), which loads the vocabulary from the ORBvoc.txt text format with its own parser.
CPU only; skipped where /root/reference was never mounted."""
import numpy as np
import pytest

from morb_slam_b200 import synth
from oracle import oracle_py as op
from oracle import oracle_bow_py as ob

pytestmark = pytest.mark.skipif(not ob.have_reference(), reason="oracle/_ref/libmorb_ref_bow.so not built (no /root/reference)")

KEYS = ("bow_word", "bow_val", "fv_node", "fv_off", "fv_feat")

VOCABS = [
    # seed, k, L, p_early_leaf, p_short, scoring, weighting
    (1, 10, 3, 0.0, 0.0, 0, 0),     # ORBvoc settings: L1_NORM, TF_IDF
    (2, 10, 4, 0.05, 0.1, 0, 0),    # ragged tree: leaves above level L, nodes with fewer than k children
    (3, 6, 5, 0.02, 0.2, 1, 0),     # L2_NORM
    (4, 9, 3, 0.0, 0.0, 5, 1),      # DOT_PRODUCT (no normalisation, divide by size), TF
    (5, 4, 6, 0.0, 0.0, 0, 2),      # IDF
    (6, 3, 4, 0.0, 0.3, 2, 3),      # CHI_SQUARE, BINARY
]


def same(a, b):
    for k in KEYS:
        if a[k].dtype == np.float64:
            assert a[k].tobytes() == b[k].tobytes(), k     # doubles bit for bit
        else:
            assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("seed,k,L,pe,ps,scoring,weighting", VOCABS)
def test_transform_equals_reference(tmp_path, seed, k, L, pe, ps, scoring, weighting):
    op.build()
    voc = synth.synth_vocabulary(seed, k, L, pe, ps, scoring=scoring, weighting=weighting)
    path = str(tmp_path / "voc.txt")
    synth.write_vocabulary_text(voc, path)
    ref = ob.ReferenceVocabulary(path)
    info = ref.info()
    assert info["k"] == k and info["L"] == L and info["scoring"] == scoring and info["weighting"] == weighting
    assert info["words"] == int(voc["is_leaf"].sum())
    orc = ob.OracleVocabulary(voc)
    # node level of the feature vector must not lie above a leaf (the reference leaves *nid uninitialised there)
    for levelsup in ((4, 2, 0, L, L + 3) if pe == 0 else (L - 2, L - 1, L + 2)):
        for n, pw in ((1200, 0.6), (300, 1.0), (1, 0.0), (0, 0.0)):
            d = synth.synth_bow_descriptors(10 * seed + n, voc, n, pw)
            a, b = orc.transform(d, levelsup), ref.transform(d, levelsup)
            same(a, b)
            if n >= 300:
                assert len(a["bow_word"]) < n and len(a["bow_word"]) > n // 20      # words repeat inside the image
                if weighting < 2 and scoring != 5:
                    assert abs((np.abs(a["bow_val"]) if scoring != 1 else a["bow_val"] ** 2).sum() - 1.0) < 1e-9


def test_known_answers(tmp_path):
    """Hand-checkable vocabulary: k = 2, L = 2, descriptors all-zero / all-one patterns."""
    op.build()
    z, o = np.zeros(32, np.uint8), np.full(32, 255, np.uint8)
    half = np.concatenate([np.zeros(16, np.uint8), np.full(16, 255, np.uint8)])
    # nodes: 0 root; 1 (zeros) and 2 (ones) under the root; 3, 4 under 1; 5, 6 under 2
    voc = dict(k=2, L=2, scoring=0, weighting=0, parent=np.array([0, 0, 0, 1, 1, 2, 2], np.int32),
               is_leaf=np.array([0, 0, 0, 1, 1, 1, 1], np.uint8),
               desc=np.stack([z, z, o, z, half, o, half]), weight=np.array([0, 0, 0, 1.0, 2.0, 4.0, 0.0]))
    path = str(tmp_path / "voc.txt")
    synth.write_vocabulary_text(voc, path)
    ref, orc = ob.ReferenceVocabulary(path), ob.OracleVocabulary(voc)
    feats = np.stack([z, z, half, o, z])           # half ties between node 1 and 2 -> first child (1); then ties 3 / 4 at 128 -> 4? no: 0 vs 128
    for impl in (ref, orc):
        r = impl.transform(feats, 1)
        # z -> node 1 -> word 0 (node 3, weight 1) three times; half -> node 1 (tie keeps the first) -> node 4 (distance 0, word 1, weight 2);
        # o -> node 2 -> node 5 (word 2, weight 4)
        assert list(r["bow_word"]) == [0, 1, 2]
        assert np.allclose(r["bow_val"], np.array([3.0, 2.0, 4.0]) / 9.0)
        assert list(r["fv_node"]) == [1, 2] and list(r["fv_off"]) == [0, 4, 5] and list(r["fv_feat"]) == [0, 1, 2, 4, 3]
    # a feature that lands on the stopped word (node 6, weight 0) is dropped from both vectors
    near6 = half.copy(); near6[0] = 255      # 136 ones: closer to node 2 than to 1, then 8 bits from node 6, 120 from node 5
    for impl in (ref, orc):
        r = impl.transform(np.stack([near6, z]), 1)
        assert list(r["bow_word"]) == [0] and list(r["fv_feat"]) == [1]
