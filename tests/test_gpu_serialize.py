"""orb_serialize_frame on the GPU: the Atlas fragments of the device-resident results are byte-identical to
what the reference's serializeVectorKeyPoints / serializeMatrix (include/SerializationUtils.h:74-152) write for the same keypoints
and descriptors (oracle: oracle/oracle_ser_py.py, equal to the reference's templates by tests/test_oracle_ser.py)."""
import numpy as np
import pytest

from morb_slam_b200 import capi, synth
from oracle import oracle_ser_py as osr

pytestmark = pytest.mark.gpu


def test_fragments_of_a_batch():
    w, h, nf, lap, fx, b = synth.CONFIGS["euroc"]
    B = 3
    imgs = np.stack([synth.mono_frame(8100, w, h), synth.flat_frame(8101, w, h), synth.mono_frame(8102, w, h)])
    ex = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B)
    n, mono, kps, desc = ex.extract_batch(imgs, lap)
    ref = osr.Reference() if osr.have_reference() else None
    with pytest.raises(capi.OrbError):
        capi.serialize_frame(ex, 0, capi.ORB_SER_KEYS_UN)      # mvKeysUn does not exist yet
    K = np.float32([[458.654, 0, 367.215], [0, 457.296, 248.375], [0, 0, 1]])
    dist = np.float32([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05])
    kun = capi.undistort_keypoints(ex, K, dist, K)
    for f in range(B):
        kb = capi.serialize_frame(ex, f, capi.ORB_SER_KEYS)
        ub = capi.serialize_frame(ex, f, capi.ORB_SER_KEYS_UN)
        db = capi.serialize_frame(ex, f, capi.ORB_SER_DESCRIPTORS)
        assert kb == osr.oracle_serialize_keypoints(kps[f, :n[f]])
        assert ub == osr.oracle_serialize_keypoints(kun[f, :n[f]])
        if n[f]:
            assert db == osr.oracle_serialize_matrix(desc[f, :n[f]])
        else:
            assert db == np.array([32, 0, 0], np.int32).tobytes() + b"\x01"
        if ref is not None:
            assert kb == ref.serialize_keypoints(kps[f, :n[f]])
            assert ref.deserialize_keypoints(kb).tobytes() == kps[f, :n[f]].tobytes()
            if n[f]:
                assert db == ref.serialize_matrix(desc[f, :n[f]])
        assert capi.deserialize_keypoints(kb).tobytes() == kps[f, :n[f]].tobytes()
        assert np.array_equal(capi.deserialize_descriptors(db), desc[f, :n[f]])
    assert n[0] > 1000
    with pytest.raises(capi.OrbError):
        capi.serialize_frame(ex, B, capi.ORB_SER_KEYS)
    with pytest.raises(capi.OrbError):
        capi.serialize_frame(ex, 0, 7)
