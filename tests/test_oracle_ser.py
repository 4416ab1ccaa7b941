"""Atlas fragments of the front-end's results: numpy restatement and the product's host-side loaders against
the reference's own serializeVectorKeyPoints / serializeMatrix (include/SerializationUtils.h:74-152) instantiated on a raw-bytes
archive (oracle/_ref/libmorb_ref_ser.so). CPU only: the loaders are plain host functions of the C ABI."""
import numpy as np
import pytest

from morb_slam_b200 import capi
from oracle import oracle_ser_py as osr
from oracle.oracle_py import KP_DTYPE

needs_ref = pytest.mark.skipif(not osr.have_reference(), reason="oracle/_ref/libmorb_ref_ser.so not built (no /root/reference)")


def _kps(seed, n):
    rng = np.random.default_rng(seed)
    k = np.zeros(n, KP_DTYPE)
    for f in ("x", "y", "size", "angle", "response"):
        k[f] = rng.uniform(-1000, 1000, n).astype(np.float32)
    k["octave"] = rng.integers(0, 8, n); k["class_id"] = rng.integers(-1, 5, n)
    return k


@needs_ref
def test_keypoint_fragment_equals_reference_template():
    ref = osr.Reference()
    for n in (0, 1, 7, 1207):
        k = _kps(n, n)
        b = ref.serialize_keypoints(k)
        assert len(b) == 4 + 28 * n == capi.lib().orb_serialized_keypoints_size(n)
        assert osr.oracle_serialize_keypoints(k) == b
        # loading branch of the template, the restatement and the product's host loader agree
        assert ref.deserialize_keypoints(b).tobytes() == k.tobytes()
        assert osr.oracle_deserialize_keypoints(b).tobytes() == k.tobytes()
        assert capi.deserialize_keypoints(b).tobytes() == k.tobytes()


@needs_ref
def test_matrix_fragment_equals_reference_template():
    ref = osr.Reference()
    rng = np.random.default_rng(3)
    for n in (1, 5, 1207):
        d = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        b = ref.serialize_matrix(d)
        assert len(b) == 13 + 32 * n == capi.lib().orb_serialized_descriptors_size(n)
        assert osr.oracle_serialize_matrix(d) == b
        assert np.array_equal(ref.deserialize_matrix(b), d) and np.array_equal(capi.deserialize_descriptors(b), d)
    # a strided view (mDescriptors.rowRange / colRange style) takes the per-row branch: continuous = 0, same payload
    wide = rng.integers(0, 256, (9, 48), dtype=np.uint8)
    view = wide[:, 8:40]
    b = ref.serialize_matrix(view)
    assert b[12] == 0 and osr.oracle_serialize_matrix(view) == b and b[13:] == np.ascontiguousarray(view).tobytes()


def test_loader_error_paths():
    k = _kps(1, 5)
    b = osr.oracle_serialize_keypoints(k)
    with pytest.raises(capi.OrbError):
        capi.deserialize_keypoints(b[:-1])           # truncated
    with pytest.raises(capi.OrbError):
        capi.deserialize_keypoints(b, cap=3)         # caller capacity
    d = osr.oracle_serialize_matrix(np.zeros((4, 32), np.uint8))
    with pytest.raises(capi.OrbError):
        capi.deserialize_descriptors(d[:20])
    bad = bytearray(d); bad[0] = 31                  # cols != 32
    with pytest.raises(capi.OrbError):
        capi.deserialize_descriptors(bytes(bad))
    assert len(capi.deserialize_keypoints(np.int32(0).tobytes())) == 0
