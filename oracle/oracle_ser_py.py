"""TEST INFRASTRUCTURE ONLY. The Atlas fragments of the front-end's results: a numpy restatement of what
serializeVectorKeyPoints / serializeMatrix (reference include/SerializationUtils.h:74-152) put into a binary archive, and ctypes
bindings of the reference's own templates instantiated on a raw-bytes archive (oracle/_ref/libmorb_ref_ser.so, ref_driver_ser.cc).
Same import rules as oracle_py."""
import ctypes as C
import os

import numpy as np

from oracle.oracle_py import KP_DTYPE, HERE, _Lib, _p

REF_SER_SO = os.path.join(HERE, "_ref", "libmorb_ref_ser.so")
# field order of the archive (:135-141): angle, response, size, pt.x, pt.y, class_id, octave
SER_KP_DTYPE = np.dtype([("angle", "<f4"), ("response", "<f4"), ("size", "<f4"), ("x", "<f4"), ("y", "<f4"), ("class_id", "<i4"), ("octave", "<i4")])


def oracle_serialize_keypoints(kps):
    kps = np.ascontiguousarray(kps, dtype=KP_DTYPE)
    rec = np.zeros(len(kps), SER_KP_DTYPE)
    for f in SER_KP_DTYPE.names:
        rec[f] = kps[f]
    return np.int32(len(kps)).tobytes() + rec.tobytes()


def oracle_serialize_matrix(mat):
    """8UC1 matrix (possibly a strided view): cols, rows, type, continuous, then the rows (:74-99)"""
    mat = np.asarray(mat, np.uint8)
    rows, cols = mat.shape
    cont = mat.flags["C_CONTIGUOUS"] or rows == 1
    return np.array([cols, rows, 0], np.int32).tobytes() + bytes([1 if cont else 0]) + np.ascontiguousarray(mat).tobytes()


def oracle_deserialize_keypoints(buf):
    n = int(np.frombuffer(buf[:4], np.int32)[0])
    rec = np.frombuffer(buf[4:4 + 28 * n], SER_KP_DTYPE)
    kps = np.zeros(n, KP_DTYPE)
    for f in SER_KP_DTYPE.names:
        kps[f] = rec[f]
    return kps


def have_reference():
    return os.path.exists(REF_SER_SO)


class Reference:
    def __init__(self):
        self.lib = _Lib.load(REF_SER_SO)
        self.lib.ref_serialize_keypoints.restype = C.c_size_t
        self.lib.ref_serialize_keypoints.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        self.lib.ref_serialize_matrix.restype = C.c_size_t
        self.lib.ref_serialize_matrix.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p]
        self.lib.ref_deserialize_keypoints.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        self.lib.ref_deserialize_matrix.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]

    def serialize_keypoints(self, kps):
        kps = np.ascontiguousarray(kps, dtype=KP_DTYPE)
        out = np.zeros(self.lib.ref_serialize_keypoints(_p(kps), len(kps), None), np.uint8)
        self.lib.ref_serialize_keypoints(_p(kps), len(kps), _p(out))
        return out.tobytes()

    def serialize_matrix(self, mat):
        mat = np.asarray(mat, np.uint8)
        assert mat.strides[1] == 1
        rows, cols = mat.shape
        stride = mat.strides[0] if rows > 1 else cols
        base = mat.ctypes.data_as(C.c_void_p)
        out = np.zeros(self.lib.ref_serialize_matrix(base, rows, cols, stride, None), np.uint8)
        self.lib.ref_serialize_matrix(base, rows, cols, stride, _p(out))
        return out.tobytes()

    def deserialize_keypoints(self, buf, cap=100000):
        b = np.frombuffer(buf, np.uint8).copy()
        kps = np.zeros(cap, KP_DTYPE)
        n = self.lib.ref_deserialize_keypoints(_p(b), _p(kps), cap)
        return kps[:n]

    def deserialize_matrix(self, buf, cap=1 << 24):
        b = np.frombuffer(buf, np.uint8).copy()
        data = np.zeros(cap, np.uint8)
        cols = C.c_int(0)
        rows = self.lib.ref_deserialize_matrix(_p(b), _p(data), cap, C.byref(cols))
        return data[:rows * cols.value].reshape(rows, cols.value)
