"""TEST INFRASTRUCTURE ONLY. ctypes bindings of the fisheye-triangulation oracle: the CPU restatement (oracle/liborb_oracle.so,
orb_oracle_kb8.cc) and the reference's own lines compiled on the mini Eigen stand-in (oracle/_ref/libmorb_ref_kb8.so,
ref_driver_kb8.cc). Same import rules as oracle_py. Parity of this row is a float tolerance (Eigen absent: see mini_eigen.h)."""
import ctypes as C
import os

import numpy as np

from oracle.oracle_py import KP_DTYPE, ORACLE_SO, HERE, _Lib, _p

REF_KB8_SO = os.path.join(HERE, "_ref", "libmorb_ref_kb8.so")
F, I, VP = C.c_float, C.c_int, C.c_void_p


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class _Impl:
    """rig = dict(cam1, cam2 (8 floats), prec1, prec2, R12 (3 x 3), t12 (3))"""

    def __init__(self, lib, pre, is_ref):
        self.lib, self.pre, self.is_ref = lib, pre, is_ref
        if not getattr(lib, "_typed_kb8", False):
            tri = getattr(lib, pre + "kb8_triangulate")
            tri.argtypes = [VP, F, VP, F, VP, VP, VP, VP, VP, VP, I, VP, VP] + ([] if is_ref else [VP])
            getattr(lib, pre + "kb8_unproject").argtypes = [VP, F, VP, I, VP]
            getattr(lib, pre + "kb8_project").argtypes = [VP, VP, I, VP]
            acc = getattr(lib, pre + "fisheye_accept")
            acc.argtypes = [VP, F, VP, F, VP, VP, VP, I, I, VP, I, I, VP, I, VP, VP, I, VP, VP, VP, VP] + ([] if is_ref else [VP, VP])
            if not is_ref:
                lib.oracle_svd4_v.argtypes = [VP, VP, VP]
            lib._typed_kb8 = True

    @staticmethod
    def _rig(rig):
        return (_f32(rig["cam1"]), float(rig["prec1"]), _f32(rig["cam2"]), float(rig["prec2"]), _f32(rig["R12"]).reshape(9), _f32(rig["t12"]))

    def triangulate(self, rig, xy1, xy2, s1, s2):
        """KannalaBrandt8::TriangulateMatches per pair: (ret[n], p3d[n, 3], quantities[n, 7] or None)"""
        c1, p1, c2, p2, R, t = self._rig(rig)
        xy1, xy2, s1, s2 = _f32(xy1), _f32(xy2), _f32(s1), _f32(s2)
        n = len(s1)
        ret = np.zeros(n, np.float32); p3d = np.zeros((n, 3), np.float32)
        f = getattr(self.lib, self.pre + "kb8_triangulate")
        if self.is_ref:
            f(_p(c1), p1, _p(c2), p2, _p(R), _p(t), _p(xy1), _p(xy2), _p(s1), _p(s2), n, _p(ret), _p(p3d))
            return ret, p3d, None
        q = np.zeros((n, 7), np.float32)
        f(_p(c1), p1, _p(c2), p2, _p(R), _p(t), _p(xy1), _p(xy2), _p(s1), _p(s2), n, _p(ret), _p(p3d), _p(q))
        return ret, p3d, q

    def unproject(self, cam, prec, xy):
        xy = _f32(xy); rays = np.zeros((len(xy), 3), np.float32)
        getattr(self.lib, self.pre + "kb8_unproject")(_p(_f32(cam)), float(prec), _p(xy), len(xy), _p(rays))
        return rays

    def project(self, cam, xyz):
        xyz = _f32(xyz); uv = np.zeros((len(xyz), 2), np.float32)
        getattr(self.lib, self.pre + "kb8_project")(_p(_f32(cam)), _p(xyz), len(xyz), _p(uv))
        return uv

    def fisheye_accept(self, rig, kL, monoL, kR, monoR, sigma2, knn_idx, knn_dist):
        """Frame::ComputeStereoFishEyeMatches after knnMatch (src/Frame.cc:1244-1273):
        (mvLeftToRightMatch, mvRightToLeftMatch, mvDepth, mvStereo3Dpoints, code or None, quantities or None)"""
        c1, p1, c2, p2, R, t = self._rig(rig)
        kL = np.ascontiguousarray(kL, dtype=KP_DTYPE); kR = np.ascontiguousarray(kR, dtype=KP_DTYPE)
        sigma2 = _f32(sigma2)
        idx = np.ascontiguousarray(knn_idx, dtype=np.int32); dist = np.ascontiguousarray(knn_dist, dtype=np.int32)
        nL, nR, nq = len(kL), len(kR), len(idx)
        l2r = np.zeros(max(nL, 1), np.int32); r2l = np.zeros(max(nR, 1), np.int32)
        depth = np.zeros(max(nL, 1), np.float32); p3d = np.zeros((max(nL, 1), 3), np.float32)
        f = getattr(self.lib, self.pre + "fisheye_accept")
        args = [_p(c1), p1, _p(c2), p2, _p(R), _p(t), _p(kL), nL, int(monoL), _p(kR), nR, int(monoR), _p(sigma2), len(sigma2), _p(idx), _p(dist), nq,
                _p(l2r), _p(r2l), _p(depth), _p(p3d)]
        if self.is_ref:
            f(*args)
            return l2r[:nL], r2l[:nR], depth[:nL], p3d[:nL], None, None
        code = np.zeros(max(nL, 1), np.int8); q = np.zeros((max(nL, 1), 7), np.float32)
        f(*args, _p(code), _p(q))
        return l2r[:nL], r2l[:nR], depth[:nL], p3d[:nL], code[:nL], q[:nL]

    def jacobi_svd4f(self, A):
        """Eigen::JacobiSVD<Matrix4f>(A, ComputeFullV) restated in float: V (row-major, columns by descending singular value), sv"""
        A = _f32(A).reshape(16); V = np.zeros((4, 4), np.float32); sv = np.zeros(4, np.float32)
        f = self.lib.oracle_jacobi_svd4f
        f.argtypes = [C.c_void_p] * 3
        f(_p(A), _p(V), _p(sv))
        return V, sv

    def svd4_v(self, A):
        """V (columns by descending singular value) and the singular values of a 4 x 4 float matrix"""
        A = _f32(A).reshape(16); V = np.zeros((4, 4), np.float64); sv = np.zeros(4, np.float64)
        self.lib.oracle_svd4_v(_p(A), _p(V), _p(sv))
        return V, sv


def oracle():
    return _Impl(_Lib.load(ORACLE_SO), "oracle_", False)


def reference():
    return _Impl(_Lib.load(REF_KB8_SO), "ref_", True)


def decisions_agree(code_dev, code_or, q, rel=2e-3):
    """Float-tolerance comparison of the accept / reject decisions: equal, or the quantity that decides differently lies within
    `rel` of its threshold (quantities of the oracle: cos, z1, z2, err1, thr1, err2, thr2). Returns the indices that disagree."""
    bad = []
    for i in np.nonzero(np.asarray(code_dev) != np.asarray(code_or))[0]:
        cos, z1, z2, e1, t1, e2, t2 = [float(v) for v in q[i]]
        near = []
        near.append(abs(cos - 0.9998) < 1e-6)                       # cos is within a few float ulps of the gate
        for z in (z1, z2):
            near.append(np.isfinite(z) and abs(z) < rel)
        for e, t in ((e1, t1), (e2, t2)):
            near.append(np.isfinite(e) and abs(e - t) <= rel * max(t, 1e-6) * 10)
        near.append(np.isfinite(z1) and abs(z1 - 1e-4) < 1e-6)
        if not any(near):
            bad.append(int(i))
    return bad
