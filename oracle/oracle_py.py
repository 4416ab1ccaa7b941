"""TEST INFRASTRUCTURE ONLY. ctypes bindings for the CPU oracle (oracle/liborb_oracle.so, the
restatement) and for oracle/_ref/libmorb_ref.so (the reference's own sources compiled unmodified).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module. The product package morb_slam_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liborb_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libmorb_ref.so")

# 28-byte cv::KeyPoint record (pt.x, pt.y, size, angle, response, octave, class_id)
KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28


def build(force=False):
    """Compile the restatement and, when /root/reference is mounted, oracle/_ref."""
    args = ["make", "-s", "-C", HERE] + (["-B"] if force else [])
    subprocess.run(args, check=True, stdout=subprocess.DEVNULL)  # keep the caller's stdout clean (bench prints one JSON line)


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_f32p = C.POINTER(C.c_float)


def _img(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 2
    return a


class _Lib:
    _cache = {}

    @classmethod
    def load(cls, path):
        if path not in cls._cache:
            if not os.path.exists(path):
                raise FileNotFoundError(path + " missing: run `make -C oracle` (or __graft_entry__.build())")
            cls._cache[path] = C.CDLL(path)
        return cls._cache[path]


def oracle_lib():
    lib = _Lib.load(ORACLE_SO)
    if not getattr(lib, "_typed", False):
        lib.oro_create.restype = C.c_void_p
        lib.oro_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        lib.oro_destroy.argtypes = [C.c_void_p]
        lib.oro_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        lib.oro_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        lib.oro_level_size.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        for f in (lib.oro_get_level, lib.oro_get_blurred):
            f.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.oro_get_candidates.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        lib.oro_get_level_keypoints.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        lib.oro_distribute.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        lib.oro_distribute_passes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        lib.oro_stereo.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oro_descriptor_distance.argtypes = [C.c_void_p, C.c_void_p]
        lib.oro_knn2.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
        lib.oro_ratio_test.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.shim_resize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
        lib.shim_gauss7.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.shim_fast.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        lib.shim_fastatan2.restype = C.c_float
        lib.shim_fastatan2.argtypes = [C.c_float, C.c_float]
        lib.shim_border101.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        for f in (lib.restated_sinf, lib.restated_cosf, lib.libm_sinf, lib.libm_cosf):
            f.restype = C.c_float
            f.argtypes = [C.c_float]
        lib.oro_introsort.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        lib._typed = True
    return lib


def ref_lib():
    lib = _Lib.load(REF_SO)
    if not getattr(lib, "_typed", False):
        lib.ref_create.restype = C.c_void_p
        lib.ref_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        lib.ref_destroy.argtypes = [C.c_void_p]
        lib.ref_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        lib.ref_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        lib.ref_level_size.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.ref_get_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.ref_keypoints_per_level.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                                C.c_int, C.c_void_p]
        lib.ref_distribute.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.c_void_p, C.c_int]
        lib.ref_descriptor_distance.argtypes = [C.c_void_p, C.c_void_p]
        lib.ref_std_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        lib.ref_stereo.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        lib._typed = True
    return lib


def ref_available():
    return os.path.exists(REF_SO)


class _ExtractorBase:
    """Shared surface of the two CPU extractors; mirrors ORBextractor (include/ORBextractor.h:44-105)."""
    _prefix = None

    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7):
        self.lib = oracle_lib() if self._prefix == "oro" else ref_lib()
        self.nlevels = nlevels
        self.nfeatures = nfeatures
        self.h = getattr(self.lib, self._prefix + "_create")(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
        self.cap = nfeatures + 3 * nlevels + 64

    def __del__(self):
        try:
            if self.h:
                getattr(self.lib, self._prefix + "_destroy")(self.h)
                self.h = None
        except Exception:
            pass

    def tables(self):
        n = self.nlevels
        sc, inv, s2, is2 = (np.zeros(n, np.float32) for _ in range(4))
        nf = np.zeros(n, np.int32)
        um = np.zeros(16, np.int32)
        getattr(self.lib, self._prefix + "_tables")(self.h, _p(sc), _p(inv), _p(s2), _p(is2), _p(nf), _p(um))
        return dict(scale=sc, inv_scale=inv, sigma2=s2, inv_sigma2=is2, nfeat=nf, umax=um)

    def __call__(self, image, lapping=(0, 0)):
        """operator(): returns (monoIndex, keypoints[KP_DTYPE], descriptors[K,32])."""
        kps = np.zeros(self.cap, KP_DTYPE)
        desc = np.zeros((self.cap, 32), np.uint8)
        n = C.c_int(0)
        if image is None or image.size == 0:
            mono = getattr(self.lib, self._prefix + "_extract")(self.h, None, 0, 0, 0, lapping[0], lapping[1],
                                                                _p(kps), _p(desc), self.cap, C.byref(n))
            return mono, kps[:0], desc[:0]
        img = _img(image)
        self._keep = img
        mono = getattr(self.lib, self._prefix + "_extract")(self.h, _p(img), img.shape[1], img.shape[0],
                                                            img.strides[0], lapping[0], lapping[1], _p(kps), _p(desc),
                                                            self.cap, C.byref(n))
        if mono < -1:
            raise RuntimeError("oracle extract failed: %d" % mono)
        return mono, kps[:n.value].copy(), desc[:n.value].copy()

    def level(self, l):
        w, h = C.c_int(), C.c_int()
        getattr(self.lib, self._prefix + "_level_size")(self.h, l, C.byref(w), C.byref(h))
        out = np.zeros((h.value, w.value), np.uint8)
        getattr(self.lib, self._prefix + "_get_level")(self.h, l, _p(out))
        return out


class OracleExtractor(_ExtractorBase):
    """The restatement (oracle/orb_oracle.cc)."""
    _prefix = "oro"

    def blurred(self, l):
        w, h = C.c_int(), C.c_int()
        self.lib.oro_level_size(self.h, l, C.byref(w), C.byref(h))
        out = np.zeros((h.value, w.value), np.uint8)
        self.lib.oro_get_blurred(self.h, l, _p(out))
        return out

    def candidates(self, l, cap=1 << 16):
        out = np.zeros((cap, 3), np.int32)
        n = self.lib.oro_get_candidates(self.h, l, _p(out), cap)
        assert n >= 0
        return out[:n].copy()

    def level_keypoints(self, l):
        out = np.zeros(self.cap, KP_DTYPE)
        n = self.lib.oro_get_level_keypoints(self.h, l, _p(out), self.cap)
        assert n >= 0
        return out[:n].copy()


class RefExtractor(_ExtractorBase):
    """The reference's ORBextractor compiled unmodified (oracle/_ref)."""
    _prefix = "ref"

    def keypoints_per_level(self, image):
        img = _img(image)
        cap = self.cap
        out = np.zeros(cap, KP_DTYPE)
        counts = np.zeros(self.nlevels, np.int32)
        n = self.lib.ref_keypoints_per_level(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(out), cap,
                                             _p(counts))
        assert n >= 0
        res, o = [], 0
        for c in counts:
            res.append(out[o:o + c].copy())
            o += c
        return res

    def distribute(self, cands_xys, w, h, N, border=16):
        """DistributeOctTree on (x,y,score) candidates relative to the border; region [0,w)x[0,h)."""
        c = np.ascontiguousarray(cands_xys, np.int32)
        kin = np.zeros(len(c), KP_DTYPE)
        kin["x"], kin["y"], kin["response"] = c[:, 0], c[:, 1], c[:, 2]
        kin["size"], kin["angle"], kin["class_id"] = 7, -1, -1
        cap = N + 64
        out = np.zeros(cap, KP_DTYPE)
        n = self.lib.ref_distribute(self.h, _p(kin), len(kin), border, border + w, border, border + h, N, 0, _p(out), cap)
        assert n >= 0, n
        o = out[:n]
        return np.stack([o["x"], o["y"], o["response"]], 1).astype(np.int32)


def oracle_distribute_passes(cands_xys, w, h, N):
    """DistributeOctTree in pass form (orb_oracle.cc: distribute_octree_passes), the formulation of a block-parallel kernel"""
    lib = oracle_lib()
    c = np.ascontiguousarray(cands_xys, np.int32)
    cap = N + 64
    out = np.zeros((cap, 3), np.int32)
    n = lib.oro_distribute_passes(_p(c), len(c), w, h, N, _p(out), cap)
    assert n >= 0, n
    return out[:n].copy()


def oracle_distribute(cands_xys, w, h, N):
    lib = oracle_lib()
    c = np.ascontiguousarray(cands_xys, np.int32)
    cap = N + 64
    out = np.zeros((cap, 3), np.int32)
    n = lib.oro_distribute(_p(c), len(c), w, h, N, _p(out), cap)
    assert n >= 0, n
    return out[:n].copy()


def oracle_stereo(exL, exR, kpsL, descL, kpsR, descR, mbf, maxD, want_best=False):
    lib = oracle_lib()
    nL, nR = len(kpsL), len(kpsR)
    kpsL = np.ascontiguousarray(kpsL); kpsR = np.ascontiguousarray(kpsR)
    descL = np.ascontiguousarray(descL); descR = np.ascontiguousarray(descR)
    uR = np.full(max(nL, 1), -1, np.float32); dp = np.full(max(nL, 1), -1, np.float32)
    bi = np.full(max(nL, 1), -1, np.int32); bd = np.full(max(nL, 1), -1, np.int32)
    lib.oro_stereo(exL.h, exR.h, _p(kpsL), _p(descL), nL, _p(kpsR), _p(descR), nR, mbf, maxD, _p(uR), _p(dp), _p(bi),
                   _p(bd))
    if want_best:
        return uR[:nL], dp[:nL], bi[:nL], bd[:nL]
    return uR[:nL], dp[:nL]


def ref_stereo(exL, exR, kpsL, descL, kpsR, descR, mbf, mb):
    lib = ref_lib()
    nL, nR = len(kpsL), len(kpsR)
    kpsL = np.ascontiguousarray(kpsL); kpsR = np.ascontiguousarray(kpsR)
    descL = np.ascontiguousarray(descL); descR = np.ascontiguousarray(descR)
    uR = np.full(max(nL, 1), -1, np.float32); dp = np.full(max(nL, 1), -1, np.float32)
    if nL:
        lib.ref_stereo(exL.h, exR.h, _p(kpsL), _p(descL), nL, _p(kpsR), _p(descR), nR, mbf, mb, _p(uR), _p(dp))
    return uR[:nL], dp[:nL]


def oracle_knn2(q, db, threads=1):
    lib = oracle_lib()
    q = np.ascontiguousarray(q, np.uint8); db = np.ascontiguousarray(db, np.uint8)
    idx = np.zeros((len(q), 2), np.int32); dist = np.zeros((len(q), 2), np.int32)
    lib.oro_knn2(_p(q), len(q), _p(db), len(db), _p(idx), _p(dist), threads)
    return idx, dist


def oracle_ratio_test(dist):
    lib = oracle_lib()
    d = np.ascontiguousarray(dist, np.int32)
    out = np.zeros(len(d), np.uint8)
    lib.oro_ratio_test(_p(d), len(d), _p(out))
    return out.astype(bool)
