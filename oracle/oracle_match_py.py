"""TEST INFRASTRUCTURE ONLY. ctypes bindings of the windowed-matcher oracle: the CPU restatement
(oracle/liborb_oracle.so, orb_oracle_match.cc) and the reference's own code compiled by line range
(oracle/_ref/libmorb_ref_match.so, ref_driver_match.cc). Same import rules as oracle_py."""
import ctypes as C
import os

import numpy as np

from oracle.oracle_py import KP_DTYPE, ORACLE_SO, HERE, _Lib, _p

REF_MATCH_SO = os.path.join(HERE, "_ref", "libmorb_ref_match.so")
GRID_COLS, GRID_ROWS = 64, 48

# one query per last-frame keypoint: projected (u, v), depth z, angle and octave of the last-frame keypoint,
# flags bit 0 = map point present and not an outlier, bit 1 = Observations() > 0   (orb_proj_query in include/orb_b200.h)
Q_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("z", "<f4"), ("angle", "<f4"), ("octave", "<i4"), ("flags", "<i4")])
assert Q_DTYPE.itemsize == 24
# orb_track_query: mTrackProjX, mTrackProjY, mTrackProjXR, mTrackViewCos, mnTrackScaleLevel, flags (bit 0 in view, bit 1 observed)
TQ_DTYPE = np.dtype([("proj_x", "<f4"), ("proj_y", "<f4"), ("proj_xr", "<f4"), ("view_cos", "<f4"), ("level", "<i4"), ("flags", "<i4")])


def grid_params(w, h):
    """mnMinX, mnMinY, mnMaxX, mnMaxY, mfGridElementWidthInv, mfGridElementHeightInv for an undistorted w x h image
    (Frame::ComputeImageBounds without distortion, src/Frame.cc:238-241)."""
    minx, miny, maxx, maxy = np.float32(0), np.float32(0), np.float32(w), np.float32(h)
    return np.array([minx, miny, maxx, maxy, np.float32(GRID_COLS) / (maxx - minx), np.float32(GRID_ROWS) / (maxy - miny)],
                    dtype=np.float32)


def _typed(lib, prefix):
    if getattr(lib, "_typed_match", False):
        return lib
    f = getattr(lib, prefix + "assign_grid")
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    f = getattr(lib, prefix + "features_in_area")
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_int]
    f = getattr(lib, prefix + "search_by_projection")
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_void_p,
                  C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int, C.c_void_p]
    f = getattr(lib, prefix + "search_local_points")
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                  C.c_float, C.c_float, C.c_void_p]
    f = getattr(lib, prefix + "search_by_bow")
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p]
    lib._typed_match = True
    return lib


class _Impl:
    def __init__(self, lib, prefix):
        self.lib, self.pre = _typed(lib, prefix), prefix

    def assign_grid(self, kps, gp):
        kps = np.ascontiguousarray(kps, dtype=KP_DTYPE)
        off = np.zeros(GRID_COLS * GRID_ROWS + 1, np.int32)
        idx = np.zeros(max(len(kps), 1), np.int32)
        n = getattr(self.lib, self.pre + "assign_grid")(_p(kps), len(kps), _p(gp), _p(off), _p(idx))
        return off, idx[:n]

    def features_in_area(self, kps, gp, x, y, r, min_level=-1, max_level=-1):
        kps = np.ascontiguousarray(kps, dtype=KP_DTYPE)
        out = np.zeros(max(len(kps), 1), np.int32)
        n = getattr(self.lib, self.pre + "features_in_area")(_p(kps), len(kps), _p(gp), float(x), float(y), float(r), int(min_level),
                                                            int(max_level), _p(out), len(out))
        assert n >= 0
        return out[:n]

    def search_by_projection(self, kps, desc, uright, scale, gp, mb, mbf, q, qdesc, th, mono=False, tlc_z=0.0, check_orientation=True):
        kps = np.ascontiguousarray(kps, dtype=KP_DTYPE)
        desc = np.ascontiguousarray(desc, dtype=np.uint8)
        uright = np.ascontiguousarray(uright, dtype=np.float32)
        scale = np.ascontiguousarray(scale, dtype=np.float32)
        q = np.ascontiguousarray(q, dtype=Q_DTYPE)
        qdesc = np.ascontiguousarray(qdesc, dtype=np.uint8)
        out = np.full(max(len(kps), 1), -1, np.int32)
        nm = getattr(self.lib, self.pre + "search_by_projection")(
            _p(kps), _p(desc), _p(uright), len(kps), _p(scale), len(scale), _p(gp), float(mb), float(mbf), _p(q), _p(qdesc), len(q),
            float(th), int(mono), float(tlc_z), int(check_orientation), _p(out))
        return nm, out[:len(kps)]


    def search_local_points(self, kps, desc, uright, locked0, scale, gp, q, qdesc, th, nnratio=0.8):
        kps = np.ascontiguousarray(kps, dtype=KP_DTYPE)
        desc = np.ascontiguousarray(desc, dtype=np.uint8)
        uright = np.ascontiguousarray(uright, dtype=np.float32)
        locked0 = np.ascontiguousarray(locked0, dtype=np.uint8)
        scale = np.ascontiguousarray(scale, dtype=np.float32)
        q = np.ascontiguousarray(q, dtype=TQ_DTYPE)
        qdesc = np.ascontiguousarray(qdesc, dtype=np.uint8)
        out = np.full(max(len(kps), 1), -1, np.int32)
        nm = getattr(self.lib, self.pre + "search_local_points")(
            _p(kps), _p(desc), _p(uright), _p(locked0), len(kps), _p(scale), len(scale), _p(gp), _p(q), _p(qdesc), len(q), float(th),
            float(nnratio), _p(out))
        return nm, out[:len(kps)]


def _search_by_bow(self, descKF, angleKF, kf_flags, fvKF, descF, angleF, fvF, nnratio=0.7, check_orientation=True):
    """fvKF / fvF: dicts with fv_node, fv_off, fv_feat (map order), as the bag-of-words oracle returns them"""
    descKF = np.ascontiguousarray(descKF, np.uint8); descF = np.ascontiguousarray(descF, np.uint8)
    angleKF = np.ascontiguousarray(angleKF, np.float32); angleF = np.ascontiguousarray(angleF, np.float32)
    kf_flags = np.ascontiguousarray(kf_flags, np.uint8)
    a = [np.ascontiguousarray(fvKF[k], t) for k, t in (("fv_node", np.uint32), ("fv_off", np.int32), ("fv_feat", np.uint32))]
    b = [np.ascontiguousarray(fvF[k], t) for k, t in (("fv_node", np.uint32), ("fv_off", np.int32), ("fv_feat", np.uint32))]
    out = np.full(max(len(descF), 1), -1, np.int32)
    nm = getattr(self.lib, self.pre + "search_by_bow")(_p(descKF), _p(angleKF), _p(kf_flags), len(descKF), _p(a[0]), _p(a[1]), _p(a[2]),
                                                      len(a[0]), _p(descF), _p(angleF), len(descF), _p(b[0]), _p(b[1]), _p(b[2]), len(b[0]),
                                                      float(nnratio), int(check_orientation), _p(out))
    return nm, out[:len(descF)]


_Impl.search_by_bow = _search_by_bow


def oracle():
    return _Impl(_Lib.load(ORACLE_SO), "oro_")


def reference():
    return _Impl(_Lib.load(REF_MATCH_SO), "refm_")


def have_reference():
    return os.path.exists(REF_MATCH_SO)


from morb_slam_b200.synth import synth_queries, synth_track_queries  # noqa: E402,F401  (the generator lives with the other synthetic inputs)
