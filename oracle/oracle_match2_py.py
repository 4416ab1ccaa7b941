"""TEST INFRASTRUCTURE ONLY. Two-camera (Nleft != -1, the fisheye rig) halves of the windowed matcher:
  * pure-Python restatement of the reference's loops, statement by statement, in float32 arithmetic (small cases only):
      Frame::AssignFeaturesToGrid / PosInGrid             src/Frame.cc:501-528, 809-820 (mGrid from mvKeys, mGridRight from mvKeysRight)
      Frame::GetFeaturesInArea(..., bRight)               src/Frame.cc:742-807
      ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono)              src/ORBmatcher.cc:1521-1733 (right half :1638-1707)
      ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th, ...)          src/ORBmatcher.cc:42-209   (right half :127-205)
      ORBmatcher::ComputeThreeMaxima                       src/ORBmatcher.cc:1844-1876
  * ctypes bindings of the reference's own code for the same calls (oracle/_ref/libmorb_ref_match.so, ref_driver_match.cc).
Same import rules as oracle_py. Combined index space: keypoint i < Nleft is left keypoint i, otherwise right keypoint i - Nleft."""
import ctypes as C
import math
import os

import numpy as np

from oracle.oracle_py import KP_DTYPE, _Lib, _p
from oracle.oracle_match_py import GRID_COLS, GRID_ROWS, Q_DTYPE, REF_MATCH_SO

F = np.float32
TH_HIGH, HISTO_LENGTH = 100, 30

# orb_proj_query2 (include/orb_b200.h): orb_proj_query + the projection into the right camera
Q2_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("z", "<f4"), ("angle", "<f4"), ("octave", "<i4"), ("flags", "<i4"), ("ur", "<f4"), ("vr", "<f4")])
assert Q2_DTYPE.itemsize == 32
# orb_track_query2: left mTrackProjX / Y / ViewCos / ScaleLevel, right mTrackProjXR / YR / ViewCosR / ScaleLevelR, flags
TQ2_DTYPE = np.dtype([("proj_x", "<f4"), ("proj_y", "<f4"), ("view_cos", "<f4"), ("level", "<i4"), ("proj_xr", "<f4"), ("proj_yr", "<f4"),
                      ("view_cos_r", "<f4"), ("level_r", "<i4"), ("flags", "<i4"), ("pad", "<i4")])
assert TQ2_DTYPE.itemsize == 40


def _round_away(x):
    """C round(): half away from zero"""
    return int(math.floor(float(x) + 0.5)) if x >= 0 else -int(math.floor(-float(x) + 0.5))


def _dist(a, b):
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


class Grid:
    """one camera's grid: cell lists in push_back order (src/Frame.cc:515-527)"""

    def __init__(self, kps, gp):
        self.kps, self.gp = kps, gp
        self.cells = {}
        minx, miny, winv, hinv = F(gp[0]), F(gp[1]), F(gp[4]), F(gp[5])
        for i in range(len(kps)):
            px = _round_away(F(F(kps["x"][i]) - minx) * winv)        # PosInGrid (:809-820)
            py = _round_away(F(F(kps["y"][i]) - miny) * hinv)
            if px < 0 or px >= GRID_COLS or py < 0 or py >= GRID_ROWS:
                continue
            self.cells.setdefault((px, py), []).append(i)

    def features_in_area(self, x, y, r, min_level=-1, max_level=-1):
        gp = self.gp
        x, y, r = F(x), F(y), F(r)
        minx, miny, winv, hinv = F(gp[0]), F(gp[1]), F(gp[4]), F(gp[5])
        out = []
        c0 = max(0, int(math.floor(F(F(F(x - minx) - r) * winv))))
        if c0 >= GRID_COLS:
            return out
        c1 = min(GRID_COLS - 1, int(math.ceil(F(F(F(x - minx) + r) * winv))))
        if c1 < 0:
            return out
        r0 = max(0, int(math.floor(F(F(F(y - miny) - r) * hinv))))
        if r0 >= GRID_ROWS:
            return out
        r1 = min(GRID_ROWS - 1, int(math.ceil(F(F(F(y - miny) + r) * hinv))))
        if r1 < 0:
            return out
        check = (min_level > 0) or (max_level >= 0)
        k = self.kps
        for ix in range(c0, c1 + 1):
            for iy in range(r0, r1 + 1):
                for j in self.cells.get((ix, iy), ()):
                    if check:
                        if k["octave"][j] < min_level:
                            continue
                        if max_level >= 0 and k["octave"][j] > max_level:
                            continue
                    if abs(F(F(k["x"][j]) - x)) < r and abs(F(F(k["y"][j]) - y)) < r:
                        out.append(j)
        return out


def _three_maxima(hist):
    max1 = max2 = max3 = 0
    ind1 = ind2 = ind3 = -1
    for i, s in enumerate(hist):
        if s > max1:
            max3, max2, max1 = max2, max1, s
            ind3, ind2, ind1 = ind2, ind1, i
        elif s > max2:
            max3, max2 = max2, s
            ind3, ind2 = ind2, i
        elif s > max3:
            max3, ind3 = s, i
    if F(max2) < F(0.1) * F(max1):
        ind2 = ind3 = -1
    elif F(max3) < F(0.1) * F(max1):
        ind3 = -1
    return ind1, ind2, ind3


def _bin(angle_last, angle_cur):
    rot = F(F(angle_last) - F(angle_cur))
    if rot < 0.0:
        rot = F(rot + F(360.0))
    b = _round_away(F(rot * F(F(1.0) / F(HISTO_LENGTH))))
    return 0 if b == HISTO_LENGTH else b


def search_by_projection2(kL, dL, kR, dR, scale, gp, mb, q, qdesc, th, mono=False, tlc_z=0.0, check_orientation=True):
    """src/ORBmatcher.cc:1521-1733 with CurrentFrame.Nleft != -1. q: Q2_DTYPE (u, v = left projection, ur, vr = project(Trl * x3Dc))."""
    nL, nR = len(kL), len(kR)
    gl, gr = Grid(kL, gp), Grid(kR, gp)
    desc = np.concatenate([dL, dR]) if nL + nR else np.zeros((0, 32), np.uint8)
    angle = np.concatenate([kL["angle"], kR["angle"]])
    holder = [-1] * (nL + nR)           # CurrentFrame.mvpMapPoints (index of the last-frame keypoint)
    hist = [[] for _ in range(HISTO_LENGTH)]
    nmatches = 0
    fwd = (F(tlc_z) > F(mb)) and not mono
    bwd = (F(-F(tlc_z)) > F(mb)) and not mono
    for i in range(len(q)):
        if not (q["flags"][i] & 1):
            continue
        invz = F(1.0 / float(F(q["z"][i])))
        if invz < 0:
            continue
        u, v = F(q["u"][i]), F(q["v"][i])
        if u < F(gp[0]) or u > F(gp[2]) or v < F(gp[1]) or v > F(gp[3]):
            continue
        octv = int(q["octave"][i])
        radius = F(F(th) * F(scale[octv]))
        lv = (octv, -1) if fwd else ((0, octv) if bwd else (octv - 1, octv + 1))
        vi = gl.features_in_area(u, v, radius, *lv)
        if not vi:
            continue                                                  # :1581 - also skips the right camera
        best, bi = 256, -1
        for i2 in vi:
            if holder[i2] >= 0 and (q["flags"][holder[i2]] & 2):      # holds a map point with observations
                continue
            d = _dist(qdesc[i], desc[i2])
            if d < best:
                best, bi = d, i2
        if best <= TH_HIGH:
            holder[bi] = i
            nmatches += 1
            if check_orientation:
                hist[_bin(q["angle"][i], angle[bi])].append(bi)
        # right camera (:1638-1707)
        ur, vr = F(q["ur"][i]), F(q["vr"][i])
        vi = gr.features_in_area(ur, vr, radius, *lv)
        best, bi = 256, -1
        for i2 in vi:
            h = holder[i2 + nL]
            if h >= 0 and (q["flags"][h] & 2):
                continue
            d = _dist(qdesc[i], desc[i2 + nL])
            if d < best:
                best, bi = d, i2
        if best <= TH_HIGH:
            holder[bi + nL] = i
            nmatches += 1
            if check_orientation:
                hist[_bin(q["angle"][i], kR["angle"][bi])].append(bi + nL)
    if check_orientation:
        keep = _three_maxima([len(h) for h in hist])
        for b in range(HISTO_LENGTH):
            if b in keep:
                continue
            for idx in hist[b]:
                holder[idx] = -1
                nmatches -= 1
    return nmatches, np.array(holder, np.int32)


def _radius_by_viewing_cos(c):
    return F(2.5) if float(F(c)) > 0.998 else F(4.0)


def search_local_points2(kL, dL, kR, dR, locked0, l2r, r2l, scale, gp, q, qdesc, th, nnratio=0.8):
    """src/ORBmatcher.cc:42-209 with F.Nleft != -1. q: TQ2_DTYPE; locked0[N]: keypoint holds a map point with observations at the start."""
    nL, nR = len(kL), len(kR)
    gl, gr = Grid(kL, gp), Grid(kR, gp)
    desc = np.concatenate([dL, dR]) if nL + nR else np.zeros((0, 32), np.uint8)
    PRIOR = -2
    holder = [PRIOR if locked0[i] else -1 for i in range(nL + nR)]

    def locked(i2):
        h = holder[i2]
        return h == PRIOR or (h >= 0 and bool(q["flags"][h] & 2))

    nmatches = 0
    bfactor = F(th) != F(1.0)
    for i in range(len(q)):
        fl = int(q["flags"][i])
        if not (fl & 1) and not (fl & 4):
            continue
        if fl & 1:
            lvl = int(q["level"][i])
            r = _radius_by_viewing_cos(q["view_cos"][i])
            if bfactor:
                r = F(r * F(th))
            vi = gl.features_in_area(q["proj_x"][i], q["proj_y"][i], F(r * F(scale[lvl])), lvl - 1, lvl)
            if vi:
                best = best2 = 256
                bl = bl2 = bi = -1
                for i2 in vi:
                    if locked(i2):
                        continue
                    d = _dist(qdesc[i], desc[i2])
                    if d < best:
                        best2, best, bl2, bl, bi = best, d, bl, int(kL["octave"][i2]), i2
                    elif d < best2:
                        bl2, best2 = int(kL["octave"][i2]), d
                if best <= TH_HIGH:
                    if bl == bl2 and F(best) > F(nnratio) * F(best2):
                        continue                                      # :123 - also skips the right camera
                    holder[bi] = i
                    if l2r[bi] != -1:
                        holder[int(l2r[bi]) + nL] = i
                        nmatches += 1
                    nmatches += 1
        if fl & 4:
            lvl = int(q["level_r"][i])
            if lvl != -1:
                r = _radius_by_viewing_cos(q["view_cos_r"][i])        # no th factor here (:131)
                vi = gr.features_in_area(q["proj_xr"][i], q["proj_yr"][i], F(r * F(scale[lvl])), lvl - 1, lvl)
                if not vi:
                    continue
                best = best2 = 256
                bl = bl2 = bi = -1
                for i2 in vi:
                    if locked(i2 + nL):
                        continue
                    d = _dist(qdesc[i], desc[i2 + nL])
                    if d < best:
                        best2, best, bl2, bl, bi = best, d, bl, int(kR["octave"][i2]), i2
                    elif d < best2:
                        bl2, best2 = int(kR["octave"][i2]), d
                if best <= TH_HIGH:
                    if bl == bl2 and F(best) > F(nnratio) * F(best2):
                        continue
                    if r2l[bi] != -1:
                        holder[int(r2l[bi])] = i
                        nmatches += 1
                    holder[bi + nL] = i
                    nmatches += 1
    return nmatches, np.array([h if h >= 0 else -1 for h in holder], np.int32)


TH_LOW = 50


def search_by_bow2(descKF, angleKF, kf_flags, fvKF, descL, angleL, descR, angleR, fvF, nnratio=0.7, check_orientation=True):
    """src/ORBmatcher.cc:218-395 with F.Nleft != -1. fvKF / fvF: dicts fv_node, fv_off, fv_feat in map order; the frame's features
    are the left keypoints followed by the right ones. Returns (nmatches, match[N])."""
    nL, nR = len(descL), len(descR)
    desc = np.concatenate([descL, descR]) if nL + nR else np.zeros((0, 32), np.uint8)
    angle = np.concatenate([np.asarray(angleL, np.float32), np.asarray(angleR, np.float32)])
    match = [-1] * (nL + nR)
    hist = [[] for _ in range(HISTO_LENGTH)]
    nmatches = 0
    fpos = {int(n): j for j, n in enumerate(fvF["fv_node"])}
    for a, node in enumerate(fvKF["fv_node"]):
        j = fpos.get(int(node))
        if j is None:
            continue
        idsF = [int(x) for x in fvF["fv_feat"][fvF["fv_off"][j]:fvF["fv_off"][j + 1]]]
        for iKF in fvKF["fv_feat"][fvKF["fv_off"][a]:fvKF["fv_off"][a + 1]]:
            iKF = int(iKF)
            if not kf_flags[iKF]:
                continue
            b1 = b2 = b1r = b2r = 256
            bi = bir = -1
            for iF in idsF:
                if match[iF] >= 0:
                    continue
                d = _dist(descKF[iKF], desc[iF])
                if iF < nL and d < b1:
                    b2, b1, bi = b1, d, iF
                elif iF < nL and d < b2:
                    b2 = d
                if iF >= nL and d < b1r:
                    b2r, b1r, bir = b1r, d, iF
                elif iF >= nL and d < b2r:
                    b2r = d
            if b1 <= TH_LOW:
                if F(b1) < F(nnratio) * F(b2):
                    match[bi] = iKF
                    if check_orientation:
                        hist[_bin(angleKF[iKF], angle[bi])].append(bi)
                    nmatches += 1
                if b1r <= TH_LOW:                                     # `|| true`: no ratio test in the right camera (:331-334)
                    match[bir] = iKF
                    if check_orientation:
                        hist[_bin(angleKF[iKF], angle[bir])].append(bir)
                    nmatches += 1
    if check_orientation:
        keep = _three_maxima([len(h) for h in hist])
        for b in range(HISTO_LENGTH):
            if b in keep:
                continue
            for idx in hist[b]:
                match[idx] = -1
                nmatches -= 1
    return nmatches, np.array(match, np.int32)


def ref_search_by_bow2(descKF, angleKF, kf_flags, fvKF, descL, angleL, descR, angleR, fvF, nnratio=0.7, check_orientation=True):
    lib = _ref()
    if not getattr(lib, "_typed_bow2", False):
        vp, i, f = C.c_void_p, C.c_int, C.c_float
        lib.refm_search_by_bow2.argtypes = [vp, vp, vp, i, vp, vp, vp, i, vp, vp, i, vp, vp, i, vp, vp, vp, i, f, i, vp]
        lib._typed_bow2 = True
    descKF, descL, descR = _c(descKF, np.uint8), _c(descL, np.uint8), _c(descR, np.uint8)
    angleKF, angleL, angleR = _c(angleKF, np.float32), _c(angleL, np.float32), _c(angleR, np.float32)
    kf_flags = _c(kf_flags, np.uint8)
    a = [_c(fvKF[k], t) for k, t in (("fv_node", np.uint32), ("fv_off", np.int32), ("fv_feat", np.uint32))]
    b = [_c(fvF[k], t) for k, t in (("fv_node", np.uint32), ("fv_off", np.int32), ("fv_feat", np.uint32))]
    out = np.full(max(len(descL) + len(descR), 1), -1, np.int32)
    nm = lib.refm_search_by_bow2(_p(descKF), _p(angleKF), _p(kf_flags), len(descKF), _p(a[0]), _p(a[1]), _p(a[2]), len(a[0]), _p(descL),
                                 _p(angleL), len(descL), _p(descR), _p(angleR), len(descR), _p(b[0]), _p(b[1]), _p(b[2]), len(b[0]),
                                 float(nnratio), int(check_orientation), _p(out))
    return nm, out[:len(descL) + len(descR)]


# ---- the reference's own code -------------------------------------------------------------------------------------
def have_reference():
    if not os.path.exists(REF_MATCH_SO):
        return False
    return hasattr(_Lib.load(REF_MATCH_SO), "refm_search_by_projection2")


def _ref():
    lib = _Lib.load(REF_MATCH_SO)
    if not getattr(lib, "_typed_match2", False):
        vp, i, f = C.c_void_p, C.c_int, C.c_float
        lib.refm_features_in_area2.argtypes = [vp, i, vp, i, vp, f, f, f, i, i, i, vp, i]
        lib.refm_search_by_projection2.argtypes = [vp, vp, i, vp, vp, i, vp, i, vp, f, f, f, vp, vp, i, f, i, f, i, vp]
        lib.refm_search_local_points2.argtypes = [vp, vp, i, vp, vp, i, vp, vp, vp, vp, i, vp, vp, vp, i, f, f, vp]
        lib._typed_match2 = True
    return lib


def _c(a, t):
    return np.ascontiguousarray(a, dtype=t)


def ref_features_in_area2(kL, kR, gp, x, y, r, min_level, max_level, right):
    kL, kR = _c(kL, KP_DTYPE), _c(kR, KP_DTYPE)
    out = np.zeros(max(len(kL), len(kR), 1), np.int32)
    n = _ref().refm_features_in_area2(_p(kL), len(kL), _p(kR), len(kR), _p(_c(gp, np.float32)), float(x), float(y), float(r), int(min_level),
                                      int(max_level), int(right), _p(out), len(out))
    assert n >= 0
    return list(out[:n])


def ref_search_by_projection2(kL, dL, kR, dR, scale, gp, mb, trl, q, qdesc, th, mono=False, tlc_z=0.0, check_orientation=True):
    """q: Q_DTYPE (the right projection is the left one + trl in the driver's pure-translation stub)"""
    kL, kR, dL, dR = _c(kL, KP_DTYPE), _c(kR, KP_DTYPE), _c(dL, np.uint8), _c(dR, np.uint8)
    scale, q, qdesc = _c(scale, np.float32), _c(q, Q_DTYPE), _c(qdesc, np.uint8)
    out = np.full(max(len(kL) + len(kR), 1), -1, np.int32)
    nm = _ref().refm_search_by_projection2(_p(kL), _p(dL), len(kL), _p(kR), _p(dR), len(kR), _p(scale), len(scale), _p(_c(gp, np.float32)),
                                           float(mb), float(trl[0]), float(trl[1]), _p(q), _p(qdesc), len(q), float(th), int(mono),
                                           float(tlc_z), int(check_orientation), _p(out))
    return nm, out[:len(kL) + len(kR)]


def ref_search_local_points2(kL, dL, kR, dR, locked0, l2r, r2l, scale, gp, q, qdesc, th, nnratio=0.8):
    kL, kR, dL, dR = _c(kL, KP_DTYPE), _c(kR, KP_DTYPE), _c(dL, np.uint8), _c(dR, np.uint8)
    scale, q, qdesc = _c(scale, np.float32), _c(q, TQ2_DTYPE), _c(qdesc, np.uint8)
    locked0, l2r, r2l = _c(locked0, np.uint8), _c(l2r, np.int32), _c(r2l, np.int32)
    out = np.full(max(len(kL) + len(kR), 1), -1, np.int32)
    nm = _ref().refm_search_local_points2(_p(kL), _p(dL), len(kL), _p(kR), _p(dR), len(kR), _p(locked0), _p(l2r), _p(r2l), _p(scale),
                                          len(scale), _p(_c(gp, np.float32)), _p(q), _p(qdesc), len(q), float(th), float(nnratio), _p(out))
    return nm, out[:len(kL) + len(kR)]


def with_right_projection(q, trl):
    """Q_DTYPE -> Q2_DTYPE with (ur, vr) = (u, v) + trl in float32, as the driver's stub computes Trl * x3Dc"""
    q2 = np.zeros(len(q), Q2_DTYPE)
    for k in Q_DTYPE.names:
        q2[k] = q[k]
    q2["ur"] = (q["u"].astype(np.float32) + np.float32(trl[0])).astype(np.float32)
    q2["vr"] = (q["v"].astype(np.float32) + np.float32(trl[1])).astype(np.float32)
    return q2
