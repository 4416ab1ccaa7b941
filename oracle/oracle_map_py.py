"""TEST INFRASTRUCTURE ONLY. The LocalMapping / Relocalization matchers (widening beyond SURVEY.md 8): numpy restatements of
  * the search of ORBmatcher::Fuse, both overloads       (reference src/ORBmatcher.cc:1131-1192, :1277-1304)
  * the map surgery that follows it, on the model of oracle/ref_driver_map.cc (:1195-1210, :1307-1317)
  * ORBmatcher::SearchByProjection(Frame, KeyFrame, ...)  (:1735-1842)
  * ORBmatcher::SearchForTriangulation                     (:821-1042, Pinhole::epipolarConstrain src/CameraModels/Pinhole.cpp:125-138)
  * MapPoint::ComputeDistinctiveDescriptors                (src/MapPoint.cc:367-431)
  * ORBmatcher::SearchForInitialization                    (:603-700; its reference lines live in libmorb_ref_match.so)
and ctypes bindings of the reference's own lines (oracle/_ref/libmorb_ref_map.so, ref_driver_map.cc). tests/test_oracle_map.py holds
restatement == reference; the GPU tests compare the library with both. Same import rules as oracle_py."""
import ctypes as C
import os

import numpy as np

from oracle.oracle_py import KP_DTYPE, HERE, _Lib, _p
from oracle import oracle_match_py as om

REF_MAP_SO = os.path.join(HERE, "_ref", "libmorb_ref_map.so")
f32 = np.float32
TH_LOW, HISTO = 50, 30
FQ_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("ur", "<f4"), ("level", "<i4"), ("flags", "<i4")])            # orb_fuse_query
S3_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("level", "<i4"), ("flags", "<i4")])                                  # Sim3PointC of the driver
FP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4"), ("min_dist", "<f4"),
                     ("max_dist", "<f4"), ("level", "<i4"), ("nobs", "<i4"), ("flags", "<i4")])                   # FusePointC of the driver
Q_DTYPE = om.Q_DTYPE
_POP = np.array([bin(i).count("1") for i in range(256)], np.int32)


def hamming(a, b):
    return int(_POP[np.bitwise_xor(a, b)].sum())


def three_maxima(h):
    """ORBmatcher::ComputeThreeMaxima (src/ORBmatcher.cc:1844-1876) on the bin sizes"""
    max1 = max2 = max3 = 0
    i1 = i2 = i3 = -1
    for i, s in enumerate(h):
        s = int(s)
        if s > max1:
            max3, max2, max1 = max2, max1, s
            i3, i2, i1 = i2, i1, i
        elif s > max2:
            max3, max2 = max2, s
            i3, i2 = i2, i
        elif s > max3:
            max3, i3 = s, i
    if f32(max2) < f32(f32(0.1) * f32(max1)):
        i2 = i3 = -1
    elif f32(max3) < f32(f32(0.1) * f32(max1)):
        i3 = -1
    return i1, i2, i3


def rot_bin(a1, a2):
    rot = f32(f32(a1) - f32(a2))
    if rot < 0.0:
        rot = f32(rot + f32(360.0))
    v = f32(rot * f32(f32(1.0) / f32(HISTO)))
    b = int(np.floor(float(v) + 0.5)) if v >= 0 else -int(np.floor(-float(v) + 0.5))   # round(): half away from zero
    return 0 if b == HISTO else b


# ---- Fuse ---------------------------------------------------------------------------------------------------------------------
def fuse_queries(pts, bf):
    """The host glue of Fuse (:1076-1128) on the driver's stub semantics (identity pose and projection, Ow = 0): orb_fuse_query
    records from FusePointC records. Margins of the synthetic inputs keep every float comparison far from its threshold."""
    q = np.zeros(len(pts), FQ_DTYPE)
    prev_valid = False
    for i, p in enumerate(pts):
        null = bool(p["flags"] & 1)
        if (p["flags"] & 4) and i > 0 and prev_valid:     # the same MapPoint object again: same record
            q[i] = q[i - 1]
            continue
        prev_valid = not null
        if null:
            continue
        x, y, z = f32(p["x"]), f32(p["y"]), f32(p["z"])
        ok = not (z < 0)
        invz = f32(1.0) / z if z != 0 else f32(np.inf)
        d = f32(np.sqrt(f32(x * x + f32(y * y + z * z))))
        ok = ok and not (d < p["min_dist"] or d > p["max_dist"])
        dot = f32(x * p["nx"] + f32(y * p["ny"] + z * p["nz"]))
        ok = ok and not (float(dot) < 0.5 * float(d))
        q[i]["u"], q[i]["v"] = x, y
        q[i]["ur"] = f32(x - f32(f32(bf) * invz)) if np.isfinite(invz) else f32(0)
        q[i]["level"] = p["level"]
        q[i]["flags"] = 1 if ok else 0
    return q


def fuse_search(kps, desc, uright, scale, inv_sigma2, gp, q, qdesc, th, mode=0, in_image=True):
    """best keypoint and distance per query: KeyFrame::GetFeaturesInArea (= Frame::GetFeaturesInArea without levels, checked against the
    reference's KeyFrame lines in tests/test_oracle_map.py), level gate, reprojection gates (mode 0), strict '<'."""
    o = om.oracle()
    n = len(q)
    bi = np.full(n, -1, np.int32); bd = np.full(n, 256, np.int32)
    ur_all = np.full(len(kps), -1, np.float32) if uright is None else np.asarray(uright, np.float32)
    for i in range(n):
        if not (q[i]["flags"] & 1):
            continue
        lvl = int(q[i]["level"])
        u, v, ur = f32(q[i]["u"]), f32(q[i]["v"]), f32(q[i]["ur"])
        r = f32(f32(th) * f32(scale[lvl]))
        for idx in o.features_in_area(kps, gp, u, v, r, -1, -1):
            k = kps[idx]
            kl = int(k["octave"])
            if kl < lvl - 1 or kl > lvl:
                continue
            if mode == 0:
                ex, ey = f32(u - k["x"]), f32(v - k["y"])
                if ur_all[idx] >= 0:
                    er = f32(ur - ur_all[idx])
                    e2 = f32(f32(f32(ex * ex) + f32(ey * ey)) + f32(er * er))
                    if float(f32(e2 * f32(inv_sigma2[kl]))) > 7.8:
                        continue
                else:
                    e2 = f32(f32(ex * ex) + f32(ey * ey))
                    if float(f32(e2 * f32(inv_sigma2[kl]))) > 5.99:
                        continue
            d = hamming(qdesc[i], desc[idx])
            if d < bd[i]:
                bd[i], bi[i] = d, idx
    return bi, bd


def fuse_replay(pts, q, bi, bd, kf_mp_nobs, kf_mp_bad, uright, gp, sim3=False):
    """The loop of Fuse around the search (:1067-1084 skips, :1128 image test, :1195-1210 / :1307-1317 surgery) replayed in map-point
    order from the search results, on the map model of ref_driver_map.cc. Returns (nFused, events, repl, kf_final, cand_bad,
    cand_nobs) like refmap_fuse."""
    n, nq = len(kf_mp_nobs), len(pts)
    stereo = (lambda idx: 2 if (uright is not None and uright[idx] >= 0) else 1)
    # objects: candidates 0..nq-1 (aliases for duplicates), own points -2-j
    alias = list(range(nq))
    for i in range(nq):
        if (pts[i]["flags"] & 4) and i > 0 and not (pts[i - 1]["flags"] & 1) and not (pts[i]["flags"] & 1):
            alias[i] = alias[i - 1]
    st = {}
    for i in range(nq):
        st[i] = dict(bad=bool(pts[i]["flags"] & 2), nobs=int(pts[i]["nobs"]), kf_idx=-1)
    kf = [-1] * n
    for j in range(n):
        if kf_mp_nobs[j] >= 0:
            st[-2 - j] = dict(bad=bool(kf_mp_bad[j]), nobs=int(kf_mp_nobs[j]), kf_idx=j)
            kf[j] = -2 - j
    events, repl = [], [-1] * nq
    nested = [False]

    def add_obs(pid, idx):
        if not nested[0]:
            events.append((1, pid, idx))
        st[pid]["kf_idx"] = idx
        st[pid]["nobs"] += stereo(idx)

    def replace(a, b):          # a->Replace(b)
        events.append((2, a, b))
        nested[0] = True
        A, B = st[a], st[b]
        if A["kf_idx"] >= 0:
            A["nobs"] -= stereo(A["kf_idx"])
            if B["kf_idx"] < 0:
                kf[A["kf_idx"]] = b
                add_obs(b, A["kf_idx"])
            else:
                kf[A["kf_idx"]] = -1
            A["kf_idx"] = -1
        B["nobs"] += A["nobs"]
        A["nobs"] = 0
        A["bad"] = True
        nested[0] = False

    already = set(p for p in kf if p != -1 and not st[p]["bad"]) if sim3 else None   # spAlreadyFound (:1233)
    nf = 0
    for i in range(nq):
        if (pts[i]["flags"] & 1) and not sim3:
            continue
        pid = alias[i]
        S = st[pid]
        if sim3:
            if S["bad"] or pid in already:
                continue
        elif S["bad"] or S["kf_idx"] >= 0:
            continue
        if not (q[i]["flags"] & 1):
            continue
        u, v = f32(q[i]["u"]), f32(q[i]["v"])
        if not (u >= gp[0] and u < gp[2] and v >= gp[1] and v < gp[3]):     # KeyFrame::IsInImage
            continue
        if bd[i] <= TH_LOW:
            j = int(bi[i])
            other = kf[j]
            if other != -1:
                if not st[other]["bad"]:
                    if sim3:
                        repl[i] = j if other <= -2 else -1000 - other
                    elif st[other]["nobs"] > S["nobs"]:
                        replace(pid, other)
                    else:
                        replace(other, pid)
            else:
                add_obs(pid, j)
                kf[j] = pid
            nf += 1
    cand_bad = np.array([int(st[i]["bad"]) for i in range(nq)], np.int32)
    cand_nobs = np.array([st[i]["nobs"] for i in range(nq)], np.int32)
    return nf, events, np.array(repl, np.int32), np.array(kf, np.int32), cand_bad, cand_nobs


# ---- SearchByProjection(Frame, KeyFrame, sAlreadyFound, th, ORBdist) ------------------------------------------------------------------
def search_by_projection_kf(kps, desc, locked0, scale, gp, q, qdesc, th, orb_dist, check_orientation=True):
    o = om.oracle()
    nC = len(kps)
    assigned = np.full(nC, -1, np.int32)
    locked = np.zeros(nC, bool) if locked0 is None else np.asarray(locked0[:nC]).astype(bool).copy()
    hist = [[] for _ in range(HISTO)]
    nm = 0
    for i in range(len(q)):
        if not (q[i]["flags"] & 1):
            continue
        u, v = f32(q[i]["u"]), f32(q[i]["v"])
        if u < gp[0] or u > gp[2] or v < gp[1] or v > gp[3]:
            continue
        lvl = int(q[i]["octave"])
        r = f32(f32(th) * f32(scale[lvl]))
        best, bi2 = 256, -1
        for i2 in o.features_in_area(kps, gp, u, v, r, lvl - 1, lvl + 1):
            if locked[i2]:
                continue
            d = hamming(qdesc[i], desc[i2])
            if d < best:
                best, bi2 = d, int(i2)
        if best <= orb_dist:
            assigned[bi2] = i
            locked[bi2] = True
            nm += 1
            if check_orientation:
                hist[rot_bin(q[i]["angle"], kps[bi2]["angle"])].append(bi2)
    if check_orientation:
        keep = three_maxima([len(b) for b in hist])
        for b in range(HISTO):
            if b not in keep:
                for i2 in hist[b]:
                    assigned[i2] = -1
                    nm -= 1
    return nm, assigned


# ---- SearchForTriangulation -----------------------------------------------------------------------------------------------------------
def search_for_triangulation(k1, k2, scale, sigma2, F12, ep, only_stereo=False, coarse=False, check_orientation=True):
    """k1 / k2: dicts kps, desc, uright (or None), has_mp, fv (fv_node, fv_off, fv_feat). Returns (nmatches, match12[n1])."""
    F = np.asarray(F12, np.float32).reshape(3, 3)
    n1 = len(k1["kps"])
    m12 = np.full(n1, -1, np.int32)
    u1 = np.full(n1, -1, np.float32) if k1.get("uright") is None else np.asarray(k1["uright"], np.float32)
    u2 = np.full(len(k2["kps"]), -1, np.float32) if k2.get("uright") is None else np.asarray(k2["uright"], np.float32)
    nodes2 = {int(nd): j for j, nd in enumerate(k2["fv"]["fv_node"])}
    off1, off2 = k1["fv"]["fv_off"], k2["fv"]["fv_off"]
    hist = [0] * HISTO
    bins = {}
    nm = 0
    for j1, nd in enumerate(k1["fv"]["fv_node"]):
        j2 = nodes2.get(int(nd))
        if j2 is None:
            continue
        for t in range(off1[j1], off1[j1 + 1]):
            idx1 = int(k1["fv"]["fv_feat"][t])
            if k1["has_mp"][idx1]:
                continue
            st1 = u1[idx1] >= 0
            if only_stereo and not st1:
                continue
            p1 = k1["kps"][idx1]
            a = f32(f32(f32(p1["x"] * F[0, 0]) + f32(p1["y"] * F[1, 0])) + F[2, 0])
            b = f32(f32(f32(p1["x"] * F[0, 1]) + f32(p1["y"] * F[1, 1])) + F[2, 1])
            c = f32(f32(f32(p1["x"] * F[0, 2]) + f32(p1["y"] * F[1, 2])) + F[2, 2])
            den = f32(f32(a * a) + f32(b * b))
            best, bi2 = TH_LOW, -1
            for s in range(off2[j2], off2[j2 + 1]):
                idx2 = int(k2["fv"]["fv_feat"][s])
                if k2["has_mp"][idx2]:
                    continue
                st2 = u2[idx2] >= 0
                if only_stereo and not st2:
                    continue
                d = hamming(k1["desc"][idx1], k2["desc"][idx2])
                if d > TH_LOW or d > best:
                    continue
                p2 = k2["kps"][idx2]
                if not st1 and not st2:
                    dx, dy = f32(f32(ep[0]) - p2["x"]), f32(f32(ep[1]) - p2["y"])
                    if f32(f32(dx * dx) + f32(dy * dy)) < f32(f32(100) * f32(scale[int(p2["octave"])])):
                        continue
                ok = bool(coarse)
                if not ok:
                    num = f32(f32(f32(a * p2["x"]) + f32(b * p2["y"])) + c)
                    if den != 0:
                        dsqr = f32(f32(num * num) / den)
                        ok = float(dsqr) < 3.84 * float(f32(sigma2[int(p2["octave"])]))
                if ok:
                    bi2, best = idx2, d
            if bi2 >= 0:
                m12[idx1] = bi2
                nm += 1
                if check_orientation:
                    bn = rot_bin(p1["angle"], k2["kps"][bi2]["angle"])
                    hist[bn] += 1
                    bins[idx1] = bn
    if check_orientation:
        keep = three_maxima(hist)
        for idx1, bn in bins.items():
            if bn not in keep:
                m12[idx1] = -1
                nm -= 1
    return nm, m12


# ---- SearchByProjection(KeyFrame, Sim3, vpPoints, vpMatched, th, ratioHamming) (:397-494) ------------------------------------------------------
def search_by_projection_sim3(kps, desc, matched0, scale, gp, q, qdesc, th, ratio):
    """q: Q_DTYPE with u, v, octave = predicted level, flags bit 0 = the candidate reaches the window. Returns (nmatches, match[n])."""
    o = om.oracle()
    n = len(kps)
    locked = np.zeros(n, bool) if matched0 is None else np.asarray(matched0[:n]).astype(bool).copy()
    match = np.full(n, -1, np.int32)
    lim = f32(f32(TH_LOW) * f32(ratio))
    nm = 0
    for i in range(len(q)):
        if not (q[i]["flags"] & 1):
            continue
        u, v = f32(q[i]["u"]), f32(q[i]["v"])
        if not (u >= gp[0] and u < gp[2] and v >= gp[1] and v < gp[3]):
            continue
        lvl = int(q[i]["octave"])
        r = f32(f32(int(th)) * f32(scale[lvl]))
        best, bi = 256, -1
        for idx in o.features_in_area(kps, gp, u, v, r, -1, -1):
            if locked[idx]:
                continue
            kl = int(kps[idx]["octave"])
            if kl < lvl - 1 or kl > lvl:
                continue
            d = hamming(qdesc[i], desc[idx])
            if d < best:
                best, bi = d, int(idx)
        if f32(best) <= lim:
            match[bi] = i
            locked[bi] = True
            nm += 1
    return nm, match


# ---- SearchBySim3 (:1323-1519) = two Fuse-style searches + the agreement test --------------------------------------------------------------
TH_HIGH = 100


def search_by_sim3_compose(best12, dist12, best21, dist21, init12, has1, has2):
    """vpMatches12 of SearchBySim3 from the two searches: best12[i1] / dist12[i1] = best keypoint of pKF2 for the map point of
    keypoint i1 of pKF1 (orb_fuse_search mode 1 on pKF2), best21 / dist21 the other direction. init12 = the matches the call starts
    with. Returns (nFound, match12)."""
    n1, n2 = len(best12), len(best21)
    m1 = np.where((dist12 <= TH_HIGH) & (best12 >= 0), best12, -1)
    m2 = np.where((dist21 <= TH_HIGH) & (best21 >= 0), best21, -1)
    out = np.array(init12, np.int32).copy()
    nf = 0
    for i1 in range(n1):
        j = m1[i1]
        if j >= 0 and m2[j] == i1:
            out[i1] = j
            nf += 1
    return nf, out


def sim3_queries(p, already):
    """orb_fuse_query records for one direction of SearchBySim3: p = Sim3PointC records, already[i] = vbAlreadyMatched"""
    q = np.zeros(len(p), FQ_DTYPE)
    q["u"], q["v"], q["level"] = p["u"], p["v"], p["level"]
    q["flags"] = ((p["flags"] & 1) != 0) & ((p["flags"] & 2) == 0) & ~np.asarray(already, bool)
    return q


# ---- SearchByBoW(KeyFrame, KeyFrame) (:702-819) ---------------------------------------------------------------------------------------
def search_by_bow_kf(k1, k2, nnratio=0.75, check_orientation=True):
    """k1 / k2: dicts kps, desc, has_mp (map point present and not bad), fv. Returns (nmatches, match12[n1])."""
    n1 = len(k1["kps"])
    m12 = np.full(n1, -1, np.int32)
    matched2 = np.zeros(len(k2["kps"]), bool)
    nodes2 = {int(nd): j for j, nd in enumerate(k2["fv"]["fv_node"])}
    off1, off2 = k1["fv"]["fv_off"], k2["fv"]["fv_off"]
    hist = [0] * HISTO
    bins = {}
    nm = 0
    for j1, nd in enumerate(k1["fv"]["fv_node"]):
        j2 = nodes2.get(int(nd))
        if j2 is None:
            continue
        for t in range(off1[j1], off1[j1 + 1]):
            idx1 = int(k1["fv"]["fv_feat"][t])
            if not k1["has_mp"][idx1]:
                continue
            b1, b2, bi = 256, 256, -1
            for s_ in range(off2[j2], off2[j2 + 1]):
                idx2 = int(k2["fv"]["fv_feat"][s_])
                if matched2[idx2] or not k2["has_mp"][idx2]:
                    continue
                d = hamming(k1["desc"][idx1], k2["desc"][idx2])
                if d < b1:
                    b2, b1, bi = b1, d, idx2
                elif d < b2:
                    b2 = d
            if b1 < TH_LOW and f32(b1) < f32(f32(nnratio) * f32(b2)):
                m12[idx1] = bi
                matched2[bi] = True
                nm += 1
                if check_orientation:
                    bn = rot_bin(k1["kps"][idx1]["angle"], k2["kps"][bi]["angle"])
                    hist[bn] += 1
                    bins[idx1] = bn
    if check_orientation:
        keep = three_maxima(hist)
        for idx1, bn in bins.items():
            if bn not in keep:
                m12[idx1] = -1
                nm -= 1
    return nm, m12


# ---- ComputeDistinctiveDescriptors ----------------------------------------------------------------------------------------------------
def distinctive(desc):
    """(BestIdx, BestMedian) of one map point's observed descriptors [N, 32]; (-1, -1) for none"""
    N = len(desc)
    if N == 0:
        return -1, -1
    D = _POP[np.bitwise_xor(desc[:, None, :], desc[None, :, :])].sum(2)
    k = int(0.5 * (N - 1))
    best, bidx = 2 ** 31 - 1, 0
    for i in range(N):
        med = int(np.sort(D[i])[k])
        if med < best:
            best, bidx = med, i
    return bidx, best


# ---- the reference's own lines ---------------------------------------------------------------------------------------------------
# ---- SearchForTriangulation between two-camera keyframes (mpCamera2 != NULL) ------------------------------------------------------------
REF_SFT2_SO = os.path.join(HERE, "_ref", "libmorb_ref_sft2.so")


def search_for_triangulation_fisheye(k1, k2, sigma2, rigs, only_stereo=False, coarse=False, check_orientation=True, tri=None):
    """reference src/ORBmatcher.cc:821-1042 with pKF1->mpCamera2 && pKF2->mpCamera2. k = dict kps (left then right), desc, has_mp, fv,
    nleft; rigs = synth.RIG_DTYPE[4] (ll, lr, rl, rr). tri: the TriangulateMatches checker (oracle_kb8_py.oracle() by default; its
    reference() gives the reference's own lines). Returns (nmatches, match12[n1])."""
    from oracle import oracle_kb8_py as ok
    tri = ok.oracle() if tri is None else tri
    n1 = len(k1["kps"])
    m12 = np.full(n1, -1, np.int32)
    if only_stereo:                      # bStereo1 is false with a second camera: every idx1 is skipped (:887-890)
        return 0, m12
    nodes2 = {int(nd): j for j, nd in enumerate(k2["fv"]["fv_node"])}
    off1, off2 = k1["fv"]["fv_off"], k2["fv"]["fv_off"]
    rigd = [dict(cam1=r["cam1"], cam2=r["cam2"], prec1=r["prec1"], prec2=r["prec2"], R12=r["R12"].reshape(3, 3), t12=r["t12"]) for r in rigs]
    hist = [0] * HISTO
    bins = {}
    nm = 0
    for j1, nd in enumerate(k1["fv"]["fv_node"]):
        j2 = nodes2.get(int(nd))
        if j2 is None:
            continue
        cand2 = [int(x) for x in k2["fv"]["fv_feat"][off2[j2]:off2[j2 + 1]] if not k2["has_mp"][int(x)]]
        for t in range(off1[j1], off1[j1 + 1]):
            idx1 = int(k1["fv"]["fv_feat"][t])
            if k1["has_mp"][idx1]:
                continue
            p1 = k1["kps"][idx1]
            right1 = idx1 >= k1["nleft"]
            dist = [hamming(k1["desc"][idx1], k2["desc"][i2]) for i2 in cand2]
            ok_tri = {}
            if not coarse:               # the constraint of every candidate that can reach it, one batched call per camera combination
                for r2 in (0, 1):
                    sel = [i2 for i2, d in zip(cand2, dist) if d <= TH_LOW and (i2 >= k2["nleft"]) == bool(r2)]
                    if not sel:
                        continue
                    p2 = k2["kps"][sel]
                    xy1 = np.tile(np.array([[p1["x"], p1["y"]]], np.float32), (len(sel), 1))
                    xy2 = np.stack([p2["x"], p2["y"]], 1).astype(np.float32)
                    s1 = np.full(len(sel), sigma2[int(p1["octave"])], np.float32)
                    s2 = np.asarray(sigma2, np.float32)[p2["octave"]]
                    ret = tri.triangulate(rigd[2 * int(right1) + r2], xy1, xy2, s1, s2)[0]
                    for i2, rv in zip(sel, ret):
                        ok_tri[i2] = bool(rv > f32(0.0001))
            best, bi2 = TH_LOW, -1
            for i2, d in zip(cand2, dist):
                if d > TH_LOW or d > best:
                    continue
                if coarse or ok_tri[i2]:
                    bi2, best = i2, d
            if bi2 >= 0:
                m12[idx1] = bi2
                nm += 1
                if check_orientation:
                    bn = rot_bin(p1["angle"], k2["kps"][bi2]["angle"])
                    hist[bn] += 1
                    bins[idx1] = bn
    if check_orientation:
        keep = three_maxima(hist)
        for idx1, bn in bins.items():
            if bn not in keep:
                m12[idx1] = -1
                nm -= 1
    return nm, m12


def ref_search_for_triangulation_fisheye(k1, k2, scale, sigma2, rigs, only_stereo=False, coarse=False, check_orientation=True):
    """the reference's own lines (oracle/_ref/libmorb_ref_sft2.so)"""
    L = _Lib.load(REF_SFT2_SO)
    vp, i = C.c_void_p, C.c_int
    L.refsft2_search.argtypes = [vp, vp, vp, i, i, vp, vp, vp, i, vp, vp, vp, i, i, vp, vp, vp, i, vp, vp, i, vp, i, i, i, vp]

    def arrs(k):
        kps = np.ascontiguousarray(k["kps"], KP_DTYPE); d = np.ascontiguousarray(k["desc"], np.uint8)
        hm = np.ascontiguousarray(k["has_mp"], np.uint8)
        fv = [np.ascontiguousarray(k["fv"][n], t) for n, t in (("fv_node", np.uint32), ("fv_off", np.int32), ("fv_feat", np.uint32))]
        return kps, d, hm, fv
    a, b = arrs(k1), arrs(k2)
    scale = np.ascontiguousarray(scale, np.float32); sigma2 = np.ascontiguousarray(sigma2, np.float32)
    rigs = np.ascontiguousarray(rigs)
    assert rigs.dtype.itemsize == 120 and len(rigs) == 4
    out = np.full(max(len(a[0]), 1), -1, np.int32)
    nm = L.refsft2_search(_p(a[0]), _p(a[1]), _p(a[2]), len(a[0]), int(k1["nleft"]), _p(a[3][0]), _p(a[3][1]), _p(a[3][2]), len(a[3][0]),
                          _p(b[0]), _p(b[1]), _p(b[2]), len(b[0]), int(k2["nleft"]), _p(b[3][0]), _p(b[3][1]), _p(b[3][2]), len(b[3][0]),
                          _p(scale), _p(sigma2), len(scale), _p(rigs), int(only_stereo), int(coarse), int(check_orientation), _p(out))
    return nm, out[:len(a[0])]


def have_reference_sft2():
    return os.path.exists(REF_SFT2_SO)


# ---- SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) ----------------------------------------------------------
IQ_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("angle", "<f4"), ("octave", "<i4")])   # orb_init_query
INT_MAX = 2147483647


def search_for_initialization(k1, d1, prev, k2, d2, gp, window, nnratio=0.9, check_orientation=True):
    """reference src/ORBmatcher.cc:603-700. prev: float32 [n1, 2] = vbPrevMatched. Returns (nmatches, vnMatches12, vbPrevMatched)."""
    o = om.oracle()
    n1, n2 = len(k1), len(k2)
    m12 = np.full(n1, -1, np.int32); m21 = np.full(n2, -1, np.int32)
    md = np.full(n2, INT_MAX, np.int64)
    prev = np.array(prev, np.float32).reshape(n1, 2).copy()
    hist = [[] for _ in range(HISTO)]
    nm = 0
    for i1 in range(n1):
        if int(k1[i1]["octave"]) > 0:
            continue
        ind = o.features_in_area(k2, gp, prev[i1, 0], prev[i1, 1], float(window), 0, 0)
        if len(ind) == 0:
            continue
        best, best2, bi2 = INT_MAX, INT_MAX, -1
        for i2 in ind:
            d = hamming(d1[i1], d2[i2])
            if md[i2] <= d:
                continue
            if d < best:
                best2, best, bi2 = best, d, int(i2)
            elif d < best2:
                best2 = d
        if best <= TH_LOW and f32(best) < f32(f32(best2) * f32(nnratio)):
            if m21[bi2] >= 0:
                m12[m21[bi2]] = -1
                nm -= 1
            m12[i1] = bi2; m21[bi2] = i1; md[bi2] = best
            nm += 1
            if check_orientation:
                hist[rot_bin(k1[i1]["angle"], k2[bi2]["angle"])].append(i1)
    if check_orientation:
        keep = three_maxima([len(b) for b in hist])
        for b in range(HISTO):
            if b in keep:
                continue
            for i1 in hist[b]:
                if m12[i1] >= 0:
                    m12[i1] = -1
                    nm -= 1
    for i1 in range(n1):
        if m12[i1] >= 0:
            prev[i1, 0], prev[i1, 1] = k2[m12[i1]]["x"], k2[m12[i1]]["y"]
    return nm, m12, prev


def ref_search_for_initialization(k1, d1, prev, k2, d2, gp, window, nnratio=0.9, check_orientation=True):
    """the reference's own lines (oracle/_ref/libmorb_ref_match.so: refm_search_for_initialization)"""
    L = _Lib.load(om.REF_MATCH_SO)
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    L.refm_search_for_initialization.argtypes = [vp, vp, i, vp, vp, i, vp, vp, i, f, i, vp]
    k1 = np.ascontiguousarray(k1, KP_DTYPE); d1 = np.ascontiguousarray(d1, np.uint8)
    k2 = np.ascontiguousarray(k2, KP_DTYPE); d2 = np.ascontiguousarray(d2, np.uint8)
    pv = np.array(prev, np.float32).reshape(len(k1), 2).copy()
    out = np.full(max(len(k1), 1), -1, np.int32)
    nm = L.refm_search_for_initialization(_p(k1), _p(d1), len(k1), _p(k2), _p(d2), len(k2), _p(gp), _p(pv), int(window), float(nnratio),
                                          int(check_orientation), _p(out))
    return nm, out[:len(k1)], pv


class _Ref:
    def __init__(self):
        self.lib = _Lib.load(REF_MAP_SO)
        L = self.lib
        vp, i, f = C.c_void_p, C.c_int, C.c_float
        L.refmap_fuse.argtypes = [vp, vp, vp, i, vp, vp, vp, i, f, vp, vp, vp, vp, i, f, i, vp, i, vp, vp, vp, vp, vp]
        L.refmap_features_in_area.argtypes = [vp, i, vp, f, f, f, vp, i]
        L.refmap_fuse_right.argtypes = [vp, vp, i, vp, vp, i, vp, vp, vp, i, f, vp, vp, vp, vp, i, f, vp, i, vp, vp, vp, vp, vp]
        L.refmap_search_for_triangulation.argtypes = [vp, vp, vp, vp, i, vp, vp, vp, i, vp, vp, vp, vp, i, vp, vp, vp, i, vp, vp, vp, i, vp, f, f,
                                                      i, i, i, vp]
        L.refmap_search_by_projection_kf.argtypes = [vp, vp, vp, i, vp, i, vp, vp, vp, i, f, i, i, vp]
        L.refmap_distinctive.argtypes = [vp, i]
        L.refmap_search_by_projection_sim3.argtypes = [vp, vp, vp, i, vp, vp, vp, i, vp, vp, vp, i, i, f, vp]
        L.refmap_search_by_sim3.argtypes = [vp, vp, i, vp, vp, vp, vp, i, vp, vp, vp, vp, vp, i, vp, f, vp]
        L.refmap_search_by_bow_kf.argtypes = [vp, vp, vp, i, vp, vp, vp, i, vp, vp, vp, i, vp, vp, vp, i, vp, f, i, vp]

    def features_in_area(self, kps, gp, x, y, r):
        kps = np.ascontiguousarray(kps, dtype=KP_DTYPE)
        out = np.zeros(max(len(kps), 1), np.int32)
        n = self.lib.refmap_features_in_area(_p(kps), len(kps), _p(gp), float(x), float(y), float(r), _p(out), len(out))
        assert n >= 0
        return out[:n]

    def fuse(self, kps, desc, uright, gp, scale, sigma2, bf, kf_mp_nobs, kf_mp_bad, pts, pdesc, th, sim3=False):
        kps = np.ascontiguousarray(kps, dtype=KP_DTYPE); desc = np.ascontiguousarray(desc, np.uint8)
        ur = None if uright is None else np.ascontiguousarray(uright, np.float32)
        scale = np.ascontiguousarray(scale, np.float32); sigma2 = np.ascontiguousarray(sigma2, np.float32)
        kf_mp_nobs = np.ascontiguousarray(kf_mp_nobs, np.int32); kf_mp_bad = np.ascontiguousarray(kf_mp_bad, np.uint8)
        pts = np.ascontiguousarray(pts, FP_DTYPE); pdesc = np.ascontiguousarray(pdesc, np.uint8)
        n, nq = len(kps), len(pts)
        ev = np.zeros(3 * (2 * nq + 8), np.int32); nev = C.c_int(0)
        repl = np.zeros(max(nq, 1), np.int32); kf_final = np.zeros(max(n, 1), np.int32)
        cb = np.zeros(max(nq, 1), np.int32); cn = np.zeros(max(nq, 1), np.int32)
        nf = self.lib.refmap_fuse(_p(kps), _p(desc), _p(ur) if ur is not None else None, n, _p(gp), _p(scale), _p(sigma2), len(scale), float(bf),
                                  _p(kf_mp_nobs), _p(kf_mp_bad), _p(pts), _p(pdesc), nq, float(th), int(sim3), _p(ev), len(ev),
                                  C.byref(nev), _p(repl), _p(kf_final), _p(cb), _p(cn))
        assert nf > -1000
        events = [tuple(int(x) for x in ev[3 * k:3 * k + 3]) for k in range(nev.value)]
        return nf, events, repl[:nq], kf_final[:n], cb[:nq], cn[:nq]

    def fuse_right(self, kpsL, descL, kpsR, descR, gp, scale, sigma2, bf, kf_mp_nobs, kf_mp_bad, pts, pdesc, th):
        """ORBmatcher::Fuse(pKF, vpMapPoints, th, bRight = true) on a two-camera keyframe; kf_mp_* over [0, nL + nR)"""
        kpsL = np.ascontiguousarray(kpsL, dtype=KP_DTYPE); descL = np.ascontiguousarray(descL, np.uint8)
        kpsR = np.ascontiguousarray(kpsR, dtype=KP_DTYPE); descR = np.ascontiguousarray(descR, np.uint8)
        scale = np.ascontiguousarray(scale, np.float32); sigma2 = np.ascontiguousarray(sigma2, np.float32)
        kf_mp_nobs = np.ascontiguousarray(kf_mp_nobs, np.int32); kf_mp_bad = np.ascontiguousarray(kf_mp_bad, np.uint8)
        pts = np.ascontiguousarray(pts, FP_DTYPE); pdesc = np.ascontiguousarray(pdesc, np.uint8)
        n, nq = len(kpsL) + len(kpsR), len(pts)
        assert len(kf_mp_nobs) == n
        ev = np.zeros(3 * (2 * nq + 8), np.int32); nev = C.c_int(0)
        repl = np.zeros(max(nq, 1), np.int32); kf_final = np.zeros(max(n, 1), np.int32)
        cb = np.zeros(max(nq, 1), np.int32); cn = np.zeros(max(nq, 1), np.int32)
        nf = self.lib.refmap_fuse_right(_p(kpsL), _p(descL), len(kpsL), _p(kpsR), _p(descR), len(kpsR), _p(gp), _p(scale), _p(sigma2),
                                        len(scale), float(bf), _p(kf_mp_nobs), _p(kf_mp_bad), _p(pts), _p(pdesc), nq, float(th), _p(ev),
                                        len(ev), C.byref(nev), _p(repl), _p(kf_final), _p(cb), _p(cn))
        assert nf > -1000
        events = [tuple(int(x) for x in ev[3 * k:3 * k + 3]) for k in range(nev.value)]
        return nf, events, repl[:nq], kf_final[:n], cb[:nq], cn[:nq]

    def search_for_triangulation(self, k1, k2, gp, scale, sigma2, F12, ep, only_stereo=False, coarse=False, check_orientation=True):
        def arrs(k):
            kps = np.ascontiguousarray(k["kps"], KP_DTYPE); d = np.ascontiguousarray(k["desc"], np.uint8)
            ur = None if k.get("uright") is None else np.ascontiguousarray(k["uright"], np.float32)
            hm = np.ascontiguousarray(k["has_mp"], np.uint8)
            fv = [np.ascontiguousarray(k["fv"][n], t) for n, t in (("fv_node", np.uint32), ("fv_off", np.int32), ("fv_feat", np.uint32))]
            return kps, d, ur, hm, fv
        a, b = arrs(k1), arrs(k2)
        scale = np.ascontiguousarray(scale, np.float32); sigma2 = np.ascontiguousarray(sigma2, np.float32)
        F = np.ascontiguousarray(F12, np.float32).reshape(9)
        out = np.full(max(len(a[0]), 1), -1, np.int32)
        nm = self.lib.refmap_search_for_triangulation(
            _p(a[0]), _p(a[1]), _p(a[2]) if a[2] is not None else None, _p(a[3]), len(a[0]), _p(a[4][0]), _p(a[4][1]), _p(a[4][2]), len(a[4][0]),
            _p(b[0]), _p(b[1]), _p(b[2]) if b[2] is not None else None, _p(b[3]), len(b[0]), _p(b[4][0]), _p(b[4][1]), _p(b[4][2]), len(b[4][0]),
            _p(gp), _p(scale), _p(sigma2), len(scale), _p(F), float(ep[0]), float(ep[1]), int(only_stereo), int(coarse), int(check_orientation),
            _p(out))
        return nm, out[:len(a[0])]

    def search_by_projection_kf(self, kps, desc, locked0, scale, gp, q, qdesc, th, orb_dist, check_orientation=True):
        kps = np.ascontiguousarray(kps, KP_DTYPE); desc = np.ascontiguousarray(desc, np.uint8)
        lk = None if locked0 is None else np.ascontiguousarray(locked0, np.uint8)
        scale = np.ascontiguousarray(scale, np.float32)
        q = np.ascontiguousarray(q, Q_DTYPE); qdesc = np.ascontiguousarray(qdesc, np.uint8)
        out = np.full(max(len(kps), 1), -1, np.int32)
        nm = self.lib.refmap_search_by_projection_kf(_p(kps), _p(desc), _p(lk) if lk is not None else None, len(kps), _p(scale), len(scale), _p(gp),
                                                     _p(q), _p(qdesc), len(q), float(th), int(orb_dist), int(check_orientation), _p(out))
        return nm, out[:len(kps)]

    def distinctive(self, desc):
        desc = np.ascontiguousarray(desc, np.uint8)
        return int(self.lib.refmap_distinctive(_p(desc), len(desc)))

    def search_by_projection_sim3(self, kps, desc, matched0, gp, scale, sigma2, q, qdesc, found_slot, th, ratio):
        kps = np.ascontiguousarray(kps, KP_DTYPE); desc = np.ascontiguousarray(desc, np.uint8)
        m0 = np.ascontiguousarray(matched0, np.uint8)
        scale = np.ascontiguousarray(scale, np.float32); sigma2 = np.ascontiguousarray(sigma2, np.float32)
        q = np.ascontiguousarray(q, Q_DTYPE); qdesc = np.ascontiguousarray(qdesc, np.uint8)
        fs = np.ascontiguousarray(found_slot, np.int32)
        out = np.full(max(len(kps), 1), -1, np.int32)
        nm = self.lib.refmap_search_by_projection_sim3(_p(kps), _p(desc), _p(m0), len(kps), _p(gp), _p(scale), _p(sigma2), len(scale), _p(q),
                                                       _p(qdesc), _p(fs), len(q), int(th), float(ratio), _p(out))
        return nm, out[:len(kps)]

    def search_by_sim3(self, kps1, desc1, p1, pdesc1, kps2, desc2, p2, pdesc2, gp, scale, sigma2, init12, th):
        kps1 = np.ascontiguousarray(kps1, KP_DTYPE); desc1 = np.ascontiguousarray(desc1, np.uint8)
        kps2 = np.ascontiguousarray(kps2, KP_DTYPE); desc2 = np.ascontiguousarray(desc2, np.uint8)
        p1 = np.ascontiguousarray(p1, S3_DTYPE); p2 = np.ascontiguousarray(p2, S3_DTYPE)
        pdesc1 = np.ascontiguousarray(pdesc1, np.uint8); pdesc2 = np.ascontiguousarray(pdesc2, np.uint8)
        scale = np.ascontiguousarray(scale, np.float32); sigma2 = np.ascontiguousarray(sigma2, np.float32)
        init12 = np.ascontiguousarray(init12, np.int32)
        out = np.full(max(len(kps1), 1), -1, np.int32)
        nf = self.lib.refmap_search_by_sim3(_p(kps1), _p(desc1), len(kps1), _p(p1), _p(pdesc1), _p(kps2), _p(desc2), len(kps2), _p(p2), _p(pdesc2),
                                            _p(gp), _p(scale), _p(sigma2), len(scale), _p(init12), float(th), _p(out))
        return nf, out[:len(kps1)]

    def search_by_bow_kf(self, k1, k2, gp, nnratio=0.75, check_orientation=True):
        def arrs(k):
            kps = np.ascontiguousarray(k["kps"], KP_DTYPE); d = np.ascontiguousarray(k["desc"], np.uint8)
            st = np.ascontiguousarray(k["mp_state"], np.uint8)
            fv = [np.ascontiguousarray(k["fv"][n], t) for n, t in (("fv_node", np.uint32), ("fv_off", np.int32), ("fv_feat", np.uint32))]
            return kps, d, st, fv
        a, b = arrs(k1), arrs(k2)
        out = np.full(max(len(a[0]), 1), -1, np.int32)
        nm = self.lib.refmap_search_by_bow_kf(_p(a[0]), _p(a[1]), _p(a[2]), len(a[0]), _p(a[3][0]), _p(a[3][1]), _p(a[3][2]), len(a[3][0]),
                                              _p(b[0]), _p(b[1]), _p(b[2]), len(b[0]), _p(b[3][0]), _p(b[3][1]), _p(b[3][2]), len(b[3][0]),
                                              _p(gp), float(nnratio), int(check_orientation), _p(out))
        return nm, out[:len(a[0])]


def reference():
    return _Ref()


def have_reference():
    return os.path.exists(REF_MAP_SO)
