"""Deterministic synthetic frames for parity tests and benchmarks.

Mid-grey background plus ~W*H/900 random rotated filled rectangles, a small Gaussian blur and
additive noise. A stereo pair renders the same rectangles shifted by a disparity that grows towards
the bottom of the image (ground-plane-like), so rows align and rectified-stereo matches exist. NumPy only (no cv2, no torch).
"""
import numpy as np

CONFIGS = {
    # name: (width, height, nfeatures, lapping area, fx, baseline)
    "euroc_mono": (752, 480, 1000, (0, 1000), 435.2, 0.11008),
    "euroc": (752, 480, 1200, (0, 0), 435.2, 0.11008),
    "tumvi": (512, 512, 1500, (0, 511), 190.98, 0.101),
    "kitti": (1241, 376, 2000, (0, 0), 718.856, 0.53716),
}


def _blur_s08(img):
    k = np.exp(-0.5 * (np.arange(-2, 3) / 0.8) ** 2)
    k /= k.sum()
    p = np.pad(img, ((0, 0), (2, 2)), mode="reflect")
    h = sum(k[i] * p[:, i:i + img.shape[1]] for i in range(5))
    p = np.pad(h, ((2, 2), (0, 0)), mode="reflect")
    return sum(k[i] * p[i:i + img.shape[0], :] for i in range(5))


def _scene(rng, w, h):
    n = max(8, (w * h) // 900)
    cx = rng.uniform(-20, w + 20, n)
    cy = rng.uniform(-20, h + 20, n)
    sa = rng.uniform(6, 60, n)
    sb = rng.uniform(6, 60, n)
    th = rng.uniform(0, np.pi, n)
    g = rng.uniform(20, 235, n)
    # ground-plane-like disparity: grows towards the bottom of the image, small per-rectangle jitter
    d = 4.0 + 40.0 * np.clip(cy, 0, h) / h + rng.uniform(-0.5, 0.5, n)
    return cx, cy, sa, sb, th, g, d


def _render(scene, w, h, shift, noise_rng):
    cx, cy, sa, sb, th, g, d = scene
    img = np.full((h, w), 128.0, np.float32)
    for i in range(len(cx)):
        x0c = cx[i] - shift * d[i]
        r = 0.5 * np.hypot(sa[i], sb[i]) + 1
        xa, xb = int(max(0, np.floor(x0c - r))), int(min(w, np.ceil(x0c + r) + 1))
        ya, yb = int(max(0, np.floor(cy[i] - r))), int(min(h, np.ceil(cy[i] + r) + 1))
        if xa >= xb or ya >= yb:
            continue
        yy, xx = np.mgrid[ya:yb, xa:xb]
        c, s = np.cos(th[i]), np.sin(th[i])
        u = (xx - x0c) * c + (yy - cy[i]) * s
        v = -(xx - x0c) * s + (yy - cy[i]) * c
        m = (np.abs(u) <= 0.5 * sa[i]) & (np.abs(v) <= 0.5 * sb[i])
        img[ya:yb, xa:xb][m] = g[i]
    img = _blur_s08(img)
    img = img + noise_rng.normal(0, 2, img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def mono_frame(seed, w=752, h=480):
    rng = np.random.default_rng(seed)
    scene = _scene(rng, w, h)
    return _render(scene, w, h, 0.0, rng)


def stereo_pair(seed, w=752, h=480):
    """Left/right rectified pair: right image = scene shifted left by each rectangle's disparity."""
    rng = np.random.default_rng(seed)
    scene = _scene(rng, w, h)
    left = _render(scene, w, h, 0.0, rng)
    right = _render(scene, w, h, 1.0, rng)
    return left, right


def flat_frame(seed, w=752, h=480):
    """Almost textureless frame: most cells fall back to minThFAST and K < nFeatures."""
    rng = np.random.default_rng(seed)
    img = 120 + 3 * rng.standard_normal((h, w))
    yy, xx = np.mgrid[0:h, 0:w]
    img += 10 * np.sin(xx / 37.0) * np.cos(yy / 23.0)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def plateau_frame(seed, w=752, h=480):
    """Frame quantised to 4 grey levels: many equal FAST scores (NMS tie handling)."""
    img = mono_frame(seed, w, h)
    return ((img // 64) * 64 + 32).astype(np.uint8)


def random_descriptors(seed, n):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (n, 32), dtype=np.uint8)


def clustered_descriptors(seed, queries, n, max_flips=80):
    """Database rows = random queries with 0..max_flips random bit flips (ties / near duplicates)."""
    rng = np.random.default_rng(seed)
    src = rng.integers(0, len(queries), n)
    db = queries[src].copy()
    bits = np.unpackbits(db, axis=1)
    nflip = rng.integers(0, max_flips + 1, n)
    for i in range(n):
        pos = rng.choice(256, nflip[i], replace=False)
        bits[i, pos] ^= 1
    return np.packbits(bits, axis=1)


# orb_proj_query records (include/orb_b200.h) for the windowed matcher
Q_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("z", "<f4"), ("angle", "<f4"), ("octave", "<i4"), ("flags", "<i4")])


def synth_queries(seed, kps_last, desc_last, kps_cur, desc_cur, w, h, p_valid=0.9, p_obs=0.8, jitter=6.0, p_dup=0.05):
    """Queries as a tracker would produce them: every last-frame keypoint carries a map point that projects near a
    current-frame keypoint with a similar descriptor (here: the nearest current keypoint of a similar octave,
    jittered by a few pixels), plus invalid points, points behind the camera, points outside the image and points
    without observations. A few queries are exact duplicates of the previous one (ties, lock conflicts)."""
    rng = np.random.default_rng(seed)
    n = len(kps_last)
    q = np.zeros(n, Q_DTYPE)
    q["u"] = kps_last["x"] + rng.normal(0, jitter, n).astype(np.float32)
    q["v"] = kps_last["y"] + rng.normal(0, jitter, n).astype(np.float32)
    q["z"] = rng.uniform(0.5, 30.0, n).astype(np.float32)
    q["angle"] = kps_last["angle"]
    q["octave"] = kps_last["octave"]
    flags = (rng.random(n) < p_valid).astype(np.int32) | ((rng.random(n) < p_obs).astype(np.int32) << 1)
    q["flags"] = flags
    behind = rng.random(n) < 0.02
    q["z"][behind] = -q["z"][behind]
    outside = rng.random(n) < 0.02
    q["u"][outside] = np.float32(w + 5)
    qdesc = np.array(desc_last, dtype=np.uint8, copy=True)
    dup = np.nonzero(rng.random(n) < p_dup)[0]
    dup = dup[dup > 0]
    for i in dup:            # same projection and descriptor as the previous query: they compete for one keypoint
        q[i] = q[i - 1]
        q["flags"][i] = flags[i] | 1
        qdesc[i] = qdesc[i - 1]
    return q, qdesc


# orb_track_query records (include/orb_b200.h) for the local-map search
TQ_DTYPE = np.dtype([("proj_x", "<f4"), ("proj_y", "<f4"), ("proj_xr", "<f4"), ("view_cos", "<f4"), ("level", "<i4"), ("flags", "<i4")])


def synth_track_queries(seed, kps_map, desc_map, uright_map, w, h, n_extra=0.5, p_view=0.9, p_obs=0.9, jitter=3.0, p_dup=0.03, mbf=40.0,
                        max_flips=40):
    """Local-map points as Frame::isInFrustum leaves them (src/Frame.cc:561-...): every keypoint of the searched frame is
    seen by a map point that projects near it (predicted level = its octave or one above, descriptor = the keypoint's
    with 0..max_flips bit flips, right-image projection near the keypoint's uRight when it has one), plus n_extra * N
    points re-drawn from the same set with a larger offset (wrong associations competing for the same keypoints),
    points out of view and points without observations."""
    rng = np.random.default_rng(seed)
    n0 = len(kps_map)
    sel = np.concatenate([np.arange(n0), rng.integers(0, max(n0, 1), int(n_extra * n0))]) if n0 else np.zeros(0, np.int64)
    n = len(sel)
    q = np.zeros(n, TQ_DTYPE)
    jit = np.where(np.arange(n) < n0, jitter, 4 * jitter).astype(np.float32)
    dx = (rng.normal(0, 1, n) * jit).astype(np.float32)
    q["proj_x"] = kps_map["x"][sel] + dx
    q["proj_y"] = kps_map["y"][sel] + (rng.normal(0, 1, n) * jit).astype(np.float32)
    depth = rng.uniform(1.0, 40.0, n).astype(np.float32)
    ur = np.asarray(uright_map, np.float32)[sel] if n else np.zeros(0, np.float32)
    q["proj_xr"] = np.where(ur > 0, ur + dx + rng.normal(0, 1, n).astype(np.float32), q["proj_x"] - np.float32(mbf) / depth)
    q["view_cos"] = rng.choice(np.array([0.9995, 0.99, 0.7], np.float32), n)
    q["level"] = np.clip(kps_map["octave"][sel] + rng.integers(0, 2, n), 0, 7)
    q["flags"] = (rng.random(n) < p_view).astype(np.int32) | ((rng.random(n) < p_obs).astype(np.int32) << 1)
    qdesc = np.array(desc_map[sel], dtype=np.uint8, copy=True)
    if n:
        bits = np.unpackbits(qdesc, axis=1)
        nflip = rng.integers(0, max_flips + 1, n)
        flip = rng.random((n, 256)).argsort(axis=1) < nflip[:, None]
        qdesc = np.packbits(bits ^ flip.astype(np.uint8), axis=1)
    dup = np.nonzero(rng.random(n) < p_dup)[0]
    for i in dup[dup > 0]:
        q[i] = q[i - 1]
        qdesc[i] = qdesc[i - 1]
    return q, qdesc


# ---- synthetic DBoW2 vocabulary (the reference's ORBvoc.txt is absent from the mount: .MISSING_LARGE_BLOBS) ----
def synth_vocabulary(seed, k=10, L=6, p_early_leaf=0.0, p_short=0.0, p_stop=0.02, scoring=0, weighting=0):
    """Vocabulary tree in the order DBoW2 creates / saves it (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h HKmeansStep:
    the children of a node get consecutive ids, then each child is expanded): node 0 = root, every other node has a
    parent, a 256-bit descriptor (the parent's with random bit flips, fewer the deeper) and a weight (idf for leaves,
    0 for inner nodes, 0 for stopped words). p_early_leaf: share of nodes above level L that stay leaves (k-means
    clusters with one member); p_short: share of inner nodes with fewer than k children.
    Returns dict(k, L, scoring, weighting, parent[int32 n], is_leaf[uint8 n], desc[uint8 n x 32], weight[float64 n])."""
    rng = np.random.default_rng(seed)
    parent, leaf, desc, weight, level = [0], [0], [np.zeros(32, np.uint8)], [0.0], [0]
    stack = [0]
    while stack:
        p = stack.pop()
        lv = level[p] + 1
        nch = k if rng.random() >= p_short else int(rng.integers(1, k + 1))
        first = len(parent)
        for _ in range(nch):
            if p == 0:
                d = rng.integers(0, 256, 32, dtype=np.uint8)
            else:
                bits = np.unpackbits(desc[p])
                flip = rng.choice(256, max(4, 96 >> lv), replace=False)
                bits[flip] ^= 1
                d = np.packbits(bits)
            is_leaf = lv == L or (lv >= 2 and rng.random() < p_early_leaf)
            parent.append(p); desc.append(d); level.append(lv); leaf.append(1 if is_leaf else 0)
            weight.append((0.0 if rng.random() < p_stop else float(rng.uniform(0.5, 12.0))) if is_leaf else 0.0)
        # expand the children in creation order (depth first like the recursion)
        for c in range(first + nch - 1, first - 1, -1):
            if not leaf[c]:
                stack.append(c)
    return dict(k=k, L=L, scoring=scoring, weighting=weighting, parent=np.array(parent, np.int32), is_leaf=np.array(leaf, np.uint8),
                desc=np.stack(desc).astype(np.uint8), weight=np.array(weight, np.float64), level=np.array(level, np.int32))


def write_vocabulary_text(voc, path):
    """The text format TemplatedVocabulary::loadFromTextFile reads (:1338-1426) = ORBvoc.txt: header 'k L scoring weighting',
    then one line per node except the root: 'parent isLeaf d0 ... d31 weight'."""
    with open(path, "w") as f:
        f.write("%d %d %d %d\n" % (voc["k"], voc["L"], voc["scoring"], voc["weighting"]))
        lines = []
        for i in range(1, len(voc["parent"])):
            lines.append("%d %d %s %s" % (voc["parent"][i], voc["is_leaf"][i], " ".join(str(int(b)) for b in voc["desc"][i]),
                                          repr(float(voc["weight"][i]))))
        f.write("\n".join(lines))   # no trailing newline: the reference's while(!f.eof()) loop would add an empty node for it


def synth_bow_descriptors(seed, voc, n, p_word=0.6, max_flips=12):
    """n descriptors: a share of them are leaf descriptors of the vocabulary with a few bit flips (so that words repeat
    inside one image), the rest uniform random."""
    rng = np.random.default_rng(seed)
    leaves = np.nonzero(voc["is_leaf"])[0]
    pool = rng.choice(leaves, max(1, n // 6))
    out = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    for i in range(n):
        if rng.random() < p_word:
            bits = np.unpackbits(voc["desc"][rng.choice(pool)])
            bits[rng.choice(256, int(rng.integers(0, max_flips + 1)), replace=False)] ^= 1
            out[i] = np.packbits(bits)
    return out


def synth_vocabulary_full(seed, k=10, L=6, p_stop=0.0, scoring=0, weighting=0):
    """Complete k-ary tree of depth L built level by level with numpy (ORBvoc.txt size k = 10, L = 6: 1 111 111 nodes in
    about a second); node ids in breadth-first order, the children of a node consecutive. Same content rules as
    synth_vocabulary."""
    rng = np.random.default_rng(seed)
    descs = [np.zeros((1, 32), np.uint8)]
    parents = [np.zeros(1, np.int32)]
    first = 0
    for lv in range(1, L + 1):
        prev = descs[-1]
        n = len(prev) * k
        if lv == 1:
            d = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        else:
            nflip = max(4, 96 >> lv)
            mask = np.zeros((n, 256), np.uint8)
            cols = rng.integers(0, 256, (n, nflip))
            mask[np.arange(n)[:, None], cols] = 1          # up to nflip distinct bits
            d = np.repeat(prev, k, axis=0) ^ np.packbits(mask, axis=1)
        parents.append((first + np.repeat(np.arange(len(prev)), k)).astype(np.int32))
        first += len(prev)
        descs.append(d)
    desc = np.concatenate(descs)
    parent = np.concatenate(parents)
    nn = len(parent)
    nleaf = len(descs[-1])
    is_leaf = np.zeros(nn, np.uint8)
    is_leaf[nn - nleaf:] = 1
    weight = np.zeros(nn, np.float64)
    w = rng.uniform(0.5, 12.0, nleaf)
    w[rng.random(nleaf) < p_stop] = 0.0
    weight[nn - nleaf:] = w
    return dict(k=k, L=L, scoring=scoring, weighting=weighting, parent=parent, is_leaf=is_leaf, desc=desc, weight=weight)


def rectify_maps(w, h, raw_w=None, raw_h=None, seed=0):
    """Float maps of the kind cv::initUndistortRectifyMap(K, D, R, P, size, CV_32F) yields for a radial-tangential camera
    (src/Settings.cc:540-545), computed here without OpenCV: for every rectified pixel the raw-image position after a
    small rotation, the EuRoC cam0 distortion and intrinsics. Parts of the rectified image fall outside the raw one."""
    raw_w, raw_h = raw_w or w, raw_h or h
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = 458.654 * raw_w / 752, 457.296 * raw_h / 480, 367.215 * raw_w / 752, 248.375 * raw_h / 480
    k1, k2, p1, p2 = -0.28340811, 0.07395907, 0.00019359, 1.76187114e-05
    nfx, nfy, ncx, ncy = 435.2 * w / 752, 435.2 * h / 480, 367.4 * w / 752, 252.2 * h / 480
    rx, ry, rz = rng.uniform(-0.02, 0.02, 3)
    R = np.array([[1, -rz, ry], [rz, 1, -rx], [-ry, rx, 1]])
    v, u = np.mgrid[0:h, 0:w].astype(np.float64)
    X = np.stack([(u - ncx) / nfx, (v - ncy) / nfy, np.ones_like(u)], -1) @ np.linalg.inv(R).T
    x, y = X[..., 0] / X[..., 2], X[..., 1] / X[..., 2]
    r2 = x * x + y * y
    kr = 1 + k1 * r2 + k2 * r2 * r2
    xd = x * kr + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    yd = y * kr + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    return (xd * fx + cx).astype(np.float32), (yd * fy + cy).astype(np.float32)


# ---- fisheye stereo rigs for Frame::ComputeStereoFishEyeMatches / KannalaBrandt8::TriangulateMatches ----
def kb8_rig(kind="tumvi"):
    """dict(cam1, cam2 = KannalaBrandt8::mvParameters fx fy cx cy k0..k3, prec1, prec2 = KannalaBrandt8::precision, R12 = mRlr,
    t12 = mtlr). "tumvi" = the calibration of Examples/Stereo/TUM-VI.yaml (Camera1.*, Camera2.*, Stereo.T_c1_c2);
    "parallel" = two equal cameras side by side without distortion (synth.stereo_pair's geometry: rows align, disparity = f b / z);
    "toed" = strong distortion coefficients and a 10-degree toe-in."""
    if kind == "tumvi":
        cam1 = [190.97847715128717, 190.9733070521226, 254.93170605935475, 256.8974428996504,
                0.0034823894022493434, 0.0007150348452162257, -0.0020532361418706202, 0.00020293673591811182]
        cam2 = [190.44236969414825, 190.4344384721956, 252.59949716835982, 254.91723064636983,
                0.0034003170790442797, 0.001766278153469831, -0.00266312569781606, 0.0003299517423931039]
        R = [[0.999999445773493, 0.000791687752817, 0.000694034010224], [-0.000823363992158, 0.998899461915674, 0.046895490788700],
             [-0.000656143613644, -0.046896036240590, 0.998899560146304]]
        t = [0.101063427414194, 0.001946204678584, 0.001015350132563]
    elif kind == "parallel":
        cam1 = cam2 = [190.0, 190.0, 255.5, 255.5, 0.0, 0.0, 0.0, 0.0]
        R = np.eye(3)
        t = [0.6, 0.0, 0.0]
    elif kind == "toed":
        cam1 = [210.0, 205.0, 250.0, 260.0, -0.03, 0.012, -0.004, 0.0006]
        cam2 = [200.0, 207.0, 262.0, 251.0, 0.02, -0.009, 0.003, -0.0004]
        a = np.deg2rad(10.0)
        R = [[np.cos(a), 0.0, np.sin(a)], [0.0, 1.0, 0.0], [-np.sin(a), 0.0, np.cos(a)]]
        t = [0.25, -0.01, 0.02]
    else:
        raise ValueError(kind)
    return {"cam1": np.asarray(cam1, np.float32), "cam2": np.asarray(cam2, np.float32), "prec1": np.float32(1e-6), "prec2": np.float32(1e-6),
            "R12": np.asarray(R, np.float32).reshape(3, 3), "t12": np.asarray(t, np.float32)}


def _kb8_project64(cam, X):
    """KannalaBrandt8::project in double (generator only)"""
    cam = np.asarray(cam, np.float64)
    r = np.hypot(X[:, 0], X[:, 1])
    th = np.arctan2(r, X[:, 2])
    psi = np.arctan2(X[:, 1], X[:, 0])
    d = th + cam[4] * th ** 3 + cam[5] * th ** 5 + cam[6] * th ** 7 + cam[7] * th ** 9
    return np.stack([cam[0] * d * np.cos(psi) + cam[2], cam[1] * d * np.sin(psi) + cam[3]], 1)


def kb8_pairs(seed, rig, n, w=512, h=512):
    """n keypoint pairs (xy1, xy2, sigma1, sigma2) that exercise every exit of TriangulateMatches: points in front of both cameras
    at 0.3 .. 30 m observed with 0 .. 1.5 px noise (accepted or rejected by the reprojection gates), far points (parallax gate),
    unrelated pixel pairs (negative depths / large errors), the principal point (theta_d <= 1e-8), pixels far outside the image."""
    rng = np.random.default_rng(seed)
    R = np.asarray(rig["R12"], np.float64); t = np.asarray(rig["t12"], np.float64)
    z = np.exp(rng.uniform(np.log(0.3), np.log(30.0), n))
    far = rng.random(n) < 0.15
    z[far] = rng.uniform(40.0, 4000.0, far.sum())
    ang = rng.uniform(0, 2 * np.pi, n); rad = np.tan(rng.uniform(0.0, 1.2, n))
    X1 = np.stack([z * rad * np.cos(ang), z * rad * np.sin(ang), z], 1)
    X2 = (X1 - t) @ R            # x2 = R12^T (x1 - t12)
    noise = rng.uniform(0.0, 1.5, (n, 1)) * rng.standard_normal((n, 2))
    xy1 = _kb8_project64(rig["cam1"], X1) + noise * (rng.random((n, 1)) < 0.7)
    xy2 = _kb8_project64(rig["cam2"], X2) + rng.uniform(0.0, 1.5, (n, 1)) * rng.standard_normal((n, 2)) * (rng.random((n, 1)) < 0.7)
    junk = rng.random(n) < 0.15
    xy2[junk] = rng.uniform(0, [w, h], (junk.sum(), 2))
    xy1 = xy1.astype(np.float32); xy2 = xy2.astype(np.float32)
    if n >= 8:
        xy1[0] = rig["cam1"][2:4]; xy2[0] = rig["cam2"][2:4]           # both principal points
        xy1[1] = rig["cam1"][2:4]                                        # one principal point
        xy1[2] = (-3000.0, 5000.0)                                       # theta_d clamps at pi / 2
        xy2[3] = xy1[3]                                                  # same pixel in both images
    lv = rng.integers(0, 8, (2, n))
    s = (np.float32(1.2) ** np.arange(8, dtype=np.float32)) ** 2
    return xy1, xy2, s[lv[0]].astype(np.float32), s[lv[1]].astype(np.float32)


# ---- two-camera frames (Nleft != -1: the fisheye rig) ---------------------------------------------------------------
# orb_proj_query2 / orb_track_query2 records (include/orb_b200.h)
Q2_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("z", "<f4"), ("angle", "<f4"), ("octave", "<i4"), ("flags", "<i4"), ("ur", "<f4"), ("vr", "<f4")])
TQ2_DTYPE = np.dtype([("proj_x", "<f4"), ("proj_y", "<f4"), ("view_cos", "<f4"), ("level", "<i4"), ("proj_xr", "<f4"), ("proj_yr", "<f4"),
                      ("view_cos_r", "<f4"), ("level_r", "<i4"), ("flags", "<i4"), ("pad", "<i4")])


def synth_queries2(seed, kL, dL, kR, dR, w, h, trl=(-14.25, 0.75), **kw):
    """Frame-to-frame queries against a two-camera frame: map points seen near the left keypoints and map points whose RIGHT
    projection (= left projection + trl, the rig's relative pose reduced to an image shift) falls near the right keypoints.
    Returns (Q_DTYPE queries for the reference driver, Q2_DTYPE queries with (ur, vr) filled in float32, descriptors)."""
    kcat = np.concatenate([kL, kR])
    kcat["x"][len(kL):] = kR["x"] - np.float32(trl[0])
    kcat["y"][len(kL):] = kR["y"] - np.float32(trl[1])
    q, qd = synth_queries(seed, kcat, np.concatenate([dL, dR]), None, None, w, h, **kw)
    order = np.random.default_rng(seed + 1).permutation(len(q))     # interleave the two kinds
    q, qd = q[order], qd[order]
    q2 = np.zeros(len(q), Q2_DTYPE)
    for k in Q_DTYPE.names:
        q2[k] = q[k]
    q2["ur"] = q["u"] + np.float32(trl[0])
    q2["vr"] = q["v"] + np.float32(trl[1])
    return q, q2, qd


def synth_stereo_pairing(seed, nL, nR, p=0.45):
    """mvLeftToRightMatch / mvRightToLeftMatch of a two-camera frame: a random partial one-to-one pairing"""
    rng = np.random.default_rng(seed)
    l2r = np.full(nL, -1, np.int32); r2l = np.full(nR, -1, np.int32)
    m = int(min(nL, nR) * p)
    li = rng.permutation(nL)[:m]; ri = rng.permutation(nR)[:m]
    l2r[li] = ri; r2l[ri] = li
    return l2r, r2l


def synth_track_queries2(seed, kL, dL, kR, dR, l2r, w, h, n_extra=0.5, p_view=0.85, p_obs=0.8, jitter=3.0, p_dup=0.03, max_flips=40):
    """Local-map points seen by a two-camera frame (Frame::isInFrustum fills the left members and, through isInFrustumChecks(...,
    bRight), the ...R members): every left keypoint is seen by a map point that projects near it and, in the right camera, near its
    stereo partner (or anywhere when it has none); more map points come from the right keypoints alone; some are visible in
    one camera only, some have no right level (mnTrackScaleLevelR = -1)."""
    rng = np.random.default_rng(seed)
    nL, nR = len(kL), len(kR)
    srcL = np.concatenate([np.arange(nL), rng.integers(0, max(nL, 1), int(n_extra * nL))]) if nL else np.zeros(0, np.int64)
    srcR = rng.integers(0, max(nR, 1), int(0.5 * nR)) if nR else np.zeros(0, np.int64)
    n = len(srcL) + len(srcR)
    q = np.zeros(n, TQ2_DTYPE)
    jit = np.float32(jitter)
    a = len(srcL)
    part = np.where(l2r[srcL] >= 0, l2r[srcL], rng.integers(0, max(nR, 1), a)) if a and nR else np.zeros(a, np.int64)
    q["proj_x"][:a] = kL["x"][srcL] + (rng.normal(0, 1, a) * jit).astype(np.float32)
    q["proj_y"][:a] = kL["y"][srcL] + (rng.normal(0, 1, a) * jit).astype(np.float32)
    q["level"][:a] = np.clip(kL["octave"][srcL] + rng.integers(0, 2, a), 0, 7)
    if nR:
        q["proj_xr"][:a] = kR["x"][part] + (rng.normal(0, 1, a) * jit).astype(np.float32)
        q["proj_yr"][:a] = kR["y"][part] + (rng.normal(0, 1, a) * jit).astype(np.float32)
        q["level_r"][:a] = np.clip(kR["octave"][part] + rng.integers(0, 2, a), 0, 7)
        q["proj_xr"][a:] = kR["x"][srcR] + (rng.normal(0, 1, n - a) * jit).astype(np.float32)
        q["proj_yr"][a:] = kR["y"][srcR] + (rng.normal(0, 1, n - a) * jit).astype(np.float32)
        q["level_r"][a:] = np.clip(kR["octave"][srcR] + rng.integers(0, 2, n - a), 0, 7)
    q["proj_x"][a:] = rng.uniform(0, w, n - a).astype(np.float32)
    q["proj_y"][a:] = rng.uniform(0, h, n - a).astype(np.float32)
    q["level"][a:] = rng.integers(0, 8, n - a)
    q["view_cos"] = rng.choice(np.array([0.9995, 0.99, 0.7], np.float32), n)
    q["view_cos_r"] = rng.choice(np.array([0.9995, 0.99, 0.7], np.float32), n)
    q["level_r"][rng.random(n) < 0.05] = -1
    inL = rng.random(n) < p_view
    inR = rng.random(n) < p_view
    inL[a:] &= rng.random(n - a) < 0.3            # map points taken from the right keypoints are mostly seen there only
    q["flags"] = inL.astype(np.int32) | ((rng.random(n) < p_obs).astype(np.int32) << 1) | (inR.astype(np.int32) << 2)
    qdesc = np.concatenate([dL[srcL], dR[srcR]]).astype(np.uint8) if n else np.zeros((0, 32), np.uint8)
    if n:
        bits = np.unpackbits(qdesc, axis=1)
        nflip = rng.integers(0, max_flips + 1, n)
        flip = rng.random((n, 256)).argsort(axis=1) < nflip[:, None]
        qdesc = np.packbits(bits ^ flip.astype(np.uint8), axis=1)
    order = rng.permutation(n)
    q, qdesc = q[order], qdesc[order]
    dup = np.nonzero(rng.random(n) < p_dup)[0]
    for i in dup[dup > 0]:
        q[i] = q[i - 1]
        qdesc[i] = qdesc[i - 1]
    return q, qdesc


def synth_bow_keyframe(seed, kL, dL, kR, dR, p_map=0.8, n_each=500, p_flip=0.02):
    """A keyframe for SearchByBoW against a two-camera frame: descriptors of some left and some right keypoints of that frame with a
    few bit flips (so both cameras find matches), angles close to theirs, flags = keypoint holds a good map point."""
    rng = np.random.default_rng(seed)
    iL = rng.permutation(len(kL))[:min(len(kL), n_each)]
    iR = rng.permutation(len(kR))[:min(len(kR), n_each)]
    d = np.concatenate([dL[iL], dR[iR]]).astype(np.uint8)
    a = np.concatenate([kL["angle"][iL], kR["angle"][iR]]).astype(np.float32)
    order = rng.permutation(len(d))
    d, a = d[order], a[order]
    bits = np.unpackbits(d, axis=1)
    d = np.packbits(bits ^ (rng.random(bits.shape) < p_flip).astype(np.uint8), axis=1)
    a = (a + rng.normal(0, 3, len(a)).astype(np.float32)).astype(np.float32)
    return d, a, (rng.random(len(d)) < p_map).astype(np.uint8)


# ---- LocalMapping-side matchers (Fuse, SearchForTriangulation, SearchByProjection(Frame, KeyFrame), ComputeDistinctiveDescriptors) ----
FP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4"), ("min_dist", "<f4"),
                     ("max_dist", "<f4"), ("level", "<i4"), ("nobs", "<i4"), ("flags", "<i4")])


def flip_bits(rng, desc, max_flips):
    """copies of 32-byte descriptors with up to max_flips random bits flipped"""
    out = np.array(desc, dtype=np.uint8, copy=True).reshape(-1, 32)
    for i in range(len(out)):
        k = int(rng.integers(0, max_flips + 1))
        bits = rng.integers(0, 256, k)
        for b in bits:
            out[i, b >> 3] ^= np.uint8(1 << (b & 7))
    return out


def synth_fuse_points(seed, kps, desc, w, h, frac=0.8, jitter=2.0, max_flips=70, nulls=True):
    """Candidate map points of ORBmatcher::Fuse for a keyframe with keypoints kps / descriptors desc, in the stub geometry of
    oracle/ref_driver_map.cc (identity pose and projection: a point at (x, y, z) projects to (x, y)): most fall near a keypoint with a
    similar descriptor and a matching predicted level; a few are NULL, bad, behind the camera, outside the image, out of the
    distance range, seen from behind, or the same MapPoint object twice. Every float test has a wide margin.
    Returns (FP_DTYPE records, descriptors, kf_mp_nobs, kf_mp_bad)."""
    rng = np.random.default_rng(seed)
    n = len(kps)
    m = int(n * frac)
    pick = rng.integers(0, max(n, 1), m)
    p = np.zeros(m, FP_DTYPE)
    p["x"] = kps["x"][pick] + rng.normal(0, jitter, m).astype(np.float32)
    p["y"] = kps["y"][pick] + rng.normal(0, jitter, m).astype(np.float32)
    p["z"] = rng.uniform(1.0, 20.0, m).astype(np.float32)
    d = np.sqrt(p["x"].astype(np.float64) ** 2 + p["y"].astype(np.float64) ** 2 + p["z"].astype(np.float64) ** 2)
    for a, b in (("nx", "x"), ("ny", "y"), ("nz", "z")):
        p[a] = (p[b] / d).astype(np.float32)
    p["min_dist"] = np.float32(0.1)
    p["max_dist"] = np.float32(1e5)
    p["level"] = np.clip(kps["octave"][pick] + rng.choice([0, 0, 0, 1, 1, -1], m), 0, 7)
    p["nobs"] = rng.integers(1, 7, m)
    pdesc = flip_bits(rng, desc[pick], max_flips)
    r = rng.random(m)
    p["z"][r < 0.02] *= -1
    p["x"][(r >= 0.02) & (r < 0.04)] = np.float32(w + 5)
    far = (r >= 0.04) & (r < 0.06)
    p["max_dist"][far] = (0.5 * d[far]).astype(np.float32)
    back = (r >= 0.06) & (r < 0.08)
    for a in ("nx", "ny", "nz"):
        p[a][back] *= -1
    p["flags"][(r >= 0.08) & (r < 0.11)] |= 2
    if nulls:
        p["flags"][(r >= 0.11) & (r < 0.14)] |= 1
    dup = np.nonzero((r >= 0.14) & (r < 0.18))[0]
    for i in dup[dup > 0]:
        keep = p["flags"][i] & 1
        p[i] = p[i - 1]
        p["flags"][i] = (p["flags"][i - 1] & ~1) | 4 | keep
        pdesc[i] = pdesc[i - 1]
    kf_mp_nobs = np.where(rng.random(n) < 0.5, rng.integers(1, 7, n), -1).astype(np.int32)
    kf_mp_bad = ((rng.random(n) < 0.05) & (kf_mp_nobs >= 0)).astype(np.uint8)
    return p, pdesc, kf_mp_nobs, kf_mp_bad


def synth_feature_vector(rng, n, nodes=100, like=None, p_same=0.9):
    """A DBoW2 FeatureVector over n features as CSR in std::map order (ascending node ids, features of a node in ascending index):
    random nodes, or for a second keyframe the node of the corresponding feature (`like` = (node_of_feature_1, correspondence))."""
    node_of = rng.integers(0, nodes, n) * 7 + 3
    if like is not None:
        node1, corr = like
        same = rng.random(n) < p_same
        ok = same & (corr >= 0)
        node_of[ok] = node1[corr[ok]]
    ids = np.unique(node_of)
    off = np.zeros(len(ids) + 1, np.int32)
    feat = []
    for j, nd in enumerate(ids):
        f = np.nonzero(node_of == nd)[0]
        feat.extend(f.tolist())
        off[j + 1] = len(feat)
    return dict(fv_node=ids.astype(np.uint32), fv_off=off, fv_feat=np.array(feat, np.uint32)), node_of


def synth_triangulation_pair(seed, kps, desc, uright, w, h, shift=12.0, jitter=1.0, max_flips=30, p_mp=0.4):
    """Two keyframes for ORBmatcher::SearchForTriangulation: the first is (kps, desc, uright), the second the same keypoints moved by
    `shift` pixels along x with jitter, descriptors with a few flipped bits, in shuffled order. Returns (k1, k2) dicts with kps, desc,
    uright, has_mp, fv."""
    rng = np.random.default_rng(seed)
    n = len(kps)
    perm = rng.permutation(n)
    k2p = np.array(kps[perm], copy=True)
    k2p["x"] += np.float32(shift) + rng.normal(0, jitter, n).astype(np.float32)
    k2p["y"] += rng.normal(0, jitter, n).astype(np.float32)
    k2p["angle"] = np.mod(k2p["angle"] + rng.normal(0, 4, n).astype(np.float32) + np.where(rng.random(n) < 0.1, 90, 0), 360).astype(np.float32)
    d2 = flip_bits(rng, desc[perm], max_flips)
    ur2 = None if uright is None else np.where(rng.random(n) < 0.7, uright[perm], -1).astype(np.float32)
    fv1, node1 = synth_feature_vector(rng, n)
    fv2, _ = synth_feature_vector(rng, n, like=(node1, perm))
    k1 = dict(kps=kps, desc=desc, uright=uright, has_mp=(rng.random(n) < p_mp).astype(np.uint8), fv=fv1)
    k2 = dict(kps=k2p, desc=d2, uright=ur2, has_mp=(rng.random(n) < p_mp).astype(np.uint8), fv=fv2)
    return k1, k2


def synth_fundamental(seed, shift_only=True):
    """A fundamental matrix whose epipolar lines are close to image rows (the second camera moved along x), with small generic terms so
    that every product of Pinhole::epipolarConstrain is exercised; row-major float32 [9]"""
    rng = np.random.default_rng(seed)
    F = np.array([[0, 0, 0], [0, 0, -1], [0, 1, 0]], np.float64)
    F += rng.normal(0, 3e-6, (3, 3))
    F[2, 2] += rng.normal(0, 0.5)
    return (F * 1e-3).astype(np.float32).reshape(9)


def synth_observations(seed, npoints, sizes=(0, 1, 2, 3, 4, 5, 8, 13, 17, 33, 64, 100)):
    """Observed descriptors of map points for MapPoint::ComputeDistinctiveDescriptors: noisy copies of one descriptor per map point
    (so the medians differ), some exact duplicates (ties)."""
    rng = np.random.default_rng(seed)
    out = []
    for p in range(npoints):
        N = int(sizes[p % len(sizes)]) if p < 2 * len(sizes) else int(rng.integers(1, 40))
        base = rng.integers(0, 256, (1, 32)).astype(np.uint8)
        d = flip_bits(rng, np.repeat(base, N, 0), 90) if N else np.zeros((0, 32), np.uint8)
        for i in range(1, N):
            if rng.random() < 0.15:
                d[i] = d[int(rng.integers(0, i))]
        out.append(d)
    return out


def synth_init_frames(seed, kA, dA, kB, dB, p_dup=0.35, max_flips=24, prev_jitter=0.0):
    """Inputs of ORBmatcher::SearchForInitialization: F1 = frame A plus near-duplicates of some of its level-0 keypoints (a few bits
    flipped, some not at all, so that several i1 compete for one keypoint of F2: skips on `vMatchedDistance[i2] <= dist`, take-overs
    on a smaller distance, exact ties), in shuffled order; F2 = frame B, where a share of the level-0 descriptors is replaced by noisy
    copies of A's so that TH_LOW and the ratio test pass often. Returns (k1, d1, prev, k2, d2) with prev = vbPrevMatched [n1, 2]."""
    rng = np.random.default_rng(seed)
    k2, d2 = kB.copy(), dB.copy()
    lvl0A = np.nonzero(kA["octave"] == 0)[0]
    lvl0B = np.nonzero(kB["octave"] == 0)[0]
    if len(lvl0A) and len(lvl0B):
        # pair every level-0 keypoint of B with the nearest level-0 keypoint of A and copy its descriptor with noise
        for j in lvl0B[rng.random(len(lvl0B)) < 0.7]:
            dx = kA["x"][lvl0A] - kB["x"][j]; dy = kA["y"][lvl0A] - kB["y"][j]
            i = lvl0A[int(np.argmin(dx * dx + dy * dy))]
            d2[j] = flip_bits(rng, dA[i:i + 1], 30)[0]
    dup = lvl0A[rng.random(len(lvl0A)) < p_dup]
    kd = kA[dup].copy()
    kd["x"] += rng.normal(0, 1.5, len(dup)).astype(np.float32)
    kd["y"] += rng.normal(0, 1.5, len(dup)).astype(np.float32)
    flips = rng.integers(0, max_flips + 1, len(dup))
    flips[rng.random(len(dup)) < 0.3] = 0
    dd = dA[dup].copy()
    for t in range(len(dup)):
        if flips[t]:
            dd[t] = flip_bits(rng, dd[t:t + 1], int(flips[t]))[0]
    k1 = np.concatenate([kA, kd]); d1 = np.concatenate([dA, dd])
    perm = rng.permutation(len(k1))
    k1, d1 = k1[perm], d1[perm]
    prev = np.stack([k1["x"], k1["y"]], axis=1).astype(np.float32)
    if prev_jitter > 0:
        prev += rng.normal(0, prev_jitter, prev.shape).astype(np.float32)
    return k1, d1, prev, k2, d2


def synth_two_camera_keyframes(seed, rig=None, npts=900, w=512, h=512, max_flips=22, p_mp=0.3):
    """Two TWO-CAMERA keyframes (fisheye rig) that see the same 3-D points, for the mpCamera2 branch of
    ORBmatcher::SearchForTriangulation. Keyframe 1's left camera is the world frame, its right camera sits at the rig's extrinsics
    (x_left = R12 x_right + t12), keyframe 2 is the same rig moved by a small rotation and ~0.3 m. Every point is observed (with pixel
    noise) by a random subset of the four cameras; observations share a noisy copy of the point's descriptor, an orientation and
    usually a vocabulary node. Returns (k1, k2, rigs): k = dict(kps [left..., right...], desc, has_mp, fv, nleft), rigs = RIG_DTYPE[4]
    in the order left-left, left-right, right-left, right-right with cam1 / cam2 and R12 / t12 of that combination
    (x_cam_of_kp1 = R12 x_cam_of_kp2 + t12)."""
    rng = np.random.default_rng(seed)
    rig = kb8_rig("tumvi") if rig is None else rig
    Rlr = np.asarray(rig["R12"], np.float64); tlr = np.asarray(rig["t12"], np.float64)
    ax = rng.normal(0, 0.03, 3)
    th = np.linalg.norm(ax); k = ax / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    Rm = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
    tm = np.array([0.3, 0.02, 0.05]) + rng.normal(0, 0.02, 3)
    # points in keyframe 1's left frame
    z = np.exp(rng.uniform(np.log(0.6), np.log(25.0), npts))
    ang = rng.uniform(0, 2 * np.pi, npts); rad = np.tan(rng.uniform(0.0, 0.9, npts))
    X_l1 = np.stack([z * rad * np.cos(ang), z * rad * np.sin(ang), z], 1)
    X_r1 = (X_l1 - tlr) @ Rlr
    X_l2 = (X_l1 - tm) @ Rm
    X_r2 = (X_l2 - tlr) @ Rlr
    base_desc = rng.integers(0, 256, (npts, 32), dtype=np.uint8)
    base_angle = rng.uniform(0, 360, npts)
    node_of_pt = rng.integers(0, 60, npts) * 7 + 3

    def observe(Xs, cams, kf):
        kps, desc, nodes, pts = [], [], [], []
        for cam_i, (X, cam) in enumerate(zip(Xs, cams)):
            uv = _kb8_project64(cam, X) + rng.normal(0, 0.4, (npts, 2)) * (rng.random((npts, 1)) < 0.8)
            seen = (rng.random(npts) < 0.7) & (X[:, 2] > 0.1) & (uv[:, 0] > 5) & (uv[:, 0] < w - 5) & (uv[:, 1] > 5) & (uv[:, 1] < h - 5)
            idx = np.nonzero(seen)[0]
            idx = idx[rng.permutation(len(idx))]
            kp = np.zeros(len(idx), KP_DTYPE)
            kp["x"] = uv[idx, 0].astype(np.float32); kp["y"] = uv[idx, 1].astype(np.float32)
            kp["size"] = 31.0; kp["response"] = 20.0
            kp["octave"] = rng.integers(0, 4, len(idx)); kp["class_id"] = -1
            wild = rng.random(len(idx)) < 0.12
            kp["angle"] = np.mod(base_angle[idx] + rng.normal(0, 3, len(idx)) + (12.0 if kf else 0.0) + np.where(wild, rng.uniform(30, 300, len(idx)), 0), 360).astype(np.float32)
            d = flip_bits(rng, base_desc[idx], max_flips)
            junk = rng.random(len(idx)) < 0.1
            d[junk] = rng.integers(0, 256, (int(junk.sum()), 32), dtype=np.uint8)
            nd = np.where(rng.random(len(idx)) < 0.92, node_of_pt[idx], rng.integers(0, 60, len(idx)) * 7 + 3)
            kps.append(kp); desc.append(d); nodes.append(nd); pts.append(idx)
        nleft = len(kps[0])
        kp = np.concatenate(kps); d = np.concatenate(desc); nd = np.concatenate(nodes)
        ids = np.unique(nd)
        off = np.zeros(len(ids) + 1, np.int32); feat = []
        for j, v in enumerate(ids):
            feat.extend(np.nonzero(nd == v)[0].tolist())
            off[j + 1] = len(feat)
        fv = dict(fv_node=ids.astype(np.uint32), fv_off=off, fv_feat=np.array(feat, np.uint32))
        return dict(kps=kp, desc=d, uright=None, has_mp=(rng.random(len(kp)) < p_mp).astype(np.uint8), fv=fv, nleft=nleft)

    k1 = observe([X_l1, X_r1], [rig["cam1"], rig["cam2"]], 0)
    k2 = observe([X_l2, X_r2], [rig["cam1"], rig["cam2"]], 1)
    combos = [(Rm, tm), (Rm @ Rlr, Rm @ tlr + tm), (Rlr.T @ Rm, Rlr.T @ (tm - tlr)), (Rlr.T @ Rm @ Rlr, Rlr.T @ (Rm @ tlr + tm - tlr))]
    rigs = np.zeros(4, RIG_DTYPE)
    for c, (R, t) in enumerate(combos):
        rigs[c]["cam1"] = rig["cam2"] if c >= 2 else rig["cam1"]
        rigs[c]["cam2"] = rig["cam2"] if c % 2 else rig["cam1"]
        rigs[c]["prec1"] = rig["prec2"] if c >= 2 else rig["prec1"]
        rigs[c]["prec2"] = rig["prec2"] if c % 2 else rig["prec1"]
        rigs[c]["R12"] = R.astype(np.float32).reshape(9)
        rigs[c]["t12"] = t.astype(np.float32)
    return k1, k2, rigs


KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"),
                     ("class_id", "<i4")])   # cv::KeyPoint / orb_keypoint
RIG_DTYPE = np.dtype([("cam1", "<f4", 8), ("cam2", "<f4", 8), ("prec1", "<f4"), ("prec2", "<f4"), ("R12", "<f4", 9), ("t12", "<f4", 3)])   # orb_kb8_rig
