"""Parity of the CUDA path (through the C ABI, liborb_b200.so) against the CPU oracle on the same seeded
inputs, stage by stage and end to end. Bit-exact for pixels, keypoints (coordinates, order, response,
size, octave), descriptors and match indices; angles / disparities are also compared bit-exactly (the
north-star tolerance is 1e-4 rad / 1e-3 px, asserted as well so a tolerance-only failure is visible)."""
import os

import numpy as np
import pytest

from morb_slam_b200 import capi, synth
from oracle import oracle_py as op
from tests.conftest import has_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")]

DIAG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def _diag(name, text):
    try:
        os.makedirs(DIAG, exist_ok=True)
        with open(os.path.join(DIAG, "diag_" + name + ".txt"), "a") as f:
            f.write(text + "\n")
    except Exception:
        pass


def _first_diff(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return "shape %s vs %s" % (a.shape, b.shape)
    d = np.argwhere(a != b)
    return "%d diffs, first at %s: %s vs %s" % (len(d), d[0] if len(d) else None,
                                                a[tuple(d[0])] if len(d) else None, b[tuple(d[0])] if len(d) else None)


def _mk(cfg, **kw):
    w, h, nf, lap, fx, b = synth.CONFIGS[cfg]
    return capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, **kw)


STAGE_CASES = [("euroc", 2000), ("euroc_mono", 1000), ("tumvi", 3000), ("kitti", 4000)]


@pytest.mark.parametrize("cfg,seed", STAGE_CASES)
def test_stages_match_oracle(cfg, seed):
    w, h, nf, lap, fx, b = synth.CONFIGS[cfg]
    img = synth.mono_frame(seed, w, h)
    ex = _mk(cfg)
    o = op.OracleExtractor(nf)
    mo, ko, do = o(img, lap)
    mg, kg, dg = ex(img, lap)
    t = ex.tables(); to = o.tables()
    for k in ("scale", "inv_scale", "sigma2", "inv_sigma2", "nfeat"):
        assert np.array_equal(t[k], to[k]), k
    errs = []
    for l in range(8):
        if not np.array_equal(ex.pyramid_level(l), o.level(l)):
            errs.append("pyramid L%d: %s" % (l, _first_diff(ex.pyramid_level(l), o.level(l))))
        if not np.array_equal(ex.blurred_level(l), o.blurred(l)):
            errs.append("blur L%d: %s" % (l, _first_diff(ex.blurred_level(l), o.blurred(l))))
        cg, co = ex.candidates(l), o.candidates(l)
        if not np.array_equal(cg, co):
            errs.append("fast L%d: n %d vs %d; %s" % (l, len(cg), len(co), _first_diff(cg[:min(len(cg), len(co))], co[:min(len(cg), len(co))])))
        sg = ex.selected(l)
        so_k = o.level_keypoints(l)
        so = np.stack([so_k["x"] - 16, so_k["y"] - 16, so_k["response"]], 1).astype(np.int32)
        if not np.array_equal(sg, so):
            errs.append("octree L%d: n %d vs %d; %s" % (l, len(sg), len(so), _first_diff(sg[:min(len(sg), len(so))], so[:min(len(sg), len(so))])))
    if mg != mo:
        errs.append("mono %d vs %d" % (mg, mo))
    if len(kg) != len(ko):
        errs.append("K %d vs %d" % (len(kg), len(ko)))
    else:
        for fld in ("x", "y", "size", "response", "octave", "class_id"):
            if not np.array_equal(kg[fld], ko[fld]):
                errs.append("kp.%s: %s" % (fld, _first_diff(kg[fld], ko[fld])))
        if kg["angle"].tobytes() != ko["angle"].tobytes():
            d = np.abs(kg["angle"] - ko["angle"])
            errs.append("angle: %d differ, max %.3g deg" % ((d > 0).sum(), d.max()))
        if not np.array_equal(dg, do):
            errs.append("desc: rows differing %d of %d" % ((dg != do).any(1).sum(), len(do)))
    if errs:
        _diag("stages_%s_%d" % (cfg, seed), "\n".join(errs))
    assert not errs, "\n".join(errs)
    # north-star tolerances (implied by the exact match above)
    assert np.all(np.abs(np.deg2rad(kg["angle"]) - np.deg2rad(ko["angle"])) <= 1e-4)


@pytest.mark.parametrize("maker", ["flat_frame", "plateau_frame", "noise", "constant"])
def test_edge_frames(maker):
    if maker == "noise":
        img = np.random.default_rng(5).integers(0, 256, (480, 752), dtype=np.uint8)
    elif maker == "constant":
        img = np.full((480, 752), 77, np.uint8)
    else:
        img = getattr(synth, maker)(11)
    ex = capi.ORBextractor(1200, max_width=752, max_height=480)
    o = op.OracleExtractor(1200)
    mo, ko, do = o(img, (0, 0))
    mg, kg, dg = ex(img, (0, 0))
    assert mg == mo and len(kg) == len(ko)
    assert kg.tobytes() == ko.tobytes(), _first_diff(kg["x"], ko["x"])
    assert np.array_equal(dg, do)


def test_empty_and_invalid_inputs():
    ex = capi.ORBextractor(500, max_width=752, max_height=480)
    assert ex(None)[0] == -1                       # operator() returns -1 on an empty image
    with pytest.raises(capi.OrbError):
        ex(np.zeros((481, 752), np.uint8))         # larger than the handle
    with pytest.raises(capi.OrbError):
        ex(np.zeros((60, 60), np.uint8))           # no room for one FAST cell at the top level
    with pytest.raises(capi.OrbError):
        capi.ORBextractor(0)                       # invalid nfeatures


def test_strided_input_and_lapping_variants():
    big = synth.mono_frame(77, 800, 500)
    view = big[10:490, 20:772]                     # non-contiguous rows: honour the stride
    ex = capi.ORBextractor(1000, max_width=752, max_height=480)
    o = op.OracleExtractor(1000)
    for lap in ((0, 1000), (0, 0), (300, 500)):
        mo, ko, do = o(np.ascontiguousarray(view), lap)
        mg, kg, dg = ex(view, lap)
        assert mg == mo and kg.tobytes() == ko.tobytes() and np.array_equal(dg, do), lap


def test_batch_equals_single_frames():
    w, h, nf, lap, fx, b = synth.CONFIGS["euroc"]
    B = 6
    imgs = np.stack([synth.mono_frame(2100 + i, w, h) for i in range(B)])
    ex = _mk("euroc", max_batch=B)
    n, mono, kps, desc = ex.extract_batch(imgs, lap)
    o = op.OracleExtractor(nf)
    for i in range(B):
        mo, ko, do = o(imgs[i], lap)
        assert n[i] == len(ko) and mono[i] == mo, i
        assert kps[i, :n[i]].tobytes() == ko.tobytes(), i
        assert np.array_equal(desc[i, :n[i]], do), i
    # a smaller batch and a smaller image on the same handle
    n2, mono2, kps2, desc2 = ex.extract_batch(imgs[:2, :400, :600].copy(), lap)
    for i in range(2):
        mo, ko, do = o(np.ascontiguousarray(imgs[i, :400, :600]), lap)
        assert n2[i] == len(ko) and kps2[i, :n2[i]].tobytes() == ko.tobytes() and np.array_equal(desc2[i, :n2[i]], do)


def _random_cands(rng, w, h, n):
    pos = rng.choice(w * h, size=min(n, w * h), replace=False)
    pos.sort()
    xs, ys = pos % w, pos // w
    c = np.stack([xs, ys, rng.integers(7, 120, len(pos))], 1).astype(np.int32)
    perm = np.argsort((ys // 38) * 1000 + (xs // 36), kind="stable")
    return c[perm]


@pytest.mark.parametrize("region", [(720, 448), (480, 480), (1209, 344), (178, 102)])
def test_octree_fuzz(region):
    w, h = region
    rng = np.random.default_rng(w * 7 + h)
    ex = capi.ORBextractor(1000, max_width=752, max_height=480)
    bad = []
    for trial in range(30):
        n = int(rng.integers(1, 7000))
        N = int(rng.integers(1, 500))
        c = _random_cands(rng, w, h, n)
        a = ex.distribute(c, w, h, N)
        b = op.oracle_distribute(c, w, h, N)
        if not np.array_equal(a, b):
            bad.append((trial, n, N, len(a), len(b)))
    if bad:
        _diag("octree_%dx%d" % region, str(bad))
    assert not bad, bad


def test_octree_clustered_and_ties():
    rng = np.random.default_rng(3)
    ex = capi.ORBextractor(1000, max_width=752, max_height=480)
    w, h = 720, 448
    for trial in range(40):
        n = int(rng.integers(1, 40)) if trial % 2 else int(rng.integers(200, 3000))
        cx, cy = rng.integers(0, w), rng.integers(0, h)
        xs = np.clip(rng.normal(cx, 25, n).astype(int), 0, w - 1)
        ys = np.clip(rng.normal(cy, 25, n).astype(int), 0, h - 1)
        pos = np.unique(ys * w + xs)
        c = np.stack([pos % w, pos // w, rng.integers(7, 30, len(pos))], 1).astype(np.int32)
        N = int(rng.integers(1, 300))
        assert np.array_equal(ex.distribute(c, w, h, N), op.oracle_distribute(c, w, h, N)), trial


@pytest.mark.parametrize("kernel", ["passes", "warp"])
def test_octree_kernels(monkeypatch, kernel):
    """both quad-tree kernels - the block-parallel pass form (default, orb_kernel_octree_passes.cuh) and the one-warp list form
    (ORB_B200_OCTREE=warp): same candidates in, same keypoints out as the oracle - fuzz through orb_debug_distribute and a whole
    extraction."""
    monkeypatch.setenv("ORB_B200_OCTREE", kernel)      # read by orb_create
    ex = capi.ORBextractor(1000, max_width=752, max_height=480)
    rng = np.random.default_rng(99)
    for trial in range(60):
        w, h = [(720, 448), (480, 480), (1209, 344), (178, 102)][trial % 4]
        if trial % 3 == 2:
            n = int(rng.integers(1, 3000))
            cx, cy = rng.integers(0, w), rng.integers(0, h)
            xs = np.clip(rng.normal(cx, 25, n).astype(int), 0, w - 1)
            ys = np.clip(rng.normal(cy, 25, n).astype(int), 0, h - 1)
            pos = np.unique(ys * w + xs)
            c = np.stack([pos % w, pos // w, rng.integers(7, 30, len(pos))], 1).astype(np.int32)
        else:
            c = _random_cands(rng, w, h, int(rng.integers(1, 7000)))
        N = int(rng.integers(1, 500))
        assert np.array_equal(ex.distribute(c, w, h, N), op.oracle_distribute(c, w, h, N)), (trial, len(c), N)
    w, h, nf, lap, fx, b = synth.CONFIGS["kitti"]
    imgs = np.stack([synth.mono_frame(4300 + i, w, h) for i in range(3)])
    exk = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=3)
    n, mono, kps, desc = exk.extract_batch(imgs, lap)
    for f in range(3):
        o = op.OracleExtractor(nf)
        mo, ko, do = o(imgs[f], lap)
        assert n[f] == len(ko) and kps[f, :n[f]].tobytes() == ko.tobytes() and np.array_equal(desc[f, :n[f]], do), f


def test_device_std_sort_equals_libstdcxx():
    """the quad-tree kernel's std::sort (one warp runs libstdc++'s __introsort_loop with ballot-built partitions, the block ranks the
    insertion-sort half) on arbitrary records: same order as the oracle's emulation and, where oracle/_ref is present, as the real
    std::sort - heavy ties (the order of equal keys is what the unstable sort leaves), sorted / reversed / organ-pipe / constant inputs"""
    ex = capi.ORBextractor(1000, max_width=752, max_height=480)
    lib = op.oracle_lib()
    ref = op.ref_lib() if os.path.exists(op.REF_SO) else None
    _p = op._p
    rng = np.random.default_rng(7)
    cases = []
    for trial in range(300):
        n = int(rng.integers(1, 900))
        nk = int(rng.integers(1, 14)) if trial % 3 else int(rng.integers(1, 1 << 20))
        keys = rng.integers(0, nk, n).astype(np.uint32)
        if trial % 5 == 0:
            keys = np.sort(keys)[:: (1 if trial % 2 else -1)].copy()
        cases.append(keys)
    for n in (17, 33, 3000, 8000):
        cases.append(np.concatenate([np.arange(n // 2), np.arange(n - n // 2)[::-1]]).astype(np.uint32))   # organ pipe
        cases.append(np.zeros(n, np.uint32))
        cases.append((np.arange(n) % 3).astype(np.uint32))
    for t, keys in enumerate(cases):
        n = len(keys)
        ko, po = ex.std_sort(keys)
        k1, p1 = keys.copy(), np.arange(n, dtype=np.uint32)
        lib.oro_introsort(_p(k1), _p(p1), n)
        assert np.array_equal(ko, k1) and np.array_equal(po, p1), (t, n)
        if ref is not None:
            k2, p2 = keys.copy(), np.arange(n, dtype=np.uint32)
            ref.ref_std_sort(_p(k2), _p(p2), n)
            assert np.array_equal(ko, k2) and np.array_equal(po, p2), (t, n)


@pytest.mark.parametrize("pipe", ["1", "0"])
def test_small_batch_pipeline_forms(monkeypatch, pipe):
    """batches of at most 8 frames run FAST + the quad-tree of every level on the level's own stream as soon as the level exists
    (default) or as one launch per stage (ORB_B200_LEVEL_PIPE=0); both equal the oracle, on plain launches (first call) and on the
    replayed CUDA graph (later calls), for batch 1, 3 and 8, with and without a lapping area, and on two interleaved handles"""
    monkeypatch.setenv("ORB_B200_LEVEL_PIPE", pipe)      # read by orb_create
    for cfg in ("euroc", "tumvi"):
        w, h, nf, lap, fx, b = synth.CONFIGS[cfg]
        exA = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=8)
        exB = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=8)
        imgs = np.stack([synth.mono_frame(7300 + i, w, h) for i in range(8)])
        oracle = []
        for f in range(8):
            o = op.OracleExtractor(nf)
            oracle.append(o(imgs[f], lap))
        pinned = {}
        for B in (1, 3, 8, 1):
            for rep in range(4):          # plain launches, capture, replay; the last repetition hands in page-locked result buffers
                sel = [(rep + i) % 8 for i in range(B)]
                outs = []
                for ex in (exA, exB):
                    out = None
                    if rep == 3 or B == 8:   # (the descriptor kernel then writes the records straight into them: no copies)
                        out = pinned.setdefault((id(ex), B), (capi.pinned_empty((B,), np.int32), capi.pinned_empty((B,), np.int32),
                                                              capi.pinned_empty((B, ex.kcap), capi.KP_DTYPE),
                                                              capi.pinned_empty((B, ex.kcap, 32), np.uint8)))
                    outs.append(ex.extract_batch(imgs[sel], lap, out=out, flags=capi.ORB_ASYNC))
                exA.sync(); exB.sync()
                for n, mono, kps, desc in outs:
                    for i, f in enumerate(sel):
                        mo, ko, do = oracle[f]
                        assert n[i] == len(ko) and mono[i] == mo and kps[i, :n[i]].tobytes() == ko.tobytes() and np.array_equal(desc[i, :n[i]], do), (cfg, B, rep, i)


def test_page_locked_result_buffers_of_any_capacity():
    """small batches write their results straight into page-locked buffers of the caller (no copies): rows of `cap` records per frame with
    cap larger than, equal to and smaller than the number of keypoints - nothing lands outside a frame's row, a too small cap is
    reported like with pageable buffers, and the stereo matcher's mvuRight / mvDepth behave the same"""
    w, h, nf, lap, fx, b = synth.CONFIGS["euroc"]
    B = 2
    pairs = [synth.stereo_pair(8800 + i, w, h) for i in range(B)]
    L = np.stack([p[0] for p in pairs]); R = np.stack([p[1] for p in pairs])
    exL, exR = _mk("euroc", max_batch=B), _mk("euroc", max_batch=B)
    n0, m0, k0, d0 = exL.extract_batch(L, lap)                     # pageable buffers: the copies
    exR.extract_batch(R, lap)
    u0 = np.zeros((B, exL.kcap), np.float32); p0 = np.zeros((B, exL.kcap), np.float32)
    capi.compute_stereo_matches_batch(exL, exR, np.float32(fx * b), np.float32(fx), out=(u0, p0))
    for cap in (exL.kcap + 57, exL.kcap, int(n0.max()), 500):
        n = capi.pinned_empty((B,), np.int32); m = capi.pinned_empty((B,), np.int32)
        kp = capi.pinned_empty((B + 1, cap), capi.KP_DTYPE); ds = capi.pinned_empty((B + 1, cap, 32), np.uint8)
        kp.view(np.uint8)[:] = 0xA5; ds[:] = 0xA5
        for rep in range(3):                                       # plain launches, capture, replay
            if cap >= n0.max():
                exL.extract_batch(L, lap, out=(n, m, kp[:B], ds[:B]))
                for f in range(B):
                    assert n[f] == n0[f] and m[f] == m0[f]
                    assert kp[f, :n[f]].tobytes() == k0[f, :n0[f]].tobytes() and np.array_equal(ds[f, :n[f]], d0[f, :n0[f]])
                    assert np.all(kp[f, n[f]:].view(np.uint8) == 0xA5) and np.all(ds[f, n[f]:] == 0xA5)    # only the n records are written
            else:
                with pytest.raises(capi.OrbError):
                    exL.extract_batch(L, lap, out=(n, m, kp[:B], ds[:B]))
                for f in range(B):
                    assert kp[f, :cap].tobytes() == k0[f, :cap].tobytes() and np.array_equal(ds[f], d0[f, :cap])
            assert np.all(kp[B].view(np.uint8) == 0xA5) and np.all(ds[B] == 0xA5)                      # the row behind the batch is untouched
        if cap >= n0.max():
            exR.extract_batch(R, lap)
            u = capi.pinned_empty((B + 1, cap), np.float32); p = capi.pinned_empty((B + 1, cap), np.float32)
            u[:] = 7.0; p[:] = 7.0
            capi.compute_stereo_matches_batch(exL, exR, np.float32(fx * b), np.float32(fx), out=(u[:B], p[:B]))
            for f in range(B):
                assert u[f, :n0[f]].tobytes() == u0[f, :n0[f]].tobytes() and p[f, :n0[f]].tobytes() == p0[f, :n0[f]].tobytes()
            assert np.all(u[B] == 7.0) and np.all(p[B] == 7.0)


@pytest.mark.parametrize("cfg,seed", [("euroc", 2000), ("euroc", 2003), ("kitti", 4000)])
def test_stereo_matches_oracle(cfg, seed):
    w, h, nf, lap, fx, b = synth.CONFIGS[cfg]
    B = 3
    pairs = [synth.stereo_pair(seed + 10 * i, w, h) for i in range(B)]
    L = np.stack([p[0] for p in pairs]); R = np.stack([p[1] for p in pairs])
    exL, exR = _mk(cfg, max_batch=B), _mk(cfg, max_batch=B)
    nL, mL, kL, dL = exL.extract_batch(L, lap)
    nR, mR, kR, dR = exR.extract_batch(R, lap)
    mbf = np.float32(fx * b)
    maxD = np.float32(fx)
    uR = np.zeros((B, exL.kcap), np.float32); dp = np.zeros((B, exL.kcap), np.float32)
    capi.compute_stereo_matches_batch(exL, exR, mbf, maxD, out=(uR, dp))
    # page-locked result buffers (the gate kernel writes them itself for a few frames): same values
    uRp = capi.pinned_empty((B, exL.kcap), np.float32); dpp = capi.pinned_empty((B, exL.kcap), np.float32)
    uRp[:] = 7.0; dpp[:] = 7.0
    capi.compute_stereo_matches_batch(exL, exR, mbf, maxD, out=(uRp, dpp))
    for i in range(B):
        assert uRp[i, :nL[i]].tobytes() == uR[i, :nL[i]].tobytes() and dpp[i, :nL[i]].tobytes() == dp[i, :nL[i]].tobytes()
    oL, oR = op.OracleExtractor(nf), op.OracleExtractor(nf)
    for i in range(B):
        _, koL, doL = oL(L[i], lap)
        _, koR, doR = oR(R[i], lap)
        assert kL[i, :nL[i]].tobytes() == koL.tobytes() and kR[i, :nR[i]].tobytes() == koR.tobytes()
        u_o, d_o, bi_o, bd_o = op.oracle_stereo(oL, oR, koL, doL, koR, doR, float(mbf), float(maxD), want_best=True)
        bi_g, bd_g = capi.stereo_best(exL, i)
        assert np.array_equal(bi_g[:nL[i]], bi_o), _first_diff(bi_g[:nL[i]], bi_o)   # match indices bit-exact
        assert np.array_equal(bd_g[:nL[i]], bd_o)
        assert (u_o >= 0).sum() > 0.3 * nL[i]
        assert np.array_equal(uR[i, :nL[i]] >= 0, u_o >= 0)
        assert np.all(np.abs(uR[i, :nL[i]] - u_o) <= 1e-3)                          # north-star tolerance
        assert uR[i, :nL[i]].tobytes() == u_o.tobytes() and dp[i, :nL[i]].tobytes() == d_o.tobytes()
    # single-frame form with host keypoints (the reference's Frame members)
    u1, d1 = capi.compute_stereo_matches(exL, exR, kL[0, :nL[0]], dL[0, :nL[0]], kR[0, :nR[0]], dR[0, :nR[0]], mbf, maxD)
    assert u1.tobytes() == uR[0, :nL[0]].tobytes() and d1.tobytes() == dp[0, :nL[0]].tobytes()


def test_stereo_without_matches_and_identical_images():
    w, h, nf, lap, fx, b = synth.CONFIGS["euroc"]
    A, Bm = synth.mono_frame(9, w, h), synth.mono_frame(10, w, h)
    exL, exR = _mk("euroc"), _mk("euroc")
    oL, oR = op.OracleExtractor(nf), op.OracleExtractor(nf)
    for left, right in ((A, Bm), (A, A)):
        _, kL, dL = exL(left, lap)
        _, kR, dR = exR(right, lap)
        oL(left, lap); oR(right, lap)
        u_g, d_g = capi.compute_stereo_matches(exL, exR, kL, dL, kR, dR, fx * b, fx)
        u_o, d_o = op.oracle_stereo(oL, oR, kL, dL, kR, dR, float(np.float32(fx * b)), float(np.float32(fx)))
        assert u_g.tobytes() == u_o.tobytes() and d_g.tobytes() == d_o.tobytes()


def test_knn2_matches_oracle_and_tie_rule():
    ex = capi.ORBextractor(1000, max_width=752, max_height=480)
    q = synth.random_descriptors(0, 1200)
    db = synth.random_descriptors(1, 50000)
    db[1234] = q[7]; db[4321] = q[7]            # exact duplicates: lower index first
    db[40000] = db[39999]
    idx, dist = capi.hamming_knn2(ex, q, db)
    io, do = op.oracle_knn2(q, db, threads=8)
    assert np.array_equal(idx, io) and np.array_equal(dist, do)
    assert list(idx[7]) == [1234, 4321] and list(dist[7]) == [0, 0]
    # clustered database: heavy distance ties
    dbc = synth.clustered_descriptors(2, q[:200], 20000, max_flips=3)
    idx, dist = capi.hamming_knn2(ex, q[:200], dbc)
    io, do = op.oracle_knn2(q[:200], dbc, threads=8)
    assert np.array_equal(idx, io) and np.array_equal(dist, do)
    # ragged sizes: 1 query, 1 row, 2 rows, sizes off tile boundaries
    for nq, ndb in ((1, 1), (1, 2), (3, 127), (33, 129), (1025, 1000), (700, 5000)):
        idx, dist = capi.hamming_knn2(ex, q[:nq] if nq <= len(q) else synth.random_descriptors(5, nq), db[:ndb])
        io, do = op.oracle_knn2(q[:nq] if nq <= len(q) else synth.random_descriptors(5, nq), db[:ndb])
        assert np.array_equal(idx, io) and np.array_equal(dist, do), (nq, ndb)
    # Lowe ratio gate, evaluated in double like the reference
    d = np.array([[7, 10], [6, 10], [14, 20], [21, 30], [0, 0], [5, -1], [69, 99], [70, 100]], np.int32)
    assert np.array_equal(capi.ratio_test(ex, d), op.oracle_ratio_test(d))


def test_knn2_sharded_merge_equals_single_scan():
    ex = capi.ORBextractor(1000, max_width=752, max_height=480)
    q = synth.random_descriptors(3, 300)
    db = synth.clustered_descriptors(4, q, 40000, max_flips=40)
    full_i, full_d = capi.hamming_knn2(ex, q, db)
    parts_i, parts_d = [], []
    G = 8
    per = len(db) // G
    for g in range(G):
        i, d = capi.hamming_knn2(ex, q, db[g * per:(g + 1) * per], index_base=g * per)
        parts_i.append(i); parts_d.append(d)
    mi, md = capi.knn2_merge(ex, np.stack(parts_i), np.stack(parts_d))
    assert np.array_equal(mi, full_i) and np.array_equal(md, full_d)


def test_descriptor_distance_host_helper():
    rng = np.random.default_rng(8)
    a = rng.integers(0, 256, (50, 32), dtype=np.uint8)
    lib = op.oracle_lib()
    for i in range(49):
        assert capi.descriptor_distance(a[i], a[i + 1]) == int(np.unpackbits(a[i] ^ a[i + 1]).sum())


PARAM_VARIANTS = [
    (752, 480, 5000, 1.2, 8, 20, 7, (0, 0)),
    (1241, 376, 2000, 1.2, 8, 12, 7, (0, 0)),
    (640, 480, 800, 2.0, 3, 20, 7, (0, 0)),
    (800, 600, 1000, 1.5, 5, 25, 10, (100, 400)),
    (512, 512, 1500, 1.1, 12, 20, 7, (0, 511)),
    (640, 480, 800, 1.8, 3, 20, 7, (0, 0)),      # widest ratio of the tile kernel: 4 columns span 8 source bytes (word loads)
    (752, 480, 1200, 1.33, 6, 20, 7, (0, 0)),
]


@pytest.mark.parametrize("w,h,nf,sf,nl,ini,mn,lap", PARAM_VARIANTS)
def test_parameter_variants(w, h, nf, sf, nl, ini, mn, lap):
    """Other constructor arguments than the EuRoC defaults: 5x nFeatures (initialisation extractor), KITTI04-12
    thresholds, exact-2x pyramid (OpenCV's box-filter shortcut), scale 1.8 / 1.5 / 1.33 / 1.1, 3 / 5 / 6 / 12 levels."""
    img = synth.mono_frame(w + nl, w, h)
    ex = capi.ORBextractor(nf, sf, nl, ini, mn, max_width=w, max_height=h)
    o = op.OracleExtractor(nf, sf, nl, ini, mn)
    mo, ko, do = o(img, lap)
    mg, kg, dg = ex(img, lap)
    errs = []
    for l in range(nl):
        if not np.array_equal(ex.pyramid_level(l), o.level(l)):
            errs.append("pyramid L%d: %s" % (l, _first_diff(ex.pyramid_level(l), o.level(l))))
        if not np.array_equal(ex.candidates(l), o.candidates(l)):
            errs.append("fast L%d" % l)
    assert not errs, errs
    assert mg == mo and len(kg) == len(ko)
    assert kg.tobytes() == ko.tobytes() and np.array_equal(dg, do)


@pytest.mark.parametrize("lap", [(0, 511), (120, 380), (200, 210), (600, 700)])
def test_fisheye_stereo_matches_batch(lap):
    """orb_stereo_fisheye_match_batch = Frame::ComputeStereoFishEyeMatches (src/Frame.cc:1222-1250) before the triangulation, on the
    device-resident descriptors of a TUM-VI-shape batch: knnMatch(k = 2) of the lapping-area descriptors + Lowe's ratio, per frame.
    Lapping areas: the whole image (TUM-VI.yaml), a part, a sliver (few or no stereo keypoints), outside the image (none)."""
    w, h, nf = synth.CONFIGS["tumvi"][:3]
    B = 4
    Ls = np.stack([synth.stereo_pair(7100 + i, w, h)[0] for i in range(B)])
    Rs = np.stack([synth.stereo_pair(7100 + i, w, h)[1] for i in range(B)])
    exL = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B)
    exR = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B)
    nL, mL, kL, dL = exL.extract_batch(Ls, lap)
    nR, mR, kR, dR = exR.extract_batch(Rs, lap)
    idx, dist, ok = capi.compute_stereo_fisheye_matches_batch(exL, exR)
    for f in range(B):
        q, t = dL[f, mL[f]:nL[f]], dR[f, mR[f]:nR[f]]
        nq = len(q)
        if nq and len(t):
            io, do = op.oracle_knn2(q, t)
            assert np.array_equal(idx[f, :nq], io) and np.array_equal(dist[f, :nq], do), f
            assert np.array_equal(ok[f, :nq], op.oracle_ratio_test(do)), f
        else:
            assert np.all(idx[f, :nq] == -1) and not ok[f].any()
        assert np.all(idx[f, nq:] == -1) and not ok[f, nq:].any()
    if lap == (0, 511):
        assert mL[0] == 0 and ok[0].sum() > 100
    if lap == (600, 700):
        assert mL[0] == nL[0]


def test_level_capacity_is_reported():
    """The internal capacities are loud, never silent (the reference has none: vToDistributeKeys only reserves nfeatures * 10).
    A lattice of isolated bright dots, 4 px apart, makes every dot a FAST corner and a local maximum: (1062 / 4)^2 = 70 k keypoint
    candidates on level 0 of an 1100 x 1100 image exceed the 65 535 the quad-tree accepts per level -> ORB_ERR_CAPACITY, no output.
    (The per-cell limit of 512 cannot be reached by 35..44-px cells: strict 3x3 non-maximum suppression leaves at most one
    keypoint per 2 x 2 block, 22 * 22 = 484.) The same image at 12-px spacing extracts normally and equals the oracle."""
    w = h = 1100
    img = np.full((h, w), 100, np.uint8)
    img[::4, ::4] = 255
    ex = capi.ORBextractor(1000, 1.2, 8, 20, 7, max_width=w, max_height=h)
    with pytest.raises(capi.OrbError) as e:
        ex(img, (0, 0))
    assert e.value.status == capi.ORB_ERR_CAPACITY
    img = np.full((h, w), 100, np.uint8)
    img[::12, ::12] = 255
    mono, kps, desc = ex(img, (0, 0))
    o = op.OracleExtractor(1000)
    mo, ko, do = o(img, (0, 0))
    assert mono == mo and kps.tobytes() == ko.tobytes() and np.array_equal(desc, do)
