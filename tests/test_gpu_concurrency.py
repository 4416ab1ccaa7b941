"""Call patterns of the reference's host side, driven concurrently through the C ABI:
  * Frame::Frame spawns one std::thread per extractor (src/Frame.cc:194-197, :1141-1149), joins both, then runs
    ComputeStereoMatches on the caller's thread - two host threads on two handles, 200 frames;
  * asynchronous calls in right-then-left order, the next extraction enqueued while the stereo matcher of the previous
    pair may still read the right handle's buffers (cross-handle stream ordering).
Results must equal the oracle bit for bit in every iteration."""
import threading

import numpy as np
import pytest

from morb_slam_b200 import capi, synth
from oracle import oracle_py as op
from tests.conftest import has_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")]


def _oracle_pair(L, R, nf, lap, mbf, maxD):
    oL, oR = op.OracleExtractor(nf), op.OracleExtractor(nf)
    mL, kL, dL = oL(L, lap)
    mR, kR, dR = oR(R, lap)
    u, d = op.oracle_stereo(oL, oR, kL, dL, kR, dR, mbf, maxD)
    return (mL, kL, dL), (mR, kR, dR), (u, d)


def test_two_host_threads_like_frame_constructor():
    w, h, nf, lap, fx, b = synth.CONFIGS["euroc"]
    mbf, maxD = float(np.float32(fx * b)), float(np.float32(fx))
    pairs = [synth.stereo_pair(7100 + i, w, h) for i in range(4)]
    want = [_oracle_pair(L, R, nf, lap, mbf, maxD) for L, R in pairs]
    exL = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h)
    exR = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h)
    res = {}

    def extract(ex, img, key):
        try:
            res[key] = ex(img, lap)
        except Exception as e:  # noqa: BLE001 - reported by the assertion below
            res[key] = e

    for it in range(200):
        L, R = pairs[it % 4]
        tL = threading.Thread(target=extract, args=(exL, L, "L"))
        tR = threading.Thread(target=extract, args=(exR, R, "R"))
        tL.start(); tR.start()
        tL.join(); tR.join()
        assert not isinstance(res["L"], Exception) and not isinstance(res["R"], Exception), (it, res)
        (mL, kL, dL), (mR, kR, dR) = res["L"], res["R"]
        uR, dp = capi.compute_stereo_matches(exL, exR, kL, dL, kR, dR, mbf, maxD)
        (woL, woR, wst) = want[it % 4]
        assert mL == woL[0] and kL.tobytes() == woL[1].tobytes() and np.array_equal(dL, woL[2]), it
        assert mR == woR[0] and kR.tobytes() == woR[1].tobytes() and np.array_equal(dR, woR[2]), it
        assert uR.tobytes() == wst[0].tobytes() and dp.tobytes() == wst[1].tobytes(), it


def test_two_host_threads_free_running():
    """the two handles run 100 extractions each without any join in between (different images per thread and iteration)"""
    w, h, nf, lap, fx, b = synth.CONFIGS["euroc"]
    imgs = [synth.mono_frame(7200 + i, w, h) for i in range(6)]
    want = []
    for im in imgs:
        o = op.OracleExtractor(nf)
        want.append(o(im, lap))
    errs = []

    def run(offset):
        try:
            ex = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h)
            for it in range(100):
                k = (offset + it) % len(imgs)
                m, kp, d = ex(imgs[k], lap)
                if not (m == want[k][0] and kp.tobytes() == want[k][1].tobytes() and np.array_equal(d, want[k][2])):
                    errs.append((offset, it))
                    return
        except Exception as e:  # noqa: BLE001
            errs.append((offset, repr(e)))

    ts = [threading.Thread(target=run, args=(o,)) for o in (0, 3)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs


@pytest.mark.parametrize("B", [1, 8])
def test_async_right_then_left(B):
    """R is extracted before L, everything asynchronous, and the next pair is enqueued before the previous results are read:
    hR's next extraction must wait for the stereo kernels on hL's stream that still read hR's pyramid / keypoints."""
    w, h, nf, lap, fx, b = synth.CONFIGS["euroc"]
    mbf, maxD = float(np.float32(fx * b)), float(np.float32(fx))
    sets = []
    for s in range(2):
        pairs = [synth.stereo_pair(7300 + 10 * s + i, w, h) for i in range(B)]
        sets.append((np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs]),
                     [_oracle_pair(p[0], p[1], nf, lap, mbf, maxD) for p in pairs]))
    exL = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B)
    exR = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B)
    pin = capi.pinned_empty
    outs = []
    for s in range(2):
        outs.append(dict(
            L=(pin((B,), np.int32), pin((B,), np.int32), pin((B, exL.kcap), capi.KP_DTYPE), pin((B, exL.kcap, 32), np.uint8)),
            R=(pin((B,), np.int32), pin((B,), np.int32), pin((B, exR.kcap), capi.KP_DTYPE), pin((B, exR.kcap, 32), np.uint8)),
            st=(pin((B, exL.kcap), np.float32), pin((B, exL.kcap), np.float32))))
    A = capi.ORB_ASYNC

    def enqueue(s):
        L, R, _ = sets[s]
        exR.extract_batch(R, lap, out=outs[s]["R"], flags=A)       # right first
        exL.extract_batch(L, lap, out=outs[s]["L"], flags=A)
        capi.compute_stereo_matches_batch(exL, exR, mbf, maxD, out=outs[s]["st"], flags=A)

    def check(s, it):
        _, _, want = sets[s]
        nL, mL, kL, dL = outs[s]["L"]
        nR, mR, kR, dR = outs[s]["R"]
        uR, dp = outs[s]["st"]
        for f in range(B):
            (woL, woR, wst) = want[f]
            assert nL[f] == len(woL[1]) and kL[f, :nL[f]].tobytes() == woL[1].tobytes() and np.array_equal(dL[f, :nL[f]], woL[2]), (it, f)
            assert nR[f] == len(woR[1]) and kR[f, :nR[f]].tobytes() == woR[1].tobytes() and np.array_equal(dR[f, :nR[f]], woR[2]), (it, f)
            assert uR[f, :nL[f]].tobytes() == wst[0].tobytes() and dp[f, :nL[f]].tobytes() == wst[1].tobytes(), (it, f)

    enqueue(0)
    for it in range(1, 40):
        s = it % 2
        enqueue(s)              # overwrites both handles' device buffers while pair it - 1 may still be in flight
        exL.sync(); exR.sync()  # (the extraction of a handle completes its own previous batch first)
        check(1 - s, it - 1)
        check(s, it)
