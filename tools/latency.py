"""Single-frame latency of the drop-in path (the way Tracking calls it: one stereo pair at a time): host image in, host
keypoints / descriptors / mvuRight out, through the C ABI. Prints median / p90 milliseconds for
  a) the reference's call sequence: operator() left, operator() right (both synchronous), ComputeStereoMatches
  b) the asynchronous form: both extractions enqueued on their own streams, stereo match, one synchronisation
Usage: python tools/latency.py [reps]"""
import sys
import time

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from morb_slam_b200 import capi, synth  # noqa: E402


def main(reps=200, quiet=False, cfg="euroc"):
    w, h, nf, lap, fx, b = synth.CONFIGS[cfg]
    fisheye = cfg == "tumvi"
    rig = capi.kb8_rig(synth.kb8_rig("parallel"))
    pairs = [synth.stereo_pair(9000 + i, w, h) for i in range(8)]
    exL = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=1)
    exR = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=1)
    mbf, maxd = float(np.float32(fx * b)), float(np.float32(fx))
    pin = lambda s, d: capi.pinned_empty(s, d)  # noqa: E731
    imL, imR = pin((1, h, w), np.uint8), pin((1, h, w), np.uint8)
    outL = (pin((1,), np.int32), pin((1,), np.int32), pin((1, exL.kcap), capi.KP_DTYPE), pin((1, exL.kcap, 32), np.uint8))
    outR = (pin((1,), np.int32), pin((1,), np.int32), pin((1, exR.kcap), capi.KP_DTYPE), pin((1, exR.kcap, 32), np.uint8))
    st = (pin((1, exL.kcap), np.float32), pin((1, exL.kcap), np.float32))
    k = exL.kcap
    fe = (pin((1, k), np.int32), pin((1, k), np.int32), pin((1, k), np.float32), pin((1, k, 3), np.float32), pin((1, k), np.int8))

    def match(flags):
        if fisheye:   # ComputeStereoFishEyeMatches: kNN + ratio on the device, triangulation, the Frame's members come back
            capi.compute_stereo_fisheye_matches_batch(exL, exR, flags=capi.ORB_ASYNC, want=False)
            exL._check(exL.L.orb_stereo_fisheye_triangulate_batch(exL.h, exR.h, __import__("ctypes").byref(rig), *[capi._p(a) for a in fe], k, flags))
        else:
            capi.compute_stereo_matches_batch(exL, exR, mbf, maxd, out=st, flags=flags)
    res = {}
    for mode in ("sync", "async"):
        ts = []
        for r in range(reps + 20):
            L, R = pairs[r % len(pairs)]
            imL[0], imR[0] = L, R
            t0 = time.perf_counter()
            if mode == "sync":
                exL.extract_batch(imL, lap, out=outL)
                exR.extract_batch(imR, lap, out=outR)
                match(0)
            else:
                exL.extract_batch(imL, lap, out=outL, flags=capi.ORB_ASYNC)
                exR.extract_batch(imR, lap, out=outR, flags=capi.ORB_ASYNC)
                match(capi.ORB_ASYNC)
                exL.sync(); exR.sync()
            ts.append((time.perf_counter() - t0) * 1e3)
        ts = np.array(ts[20:])
        res[mode] = (float(np.median(ts)), float(np.percentile(ts, 90)))
        if not quiet:
            nmatch = int((fe[0][0, :outL[0][0]] >= 0).sum()) if fisheye else int((st[0][0, :outL[0][0]] >= 0).sum())
            print("%-6s %-5s median %.3f ms  p90 %.3f ms  (K = %d / %d, %d stereo matches)" % (cfg, mode, res[mode][0], res[mode][1], outL[0][0],
                                                                                              outR[0][0], nmatch))
    return res


if __name__ == "__main__":
    for c in (sys.argv[2:] or ["euroc"]):
        main(int(sys.argv[1]) if len(sys.argv) > 1 else 200, cfg=c)
