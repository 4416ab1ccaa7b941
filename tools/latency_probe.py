import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from morb_slam_b200 import capi, synth
w, h, nf, lap, fx, b = synth.CONFIGS["euroc"]
pairs = [synth.stereo_pair(9000 + i, w, h) for i in range(4)]
exL = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=1)
pin = lambda s, d: capi.pinned_empty(s, d)
imL = pin((1, h, w), np.uint8)
outL = (pin((1,), np.int32), pin((1,), np.int32), pin((1, exL.kcap), capi.KP_DTYPE), pin((1, exL.kcap, 32), np.uint8))
for mode in ("full", "no_output"):
    gpu, wall, enq = [], [], []
    for r in range(220):
        imL[0] = pairs[r % 4][0]
        fl = capi.ORB_ASYNC | (capi.ORB_NO_OUTPUT if mode == "no_output" else 0)
        t0 = time.perf_counter()
        exL.timer_start()
        exL.extract_batch(imL, lap, out=outL, flags=fl)
        t1 = time.perf_counter()
        ms = exL.timer_stop()
        t2 = time.perf_counter()
        if r >= 20:
            gpu.append(ms); wall.append((t2 - t0) * 1e3); enq.append((t1 - t0) * 1e3)
    print(mode, "gpu-side ms %.3f  wall %.3f  enqueue (host) %.3f" % (np.median(gpu), np.median(wall), np.median(enq)))
