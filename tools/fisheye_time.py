"""Device time of the two matcher kernels of the TUM-VI workload at bench batch size (CUDA events on the handle's stream):
python tools/fisheye_time.py [batch]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morb_slam_b200 import capi, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
w, h, nf, lap = synth.CONFIGS["tumvi"][:4]
D = 8
pairs = [synth.stereo_pair(2000 + i, w, h) for i in range(D)]
Ls = np.stack([pairs[i % D][0] for i in range(B)]); Rs = np.stack([pairs[i % D][1] for i in range(B)])
exL = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B)
exR = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B)
nL, mL, _, _ = exL.extract_batch(Ls, lap)
nR, mR, _, _ = exR.extract_batch(Rs, lap)
rig = capi.kb8_rig(synth.kb8_rig("parallel"))
AS = capi.ORB_ASYNC
for _ in range(3):
    capi.compute_stereo_fisheye_matches_batch(exL, exR, flags=AS, want=False)
    capi.compute_stereo_fisheye_triangulation_batch(exL, exR, rig, flags=AS, want=False)
exL.sync()
R = 20
exL.timer_start()
for _ in range(R):
    capi.compute_stereo_fisheye_matches_batch(exL, exR, flags=AS, want=False)
t_knn = exL.timer_stop() / R
exL.timer_start()
for _ in range(R):
    capi.compute_stereo_fisheye_triangulation_batch(exL, exR, rig, flags=AS, want=False)
t_tri = exL.timer_stop() / R
l2r = capi.compute_stereo_fisheye_triangulation_batch(exL, exR, rig)[0]
q = float(np.mean(nL - mL)); t = float(np.mean(nR - mR))
print("batch %d: knnMatch + ratio %.3f ms (%.0f x %.0f descriptors per frame, %.3g pairs/s), triangulation %.3f ms (%.0f accepted per frame)"
      % (B, t_knn, q, t, B * q * t / (t_knn * 1e-3), t_tri, float((l2r >= 0).sum()) / B))
