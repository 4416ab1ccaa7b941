"""One EuRoC-shape extraction through the C ABI (debug helper for compute-sanitizer / ncu runs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from morb_slam_b200 import capi, synth
w, h, nf, lap, fx, b = synth.CONFIGS["euroc"]
img = synth.mono_frame(1, w, h)
ex = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h)
m, k, d = ex(img, lap)
print("K =", len(k), "mono =", m)
