#!/usr/bin/env python3
"""Golden vectors of the rank-3 / rank-4 rows from the REFERENCE's own lines (oracle/_ref/libmorb_ref_kb8.so: KannalaBrandt8.cpp
:68-94,111-147,323-395,415-428 on the Eigen stand-in; oracle/_ref/libmorb_ref_ser.so: SerializationUtils.h:74-152 on a raw-bytes
archive) on the deterministic inputs of morb_slam_b200/synth.py. Run in the build container (needs /root/reference):

    python tools/gen_golden_kb8.py
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from morb_slam_b200 import synth  # noqa: E402
from oracle import oracle_kb8_py as ok  # noqa: E402
from oracle import oracle_ser_py as osr  # noqa: E402
from oracle.oracle_py import KP_DTYPE  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
KB8 = [("tumvi", 900), ("parallel", 901), ("toed", 902)]
N = 2000


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


def ser_keypoints(seed=77, n=500):
    rng = np.random.default_rng(seed)
    k = np.zeros(n, KP_DTYPE)
    for f in ("x", "y", "size", "angle", "response"):
        k[f] = rng.uniform(0, 700, n).astype(np.float32)
    k["octave"] = rng.integers(0, 8, n); k["class_id"] = -1
    return k


def main():
    r = ok.reference()
    for kind, seed in KB8:
        rig = synth.kb8_rig(kind)
        xy1, xy2, s1, s2 = synth.kb8_pairs(seed, rig, N)
        ret, p3d, _ = r.triangulate(rig, xy1, xy2, s1, s2)
        np.savez_compressed(os.path.join(OUT, "kb8_%s.npz" % kind), in_crc=np.uint64(crc(xy1) ^ crc(xy2) ^ crc(s1) ^ crc(s2)), ret=ret, p3d=p3d)
        print("kb8", kind, "accepted", int((ret > 0).sum()))
    ref = osr.Reference()
    k = ser_keypoints()
    d = synth.random_descriptors(78, 500)
    np.savez_compressed(os.path.join(OUT, "ser_fragments.npz"), kps_crc=np.uint64(crc(k)), desc_crc=np.uint64(crc(d)),
                        keypoints=np.frombuffer(ref.serialize_keypoints(k), np.uint8), descriptors=np.frombuffer(ref.serialize_matrix(d), np.uint8))
    print("ser", 4 + 28 * len(k), 13 + 32 * len(d))


if __name__ == "__main__":
    main()
