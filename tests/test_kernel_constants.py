"""Host-side check of the packed coefficient words of the blur kernel (morb_slam_b200/csrc/orb_kernel_blur.cuh) against the
separable kernel OpenCV uses for GaussianBlur 7x7 sigma 2 on 8-bit images, [18 34 48 56 48 34 18] / 256 (reference
src/ORBextractor.cc:1049-1050). The words are read from the source text, so an edit of a constant that
the GPU parity tests would catch only on a GPU box fails here already. No GPU, no compute call."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "morb_slam_b200", "csrc", "orb_kernel_blur.cuh")).read()
TAPS = [18, 34, 48, 56, 48, 34, 18]


def _bytes(word):
    return [(word >> (8 * i)) & 0xFF for i in range(4)]


def test_horizontal_pass_coefficient_words():
    """Output j of a quad (tile byte 16 + 4q + j = byte 4 + j of the 12-byte window w0 w1 w2) takes its 7 taps from window bytes
    j + 1 .. j + 7: the coefficient words of its IDP.4A chain must hold the taps there and zero elsewhere, and start from 128."""
    for j, comp in enumerate("xyzw"):
        m = re.search(r"v\[e\]\.%s = (.*?);" % comp, SRC)
        assert m, "horizontal sum of component %s not found" % comp
        expr = m.group(1)
        coef = {0: 0, 1: 0, 2: 0}
        for w, c in re.findall(r"__dp4a\(w(\d), (0x[0-9a-fA-F]{8})u", expr):
            assert coef[int(w)] == 0, "word w%s used twice" % w
            coef[int(w)] = int(c, 16)
        window = _bytes(coef[0]) + _bytes(coef[1]) + _bytes(coef[2])
        want = [0] * 12
        want[j + 1:j + 8] = TAPS
        assert window == want, (comp, window)
        assert expr.count("128u") == 1, "every horizontal sum starts from 128 exactly once (256 * 128 = the vertical rounding constant)"


def test_vertical_pass_coefficient_pairs():
    """An output row is four IDP.2A of vertical row pairs (rows 2p, 2p + 1 in one word) with coefficient byte pairs: even rows use
    the .lo halves of C0..C3 on the pairs that start at their own row (taps r .. r + 6, row r + 7 unused), odd rows the .hi halves
    on the pairs that start one row above (row r - 1 unused, taps r .. r + 6)."""
    words = [int(x, 16) for x in re.findall(r"\bC[0-3] = (0x[0-9a-fA-F]{8})u", SRC)]
    assert len(words) == 4
    even, odd = [], []
    for c in words:
        b = _bytes(c)
        even += b[0:2]
        odd += b[2:4]
    assert even == TAPS + [0]
    assert odd == [0] + TAPS


def test_pair_formulation_equals_seven_tap_sum():
    """The arithmetic of the two passes restated with numpy on random bytes: horizontal sums of 16 bits (+128), stored as vertical
    pairs, four 2-way dot products per output, byte 2 of the accumulator = (sum + 32768) >> 16 of the plain separable filter."""
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, size=(22, 40), dtype=np.int64)
    k = np.array(TAPS, dtype=np.int64)
    h = np.zeros((22, 34), dtype=np.int64)
    for x in range(34):
        h[:, x] = (img[:, x:x + 7] * k).sum(axis=1) + 128
    assert h.max() < 65536
    words = [int(x, 16) for x in re.findall(r"\bC[0-3] = (0x[0-9a-fA-F]{8})u", SRC)]
    pairs = h[0::2] | (h[1::2] << 16)                      # word = row 2p | row 2p + 1 << 16
    for r in range(16 - 1):
        m, hi = r >> 1, r & 1
        acc = np.zeros(34, dtype=np.int64)
        for t in range(4):
            c = _bytes(words[t])[2 * hi:2 * hi + 2]
            p = pairs[m + t]
            acc += (p & 0xFFFF) * c[0] + (p >> 16) * c[1]
        ref = np.zeros(34, dtype=np.int64)
        for x in range(34):
            ref[x] = int((img[r:r + 7, x:x + 7] * np.outer(k, k)).sum())
        assert np.array_equal((acc >> 16) & 0xFF, (ref + 32768) >> 16)
        assert acc.max() < (1 << 24)
