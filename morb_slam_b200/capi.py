"""ctypes binding of liborb_b200.so (include/orb_b200.h) plus thin Python mirrors of the reference's
call surface for the hot path:

  ORBextractor.__call__(image, lapping)   <-> ORBextractor::operator() (include/ORBextractor.h:56-58)
  ORBextractor.extract_batch(...)          batched form (independent frames)
  compute_stereo_matches(exL, exR, ...)   <-> Frame::ComputeStereoMatches (src/Frame.cc:889-1047)
  hamming_knn2 / knn2_merge / ratio_test  <-> BFMatcher.knnMatch + Lowe test (src/Frame.cc:1242-1250)
  descriptor_distance                     <-> ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:1880-1894)

There is no CPU fallback: importing works anywhere, but every compute call needs the CUDA library
and a CUDA device, and raises otherwise.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ORB_B200_LIB") or os.path.join(HERE, "lib", "liborb_b200.so")   # ORB_B200_LIB: a variant build for A/B measurements

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

ORB_SRC_DEVICE, ORB_DST_DEVICE, ORB_ASYNC, ORB_NO_OUTPUT, ORB_INPUT_REMAP, ORB_INPUT_RESIZE = 1, 2, 4, 8, 16, 32
ORB_ERR_EMPTY_IMAGE = -1
ORB_ERR_INVALID_ARG, ORB_ERR_CUDA, ORB_ERR_UNSUPPORTED_SIZE, ORB_ERR_CAPACITY, ORB_ERR_STATE = -2, -3, -4, -5, -6


class OrbError(RuntimeError):
    def __init__(self, status, msg=""):
        super().__init__("orb_b200 status %d (%s) %s" % (status, _status_string(status), msg))
        self.status = status


class _Params(C.Structure):
    _fields_ = [("nfeatures", C.c_int), ("scale_factor", C.c_float), ("nlevels", C.c_int),
                ("ini_th_fast", C.c_int), ("min_th_fast", C.c_int)]


class _Kb8Rig(C.Structure):
    # orb_kb8_rig (include/orb_b200.h)
    _fields_ = [("cam1", C.c_float * 8), ("cam2", C.c_float * 8), ("precision1", C.c_float), ("precision2", C.c_float),
                ("R12", C.c_float * 9), ("t12", C.c_float * 3)]


def kb8_rig(rig):
    """orb_kb8_rig from dict(cam1, cam2, prec1, prec2, R12, t12) (synth.kb8_rig)"""
    r = _Kb8Rig()
    r.cam1[:] = [float(v) for v in np.asarray(rig["cam1"], np.float32)]
    r.cam2[:] = [float(v) for v in np.asarray(rig["cam2"], np.float32)]
    r.precision1, r.precision2 = float(rig["prec1"]), float(rig["prec2"])
    r.R12[:] = [float(v) for v in np.asarray(rig["R12"], np.float32).reshape(9)]
    r.t12[:] = [float(v) for v in np.asarray(rig["t12"], np.float32)]
    return r


_lib = None


def lib():
    """Load liborb_b200.so; raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("liborb_b200.so is missing (%s): run `python -c 'import __graft_entry__ as g; g.build()'`"
                           % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i, f, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    ip = C.POINTER(C.c_int)
    L.orb_create.argtypes = [C.POINTER(_Params), i, i, i, i, C.POINTER(vp)]
    L.orb_destroy.argtypes = [vp]
    L.orb_last_error.argtypes = [vp]
    L.orb_last_error.restype = C.c_char_p
    L.orb_status_string.argtypes = [i]
    L.orb_status_string.restype = C.c_char_p
    L.orb_keypoint_capacity.argtypes = [vp]
    L.orb_get_tables.argtypes = [vp, vp, vp, vp, vp, vp]
    L.orb_compute_tables.argtypes = [vp, vp, vp, vp, vp, vp]
    L.orb_extract.argtypes = [vp, vp, i, i, sz, i, i, vp, vp, i, ip, ip]
    L.orb_extract_batch.argtypes = [vp, vp, i, i, i, sz, sz, i, i, vp, vp, i, vp, vp, i]
    L.orb_sync.argtypes = [vp]
    L.orb_set_rectify_maps.argtypes = [vp, vp, vp, i, i]
    L.orb_set_input_size.argtypes = [vp, i, i]
    L.orb_pyramid_level_size.argtypes = [vp, i, ip, ip]
    L.orb_pyramid_level.argtypes = [vp, i, i, vp, sz]
    L.orb_stereo_match_batch.argtypes = [vp, vp, f, f, vp, vp, i, i]
    L.orb_stereo_match.argtypes = [vp, vp, vp, vp, i, vp, vp, i, f, f, vp, vp]
    L.orb_stereo_fisheye_match_batch.argtypes = [vp, vp, vp, vp, vp, i, i]
    L.orb_stereo_fisheye_triangulate_batch.argtypes = [vp, vp, C.POINTER(_Kb8Rig), vp, vp, vp, vp, vp, i, i]
    L.orb_kb8_triangulate_matches.argtypes = [vp, C.POINTER(_Kb8Rig), vp, vp, vp, vp, i, vp, vp]
    L.orb_hamming_knn2.argtypes = [vp, vp, i, vp, C.c_int64, C.c_int32, vp, vp, i]
    L.orb_knn2_merge.argtypes = [vp, vp, vp, i, i, vp, vp, i]
    L.orb_knn_exchange_create.argtypes = [vp, i, i, i, C.POINTER(vp), vp]
    L.orb_knn_exchange_connect.argtypes = [vp, vp]
    L.orb_knn_exchange_connect_local.argtypes = [vp, C.POINTER(vp)]
    L.orb_knn_exchange_destroy.argtypes = [vp]
    L.orb_knn_exchange_check.argtypes = [vp]
    L.orb_hamming_knn2_sharded.argtypes = [vp, vp, vp, i, vp, C.c_int64, C.c_int32, vp, vp, i]
    L.orb_ratio_test.argtypes = [vp, vp, i, vp, i]
    L.orb_hamming_distance.argtypes = [vp, vp]
    L.orb_serialized_keypoints_size.argtypes = [i]
    L.orb_serialized_keypoints_size.restype = sz
    L.orb_serialized_descriptors_size.argtypes = [i]
    L.orb_serialized_descriptors_size.restype = sz
    L.orb_serialize_frame.argtypes = [vp, i, i, vp, sz, C.POINTER(sz)]
    L.orb_deserialize_keypoints.argtypes = [vp, sz, vp, i, ip]
    L.orb_deserialize_descriptors.argtypes = [vp, sz, vp, i, ip]
    L.orb_host_alloc.argtypes = [C.POINTER(vp), sz]
    L.orb_host_alloc_ex.argtypes = [C.POINTER(vp), sz, i]
    L.orb_host_free.argtypes = [vp]
    L.orb_device_alloc.argtypes = [vp, C.POINTER(vp), sz]
    L.orb_device_free.argtypes = [vp, vp]
    L.orb_memcpy_h2d.argtypes = [vp, vp, vp, sz]
    L.orb_memcpy_d2h.argtypes = [vp, vp, vp, sz]
    L.orb_memcpy_h2d_async.argtypes = [vp, vp, vp, sz]
    L.orb_memcpy_d2h_async.argtypes = [vp, vp, vp, sz]
    L.orb_timer_start.argtypes = [vp]
    L.orb_timer_stop.argtypes = [vp, C.POINTER(f)]
    L.orb_launch_count.argtypes = [vp]
    L.orb_launch_count.restype = C.c_int64
    L.orb_set_stage_timing.argtypes = [vp, i]
    L.orb_get_stage_times.argtypes = [vp, vp]
    L.orb_debug_get_blurred.argtypes = [vp, i, i, vp, sz]
    L.orb_debug_get_candidates.argtypes = [vp, i, i, vp, i, ip]
    L.orb_debug_get_level_counts.argtypes = [vp, vp, i]
    L.orb_debug_get_selected.argtypes = [vp, i, i, vp, i, ip]
    L.orb_debug_distribute.argtypes = [vp, vp, i, i, i, i, vp, i, ip]
    L.orb_debug_get_stereo_best.argtypes = [vp, i, vp, vp, i]
    L.orb_assign_features_to_grid.argtypes = [vp, vp, i]
    L.orb_undistort_keypoints.argtypes = [vp, vp, vp, i, vp, vp, i, i]
    L.orb_debug_get_grid.argtypes = [vp, i, vp, vp, i, ip]
    L.orb_search_by_projection.argtypes = [vp, vp, vp, vp, i, f, i, vp, f, f, i, vp, vp, i]
    L.orb_search_local_points.argtypes = [vp, vp, vp, vp, i, vp, f, f, vp, vp, i]
    L.orb_search_by_projection_stereo.argtypes = [vp, vp, vp, vp, vp, i, f, i, vp, f, i, vp, vp, vp, i]
    L.orb_search_local_points_stereo.argtypes = [vp, vp, vp, vp, vp, i, vp, vp, vp, vp, f, f, vp, vp, vp, i]
    L.orb_vocab_create.argtypes = [i, i, i, i, i, i, vp, vp, vp, vp, C.POINTER(vp)]
    L.orb_vocab_load_text.argtypes = [i, C.c_char_p, C.POINTER(vp)]
    L.orb_vocab_info.argtypes = [vp, vp]
    L.orb_vocab_destroy.argtypes = [vp]
    L.orb_compute_bow.argtypes = [vp, vp, i, vp, i]
    L.orb_search_by_bow.argtypes = [vp, vp, f, i, vp, vp, i]
    _lib = L
    return L


def _status_string(s):
    try:
        return lib().orb_status_string(s).decode()
    except Exception:
        return "?"


def _p(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def pinned_empty(shape, dtype, write_combined=False):
    """numpy array over page-locked host memory (orb_host_alloc); keep a reference to free it later.
    write_combined: cudaHostAllocWriteCombined, for input buffers the CPU only writes."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    ptr = C.c_void_p()
    st = lib().orb_host_alloc_ex(C.byref(ptr), max(n, 1), 1 if write_combined else 0)
    if st:
        raise OrbError(st)
    buf = (C.c_uint8 * max(n, 1)).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    return arr


def compute_tables(nfeatures, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
    """orb_compute_tables: the constructor tables without a handle or a device (host arithmetic of src/ORBextractor.cc:413-443)."""
    prm = _Params(nfeatures, scale_factor, nlevels, ini_th, min_th)
    sc, inv, s2, is2 = (np.zeros(nlevels, np.float32) for _ in range(4))
    nf = np.zeros(nlevels, np.int32)
    st = lib().orb_compute_tables(C.byref(prm), _p(sc), _p(inv), _p(s2), _p(is2), _p(nf))
    if st != 0:
        raise OrbError(st, "orb_compute_tables")
    return sc, inv, s2, is2, nf


class ORBextractor:
    """Mirror of ORB_SLAM3::ORBextractor (include/ORBextractor.h:44-105) on top of the C ABI."""

    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7, max_width=752,
                 max_height=480, max_batch=1, device=0):
        self.L = lib()
        self.h = C.c_void_p()
        self.nlevels = nlevels
        self.nfeatures = nfeatures
        p = _Params(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
        st = self.L.orb_create(C.byref(p), max_width, max_height, max_batch, device, C.byref(self.h))
        if st:
            self.h = C.c_void_p()
            raise OrbError(st, "orb_create failed (no CUDA device, or unsupported parameters)")
        self.kcap = self.L.orb_keypoint_capacity(self.h)
        self.device = device
        self.cur_batch = 0   # frames of the last extraction (what the device-resident consumers work on)

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.orb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st:
            raise OrbError(st, self.L.orb_last_error(self.h).decode())

    # ---- getters (include/ORBextractor.h:60-74)
    def GetLevels(self):
        return self.nlevels

    def tables(self):
        n = self.nlevels
        sc, inv, s2, is2 = (np.zeros(n, np.float32) for _ in range(4))
        nf = np.zeros(n, np.int32)
        self._check(self.L.orb_get_tables(self.h, _p(sc), _p(inv), _p(s2), _p(is2), _p(nf)))
        return dict(scale=sc, inv_scale=inv, sigma2=s2, inv_sigma2=is2, nfeat=nf)

    def GetScaleFactors(self):
        return self.tables()["scale"]

    def GetInverseScaleFactors(self):
        return self.tables()["inv_scale"]

    def GetScaleSigmaSquares(self):
        return self.tables()["sigma2"]

    def GetInverseScaleSigmaSquares(self):
        return self.tables()["inv_sigma2"]

    # ---- operator()
    def __call__(self, image, lapping=(0, 0)):
        """Returns (monoIndex, keypoints[KP_DTYPE], descriptors[K, 32]); monoIndex = -1 for an empty image."""
        kps = np.zeros(self.kcap, KP_DTYPE)
        desc = np.zeros((self.kcap, 32), np.uint8)
        n, mono = C.c_int(0), C.c_int(0)
        if image is None or image.size == 0:
            st = self.L.orb_extract(self.h, None, 0, 0, 0, lapping[0], lapping[1], _p(kps), _p(desc), self.kcap,
                                    C.byref(n), C.byref(mono))
            assert st == ORB_ERR_EMPTY_IMAGE
            return -1, kps[:0], desc[:0]
        assert image.dtype == np.uint8 and image.ndim == 2 and image.strides[1] == 1
        st = self.L.orb_extract(self.h, _p(image), image.shape[1], image.shape[0], image.strides[0], lapping[0],
                                lapping[1], _p(kps), _p(desc), self.kcap, C.byref(n), C.byref(mono))
        self._check(st)
        self.cur_batch = 1
        return mono.value, kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch(self, images, lapping=(0, 0), out=None, flags=0):
        """images: uint8 [B, H, W] (host array, or a (device_ptr, B, H, W) tuple with ORB_SRC_DEVICE).
        Returns (n[B], mono[B], kps[B, kcap], desc[B, kcap, 32]); out may supply preallocated (pinned) arrays."""
        if isinstance(images, tuple):
            ptr, B, H, W = images
            src = C.c_void_p(ptr)
            stride, istride = W, W * H
            flags |= ORB_SRC_DEVICE
        else:
            assert images.dtype == np.uint8 and images.ndim == 3 and images.strides[2] == 1
            B, H, W = images.shape
            src = _p(images)
            stride, istride = images.strides[1], images.strides[0]
        if out is None:
            out = (np.zeros(B, np.int32), np.zeros(B, np.int32), np.zeros((B, self.kcap), KP_DTYPE),
                   np.zeros((B, self.kcap, 32), np.uint8))
        n, mono, kps, desc = out
        if flags & ORB_NO_OUTPUT:
            st = self.L.orb_extract_batch(self.h, src, B, W, H, stride, istride, lapping[0], lapping[1], None, None, 0,
                                          _p(n), _p(mono), flags)
        else:
            st = self.L.orb_extract_batch(self.h, src, B, W, H, stride, istride, lapping[0], lapping[1], _p(kps),
                                          _p(desc), kps.shape[1], _p(n), _p(mono), flags)
        self._check(st)
        self.cur_batch = B
        return out

    def sync(self):
        self._check(self.L.orb_sync(self.h))

    def set_input_size(self, new_w, new_h):
        """settings_->newImSize() (needToResize, src/System.cc:262-264): extract_batch(..., flags=ORB_INPUT_RESIZE) then resizes
        the frames it is given to new_w x new_h on the device like cv::resize (INTER_LINEAR). (0, 0) clears."""
        self._check(self.L.orb_set_input_size(self.h, int(new_w), int(new_h)))

    def set_rectify_maps(self, map_x, map_y):
        """M1 / M2 of cv::initUndistortRectifyMap (CV_32FC1, src/Settings.cc:540-545); extract_batch(..., flags=ORB_INPUT_REMAP)
        then takes raw frames and rectifies them on the device like System::TrackStereo's cv::remap. None clears."""
        if map_x is None:
            self._check(self.L.orb_set_rectify_maps(self.h, None, None, 0, 0))
            return
        map_x = np.ascontiguousarray(map_x, np.float32); map_y = np.ascontiguousarray(map_y, np.float32)
        assert map_x.shape == map_y.shape and map_x.ndim == 2
        self._check(self.L.orb_set_rectify_maps(self.h, _p(map_x), _p(map_y), map_x.shape[1], map_x.shape[0]))

    # ---- mvImagePyramid
    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        self._check(self.L.orb_pyramid_level_size(self.h, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    def pyramid_level(self, level, frame=0):
        w, h = self.level_size(level)
        out = np.zeros((h, w), np.uint8)
        self._check(self.L.orb_pyramid_level(self.h, frame, level, _p(out), w))
        return out

    @property
    def mvImagePyramid(self):
        return [self.pyramid_level(l) for l in range(self.nlevels)]

    # ---- stage outputs (parity tests)
    def blurred_level(self, level, frame=0):
        w, h = self.level_size(level)
        out = np.zeros((h, w), np.uint8)
        self._check(self.L.orb_debug_get_blurred(self.h, frame, level, _p(out), w))
        return out

    def candidates(self, level, frame=0, cap=1 << 16):
        out = np.zeros((cap, 3), np.int32)
        n = C.c_int()
        self._check(self.L.orb_debug_get_candidates(self.h, frame, level, _p(out), cap, C.byref(n)))
        return out[:n.value].copy()

    def level_counts(self, batch):
        out = np.zeros((batch, self.nlevels), np.int32)
        self._check(self.L.orb_debug_get_level_counts(self.h, _p(out), out.size))
        return out

    def selected(self, level, frame=0, cap=1 << 14):
        out = np.zeros((cap, 3), np.int32)
        n = C.c_int()
        self._check(self.L.orb_debug_get_selected(self.h, frame, level, _p(out), cap, C.byref(n)))
        return out[:n.value].copy()

    def distribute(self, cands_xys, w, h, N):
        c = np.ascontiguousarray(cands_xys, np.int32)
        cap = N + 64
        out = np.zeros((cap, 3), np.int32)
        n = C.c_int()
        self._check(self.L.orb_debug_distribute(self.h, _p(c), len(c), w, h, N, _p(out), cap, C.byref(n)))
        return out[:n.value].copy()

    def std_sort(self, keys):
        """the quad-tree kernel's std::sort emulation on (key, input index) records: returns (sorted keys, payload order)"""
        k = np.ascontiguousarray(keys, np.uint32)
        ko = np.zeros(len(k), np.uint32); po = np.zeros(len(k), np.uint32)
        self.L.orb_debug_std_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        self._check(self.L.orb_debug_std_sort(self.h, _p(k), len(k), _p(ko), _p(po)))
        return ko, po

    # ---- measurement
    def timer_start(self):
        self._check(self.L.orb_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        self._check(self.L.orb_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(self.L.orb_launch_count(self.h))

    def set_stage_timing(self, on):
        self._check(self.L.orb_set_stage_timing(self.h, int(on)))

    def stage_times(self):
        ms = np.zeros(8, np.float32)
        self._check(self.L.orb_get_stage_times(self.h, _p(ms)))
        return ms


def compute_stereo_matches_batch(exL, exR, mbf, maxD, out=None, flags=0):
    """Frame::ComputeStereoMatches for every frame of the two extractors' last batches.
    Returns (uRight[B, kcap], depth[B, kcap]); entries beyond a frame's keypoint count are undefined."""
    B = None
    if out is None:
        raise ValueError("pass out=(uright, depth) arrays of shape [B, kcap]")
    uR, dp = out
    cap = uR.shape[1] if uR is not None else 0
    st = exL.L.orb_stereo_match_batch(exL.h, exR.h, float(mbf), float(maxD), _p(uR), _p(dp), cap, flags)
    exL._check(st)
    return out


def compute_stereo_fisheye_matches_batch(exL, exR, flags=0, want=True):
    """Frame::ComputeStereoFishEyeMatches up to the triangulation for every frame of the two extractors' last batches: returns
    (idx[B, kcap, 2], dist[B, kcap, 2], passed[B, kcap]) indexed by query i = left keypoint monoLeft + i."""
    if not want:
        exL._check(exL.L.orb_stereo_fisheye_match_batch(exL.h, exR.h, None, None, None, 0, flags | ORB_NO_OUTPUT))
        return None
    B, k = exL.cur_batch, exL.kcap
    idx = np.full((B, k, 2), -1, np.int32); dist = np.full((B, k, 2), -1, np.int32); ok = np.zeros((B, k), np.uint8)
    exL._check(exL.L.orb_stereo_fisheye_match_batch(exL.h, exR.h, _p(idx), _p(dist), _p(ok), k, flags))
    return idx, dist, ok


ORB_SER_KEYS, ORB_SER_KEYS_UN, ORB_SER_DESCRIPTORS = 0, 1, 2


def serialize_frame(ex, frame, what):
    """The Atlas fragment of frame `frame` of the extractor's last batch (include/orb_b200.h: orb_serialize_frame): bytes as
    serializeVectorKeyPoints(mvKeys / mvKeysUn) or serializeMatrix(mDescriptors) write them into a binary archive."""
    cap = max(int(ex.L.orb_serialized_keypoints_size(ex.kcap)), int(ex.L.orb_serialized_descriptors_size(ex.kcap)))
    out = np.zeros(cap, np.uint8)
    n = C.c_size_t(0)
    ex._check(ex.L.orb_serialize_frame(ex.h, int(frame), int(what), _p(out), cap, C.byref(n)))
    return out[:n.value].tobytes()


def deserialize_keypoints(buf, cap=None):
    """host-only loader of a keypoint fragment -> KP_DTYPE records"""
    b = np.frombuffer(bytes(buf), np.uint8).copy()
    n = C.c_int(0)
    if cap is None:
        cap = max((len(b) - 4) // 28, 0)
    kps = np.zeros(max(cap, 1), KP_DTYPE)
    st = lib().orb_deserialize_keypoints(_p(b), len(b), _p(kps), cap, C.byref(n))
    if st != 0:
        raise OrbError(st)
    return kps[:n.value]


def deserialize_descriptors(buf, cap_rows=None):
    b = np.frombuffer(bytes(buf), np.uint8).copy()
    n = C.c_int(0)
    if cap_rows is None:
        cap_rows = max((len(b) - 13) // 32, 0)
    d = np.zeros((max(cap_rows, 1), 32), np.uint8)
    st = lib().orb_deserialize_descriptors(_p(b), len(b), _p(d), cap_rows, C.byref(n))
    if st != 0:
        raise OrbError(st)
    return d[:n.value]


def compute_stereo_fisheye_triangulation_batch(exL, exR, rig, flags=0, want=True, pinned=False):
    """The rest of Frame::ComputeStereoFishEyeMatches (src/Frame.cc:1244-1273) on the device results of
    compute_stereo_fisheye_matches_batch: returns (mvLeftToRightMatch[B, kcap], mvRightToLeftMatch[B, kcap], mvDepth[B, kcap],
    mvStereo3Dpoints[B, kcap, 3], code[B, kcap]); entries beyond a frame's keypoint count are -1 / -1 / -1 / 0 / 0."""
    r = rig if isinstance(rig, _Kb8Rig) else kb8_rig(rig)
    if not want:
        exL._check(exL.L.orb_stereo_fisheye_triangulate_batch(exL.h, exR.h, C.byref(r), None, None, None, None, None, 0, flags | ORB_NO_OUTPUT))
        return None
    B, k = exL.cur_batch, max(exL.kcap, exR.kcap)
    if pinned:   # page-locked result buffers: a few frames get them written by the kernel itself
        l2r = pinned_empty((B, k), np.int32); r2l = pinned_empty((B, k), np.int32); depth = pinned_empty((B, k), np.float32)
        p3d = pinned_empty((B, k, 3), np.float32); code = pinned_empty((B, k), np.int8)
        l2r[:] = -1; r2l[:] = -1; depth[:] = -1; p3d[:] = 0; code[:] = 0
    else:
        l2r = np.full((B, k), -1, np.int32); r2l = np.full((B, k), -1, np.int32); depth = np.full((B, k), -1, np.float32)
        p3d = np.zeros((B, k, 3), np.float32); code = np.zeros((B, k), np.int8)
    exL._check(exL.L.orb_stereo_fisheye_triangulate_batch(exL.h, exR.h, C.byref(r), _p(l2r), _p(r2l), _p(depth), _p(p3d), _p(code), k, flags))
    return l2r, r2l, depth, p3d, code


def kb8_triangulate_matches(ex, rig, xy1, xy2, sigma1, sigma2):
    """KannalaBrandt8::TriangulateMatches (src/CameraModels/KannalaBrandt8.cpp:323-395) on explicit keypoint pairs: (ret[n], p3d[n, 3])"""
    r = rig if isinstance(rig, _Kb8Rig) else kb8_rig(rig)
    xy1 = np.ascontiguousarray(xy1, np.float32); xy2 = np.ascontiguousarray(xy2, np.float32)
    s1 = np.ascontiguousarray(sigma1, np.float32); s2 = np.ascontiguousarray(sigma2, np.float32)
    n = len(s1)
    ret = np.zeros(max(n, 1), np.float32); p3d = np.zeros((max(n, 1), 3), np.float32)
    ex._check(ex.L.orb_kb8_triangulate_matches(ex.h, C.byref(r), _p(xy1), _p(xy2), _p(s1), _p(s2), n, _p(ret), _p(p3d)))
    return ret[:n], p3d[:n]


def compute_stereo_matches(exL, exR, kpsL, descL, kpsR, descR, mbf, maxD):
    """Single-frame form with host keypoints / descriptors (mvKeys, mDescriptors, ...): returns
    (mvuRight, mvDepth). Pyramids are those of the extractors' last call."""
    nL, nR = len(kpsL), len(kpsR)
    kpsL = np.ascontiguousarray(kpsL); kpsR = np.ascontiguousarray(kpsR)
    descL = np.ascontiguousarray(descL); descR = np.ascontiguousarray(descR)
    uR = np.full(max(nL, 1), -1, np.float32)
    dp = np.full(max(nL, 1), -1, np.float32)
    st = exL.L.orb_stereo_match(exL.h, exR.h, _p(kpsL), _p(descL), nL, _p(kpsR), _p(descR), nR, float(mbf),
                                float(maxD), _p(uR), _p(dp))
    exL._check(st)
    return uR[:nL], dp[:nL]


def stereo_best(exL, frame=0):
    bi = np.zeros(exL.kcap, np.int32)
    bd = np.zeros(exL.kcap, np.int32)
    exL._check(exL.L.orb_debug_get_stereo_best(exL.h, frame, _p(bi), _p(bd), exL.kcap))
    return bi, bd


def hamming_knn2(ex, q, db, index_base=0, flags=0, ndb=None, nq=None, out=None):
    """Top-2 Hamming neighbours of every query row in db. Host arrays, or raw device pointers (ints)
    with the ORB_*_DEVICE flags. Returns (idx[nq, 2], dist[nq, 2]) int32."""
    if not isinstance(q, int):
        q = np.ascontiguousarray(q, np.uint8)
        nq = len(q)
    if not isinstance(db, int):
        db = np.ascontiguousarray(db, np.uint8)
        ndb = len(db)
    if out is None:
        out = (np.zeros((nq, 2), np.int32), np.zeros((nq, 2), np.int32))
    idx, dist = out
    st = ex.L.orb_hamming_knn2(ex.h, _p(q), nq, _p(db), ndb, index_base, _p(idx), _p(dist), flags)
    ex._check(st)
    return out


ORB_IPC_HANDLE_BYTES = 64


class KnnExchange:
    """orb_knn_exchange (include/orb_b200.h): this rank's peer-writable buffer of the sharded kNN. `handle` is the CUDA IPC handle
    the other ranks need (bytes); connect(all_handles) maps theirs, connect_local(peers) is for ranks inside one process."""

    def __init__(self, ex, rank, world, max_nq):
        self.ex, self.rank, self.world, self.max_nq = ex, rank, world, max_nq
        self.x = C.c_void_p()
        hb = np.zeros(ORB_IPC_HANDLE_BYTES, np.uint8)
        ex._check(ex.L.orb_knn_exchange_create(ex.h, rank, world, max_nq, C.byref(self.x), _p(hb)))
        self.handle = hb

    def connect(self, all_handles):
        a = np.ascontiguousarray(all_handles, np.uint8).reshape(self.world, ORB_IPC_HANDLE_BYTES)
        self.ex._check(self.ex.L.orb_knn_exchange_connect(self.x, _p(a)))

    def connect_local(self, peers):
        arr = (C.c_void_p * self.world)(*[p.x for p in peers])
        self.ex._check(self.ex.L.orb_knn_exchange_connect_local(self.x, arr))

    def search(self, q_ptr, nq, db_ptr, ndb, index_base, idx_ptr, dist_ptr, flags=0):
        """orb_hamming_knn2_sharded with raw device pointers (collective: every rank calls it)"""
        self.ex._check(self.ex.L.orb_hamming_knn2_sharded(self.ex.h, self.x, q_ptr, nq, db_ptr, ndb, index_base, idx_ptr, dist_ptr,
                                                          flags | ORB_SRC_DEVICE | ORB_DST_DEVICE))

    def check(self):
        """completes the ORB_ASYNC searches; raises if one of them ran into the peer time-out"""
        self.ex._check(self.ex.L.orb_knn_exchange_check(self.x))

    def close(self):
        if self.x:
            self.ex.L.orb_knn_exchange_destroy(self.x)
            self.x = C.c_void_p()


def knn2_merge(ex, idx_parts, dist_parts, flags=0, nparts=None, nq=None, out=None):
    if not isinstance(idx_parts, int):
        idx_parts = np.ascontiguousarray(idx_parts, np.int32)
        dist_parts = np.ascontiguousarray(dist_parts, np.int32)
        nparts, nq = idx_parts.shape[0], idx_parts.shape[1]
    if out is None:
        out = (np.zeros((nq, 2), np.int32), np.zeros((nq, 2), np.int32))
    st = ex.L.orb_knn2_merge(ex.h, _p(idx_parts), _p(dist_parts), nparts, nq, _p(out[0]), _p(out[1]), flags)
    ex._check(st)
    return out


def ratio_test(ex, dist):
    d = np.ascontiguousarray(dist, np.int32)
    out = np.zeros(len(d), np.uint8)
    ex._check(ex.L.orb_ratio_test(ex.h, _p(d), len(d), _p(out), 0))
    return out.astype(bool)


def descriptor_distance(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return lib().orb_hamming_distance(_p(a), _p(b))


# ---- windowed matcher (include/orb_b200.h: orb_assign_features_to_grid, orb_search_by_projection) ----
GRID_COLS, GRID_ROWS = 64, 48
Q_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("z", "<f4"), ("angle", "<f4"), ("octave", "<i4"), ("flags", "<i4")])  # orb_proj_query


def grid_params(w, h):
    """orb_grid_params of an undistorted w x h image: mnMinX, mnMinY, mnMaxX, mnMaxY and the inverse cell sizes as
    the reference computes them (src/Frame.cc:236-241)."""
    minx, miny, maxx, maxy = np.float32(0), np.float32(0), np.float32(w), np.float32(h)
    return np.array([minx, miny, maxx, maxy, np.float32(GRID_COLS) / (maxx - minx), np.float32(GRID_ROWS) / (maxy - miny)],
                    dtype=np.float32)


def undistort_keypoints(ex, K, dist, P=None, flags=0):
    """Frame::UndistortKeyPoints for every frame of the extractor's last batch: K = toK(), dist = mDistCoef, P = mK (3 x 3 float;
    default K). Returns mvKeysUn [B, kcap]; the grid and the searches use it from now on."""
    K = np.ascontiguousarray(K, np.float32).reshape(3, 3)
    P = K if P is None else np.ascontiguousarray(P, np.float32).reshape(3, 3)
    dist = np.ascontiguousarray(dist, np.float32).ravel()
    out = np.zeros((ex.cur_batch, ex.kcap), KP_DTYPE)
    ex._check(ex.L.orb_undistort_keypoints(ex.h, _p(K), _p(dist) if len(dist) else None, len(dist), _p(P), _p(out), ex.kcap, flags))
    return out


def assign_features_to_grid(ex, gp, flags=0):
    """Frame::AssignFeaturesToGrid for every frame of the extractor's last batch (device-resident)."""
    gp = np.ascontiguousarray(gp, dtype=np.float32)
    ex._check(ex.L.orb_assign_features_to_grid(ex.h, _p(gp), flags))


def get_grid(ex, frame=0):
    off = np.zeros(GRID_COLS * GRID_ROWS + 1, np.int32)
    idx = np.zeros(ex.kcap, np.int32)
    n = C.c_int(0)
    ex._check(ex.L.orb_debug_get_grid(ex.h, frame, _p(off), _p(idx), ex.kcap, C.byref(n)))
    return off, idx[:n.value]


def search_by_projection(ex, queries, qdesc, nq, th, mono, tlc_z, mb, mbf, check_orientation=True, out=None, flags=0):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) for every frame of the extractor's last batch.
    queries: Q_DTYPE [B, qcap], qdesc: uint8 [B, qcap, 32], nq: int32 [B], tlc_z: float32 [B] (host arrays, or device
    pointers as ints with ORB_SRC_DEVICE plus out=(match_ptr, nmatches_ptr) with ORB_DST_DEVICE).
    Returns (nmatches[B], match[B, kcap])."""
    if flags & ORB_SRC_DEVICE:
        q_p, d_p, n_p, t_p, B, qcap = queries
        args = (C.c_void_p(q_p), C.c_void_p(d_p), C.c_void_p(n_p), qcap)
        tz = C.c_void_p(t_p)
    else:
        queries = np.ascontiguousarray(queries, dtype=Q_DTYPE)
        qdesc = np.ascontiguousarray(qdesc, dtype=np.uint8)
        nq = np.ascontiguousarray(nq, dtype=np.int32)
        tlc_z = np.ascontiguousarray(tlc_z, dtype=np.float32)
        B, qcap = queries.shape
        args = (_p(queries), _p(qdesc), _p(nq), qcap)
        tz = _p(tlc_z)
    if out is None:
        out = (np.zeros(B, np.int32), np.full((B, ex.kcap), -1, np.int32))
    nm, match = out
    mp = C.c_void_p(match) if isinstance(match, int) else _p(match)
    np_ = C.c_void_p(nm) if isinstance(nm, int) else _p(nm)
    ex._check(ex.L.orb_search_by_projection(ex.h, args[0], args[1], args[2], args[3], float(th), int(mono), tz, float(mb), float(mbf),
                                            int(check_orientation), mp, np_, flags))
    return out


TQ_DTYPE = np.dtype([("proj_x", "<f4"), ("proj_y", "<f4"), ("proj_xr", "<f4"), ("view_cos", "<f4"), ("level", "<i4"),
                     ("flags", "<i4")])  # orb_track_query


def search_local_points(ex, queries, qdesc, nq, locked0, th, nnratio=0.8, out=None, flags=0):
    """ORBmatcher::SearchByProjection(F, vpMapPoints, th) (the local-map search) for every frame of the extractor's last
    batch. queries: TQ_DTYPE [B, qcap], qdesc: uint8 [B, qcap, 32], nq: int32 [B], locked0: uint8 [B, kcap] or None
    (host arrays, or (q_ptr, desc_ptr, nq_ptr, locked_ptr_or_0, B, qcap) device pointers with ORB_SRC_DEVICE plus
    out=(nmatches_ptr, match_ptr) with ORB_DST_DEVICE). Returns (nmatches[B], match[B, kcap])."""
    if flags & ORB_SRC_DEVICE:
        q_p, d_p, n_p, l_p, B, qcap = queries
        args = (C.c_void_p(q_p), C.c_void_p(d_p), C.c_void_p(n_p), qcap, C.c_void_p(l_p) if l_p else None)
    else:
        queries = np.ascontiguousarray(queries, dtype=TQ_DTYPE)
        qdesc = np.ascontiguousarray(qdesc, dtype=np.uint8)
        nq = np.ascontiguousarray(nq, dtype=np.int32)
        B, qcap = queries.shape
        if locked0 is not None:
            locked0 = np.ascontiguousarray(locked0, dtype=np.uint8)
            assert locked0.shape == (B, ex.kcap)
        args = (_p(queries), _p(qdesc), _p(nq), qcap, _p(locked0) if locked0 is not None else None)
    if out is None:
        out = (np.zeros(B, np.int32), np.full((B, ex.kcap), -1, np.int32))
    nm, match = out
    mp = C.c_void_p(match) if isinstance(match, int) else _p(match)
    np_ = C.c_void_p(nm) if isinstance(nm, int) else _p(nm)
    ex._check(ex.L.orb_search_local_points(ex.h, args[0], args[1], args[2], args[3], args[4], float(th), float(nnratio), mp, np_, flags))
    return out


# ---- two-camera frames (Nleft != -1): the right-camera halves of both searches (include/orb_b200.h) ----
Q2_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("z", "<f4"), ("angle", "<f4"), ("octave", "<i4"), ("flags", "<i4"), ("ur", "<f4"),
                     ("vr", "<f4")])  # orb_proj_query2
TQ2_DTYPE = np.dtype([("proj_x", "<f4"), ("proj_y", "<f4"), ("view_cos", "<f4"), ("level", "<i4"), ("proj_xr", "<f4"), ("proj_yr", "<f4"),
                      ("view_cos_r", "<f4"), ("level_r", "<i4"), ("flags", "<i4"), ("pad", "<i4")])  # orb_track_query2


def search_by_projection_stereo(exL, exR, queries, qdesc, nq, th, mono, tlc_z, mb, check_orientation=True, flags=0):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) with a two-camera CurrentFrame = (exL, exR).
    queries: Q2_DTYPE [B, qcap], qdesc uint8 [B, qcap, 32], nq int32 [B], tlc_z float32 [B] (host arrays).
    Returns (nmatches[B], match_left[B, kcapL], match_right[B, kcapR])."""
    queries = np.ascontiguousarray(queries, dtype=Q2_DTYPE)
    qdesc = np.ascontiguousarray(qdesc, dtype=np.uint8)
    nq = np.ascontiguousarray(nq, dtype=np.int32)
    tlc_z = np.ascontiguousarray(tlc_z, dtype=np.float32)
    B, qcap = queries.shape
    nm = np.zeros(B, np.int32)
    mL = np.full((B, exL.kcap), -1, np.int32); mR = np.full((B, exR.kcap), -1, np.int32)
    exL._check(exL.L.orb_search_by_projection_stereo(exL.h, exR.h, _p(queries), _p(qdesc), _p(nq), qcap, float(th), int(mono), _p(tlc_z),
                                                     float(mb), int(check_orientation), _p(mL), _p(mR), _p(nm), flags))
    return nm, mL, mR


def search_local_points_stereo(exL, exR, queries, qdesc, nq, locked0_left, locked0_right, left_to_right, right_to_left, th, nnratio=0.8,
                               flags=0):
    """ORBmatcher::SearchByProjection(F, vpMapPoints, th) with a two-camera F = (exL, exR). queries: TQ2_DTYPE [B, qcap];
    locked0_*: uint8 [B, kcap] or None; left_to_right int32 [B, kcapL] / right_to_left int32 [B, kcapR], or both None for the
    device-resident pairing of compute_stereo_fisheye_triangulation_batch. Returns (nmatches[B], match_left, match_right)."""
    queries = np.ascontiguousarray(queries, dtype=TQ2_DTYPE)
    qdesc = np.ascontiguousarray(qdesc, dtype=np.uint8)
    nq = np.ascontiguousarray(nq, dtype=np.int32)
    B, qcap = queries.shape
    opt = lambda a, t, shp: None if a is None else np.ascontiguousarray(a, dtype=t).reshape(shp)  # noqa: E731
    lkL = opt(locked0_left, np.uint8, (B, exL.kcap)); lkR = opt(locked0_right, np.uint8, (B, exR.kcap))
    l2r = opt(left_to_right, np.int32, (B, exL.kcap)); r2l = opt(right_to_left, np.int32, (B, exR.kcap))
    nm = np.zeros(B, np.int32)
    mL = np.full((B, exL.kcap), -1, np.int32); mR = np.full((B, exR.kcap), -1, np.int32)
    pp = lambda a: None if a is None else _p(a)  # noqa: E731
    exL._check(exL.L.orb_search_local_points_stereo(exL.h, exR.h, _p(queries), _p(qdesc), _p(nq), qcap, pp(lkL), pp(lkR), pp(l2r), pp(r2l),
                                                    float(th), float(nnratio), _p(mL), _p(mR), _p(nm), flags))
    return nm, mL, mR


# ---- bag of words (include/orb_b200.h: orb_vocab_*, orb_compute_bow) ----
class _BowOut(C.Structure):   # orb_bow_out
    _fields_ = [(n, C.c_void_p) for n in ("bow_n", "bow_word", "bow_val", "fv_n", "fv_node", "fv_off", "fv_feat", "feat_word", "feat_node")]


class ORBVocabulary:
    """Mirror of ORB_SLAM3::ORBVocabulary (include/ORBVocabulary.h) for the transform path: either the arrays of
    morb_slam_b200.synth.synth_vocabulary or a text file in the ORBvoc.txt format (loadFromTextFile)."""

    def __init__(self, voc=None, path=None, device=0):
        self.L = lib()
        self.h = C.c_void_p()
        if path is not None:
            st = self.L.orb_vocab_load_text(device, path.encode(), C.byref(self.h))
        else:
            parent = np.ascontiguousarray(voc["parent"], np.int32)
            leaf = np.ascontiguousarray(voc["is_leaf"], np.uint8)
            desc = np.ascontiguousarray(voc["desc"], np.uint8)
            weight = np.ascontiguousarray(voc["weight"], np.float64)
            st = self.L.orb_vocab_create(device, voc["k"], voc["L"], voc["scoring"], voc["weighting"], len(parent), _p(parent), _p(leaf),
                                         _p(desc), _p(weight), C.byref(self.h))
        if st:
            raise OrbError(st)

    def info(self):
        a = np.zeros(6, np.int32)
        self.L.orb_vocab_info(self.h, _p(a))
        return dict(k=int(a[0]), L=int(a[1]), scoring=int(a[2]), weighting=int(a[3]), nodes=int(a[4]), words=int(a[5]))

    def close(self):
        if self.h:
            self.L.orb_vocab_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def compute_bow(ex, voc, levelsup=4, flags=0, want=True):
    """Frame::ComputeBoW for every frame of the extractor's last batch. Returns one dict per frame with the keys of the
    oracle (bow_word, bow_val, fv_node, fv_off, fv_feat, feat_word, feat_node); want=False only launches (ORB_NO_OUTPUT)."""
    B, k = ex.cur_batch, ex.kcap
    if not want:
        ex._check(ex.L.orb_compute_bow(ex.h, voc.h, levelsup, None, flags | ORB_NO_OUTPUT))
        return None
    bn, fn = np.zeros(B, np.int32), np.zeros(B, np.int32)
    bw, bv = np.zeros((B, k), np.uint32), np.zeros((B, k), np.float64)
    fnode, foff, ffeat = np.zeros((B, k), np.uint32), np.zeros((B, k + 1), np.int32), np.zeros((B, k), np.uint32)
    fw, fnd = np.zeros((B, k), np.int32), np.zeros((B, k), np.int32)
    o = _BowOut(*[a.ctypes.data for a in (bn, bw, bv, fn, fnode, foff, ffeat, fw, fnd)])
    ex._check(ex.L.orb_compute_bow(ex.h, voc.h, levelsup, C.byref(o), flags))
    out = []
    for f in range(B):
        nb, nn = int(bn[f]), int(fn[f])
        out.append(dict(bow_word=bw[f, :nb].copy(), bow_val=bv[f, :nb].copy(), fv_node=fnode[f, :nn].copy(), fv_off=foff[f, :nn + 1].copy(),
                        fv_feat=ffeat[f, :foff[f, nn]].copy(), feat_word=fw[f], feat_node=fnd[f]))
    return out


class _BowKeyframes(C.Structure):   # orb_bow_keyframes
    _fields_ = [(n, C.c_void_p) for n in ("desc", "angle", "flags", "n", "fv_node", "fv_off", "fv_feat", "fv_n")] + [("cap", C.c_int32)]


def pack_bow_keyframes(keyframes):
    """Padded arrays of orb_bow_keyframes for a list of per-frame keyframe dicts (desc, angle, flags, fv):
    (desc, angle, flags, n, fv_node, fv_off, fv_feat, fv_n, cap)."""
    B = len(keyframes)
    cap = max([len(k["desc"]) for k in keyframes] + [1])
    desc = np.zeros((B, cap, 32), np.uint8); angle = np.zeros((B, cap), np.float32); fl = np.zeros((B, cap), np.uint8)
    n = np.zeros(B, np.int32); nn = np.zeros(B, np.int32)
    node = np.zeros((B, cap), np.uint32); off = np.zeros((B, cap + 1), np.int32); feat = np.zeros((B, cap), np.uint32)
    for i, k in enumerate(keyframes):
        m = len(k["desc"]); n[i] = m
        desc[i, :m] = k["desc"]; angle[i, :m] = k["angle"]; fl[i, :m] = k["flags"]
        fv = k["fv"]; j = len(fv["fv_node"]); nn[i] = j
        node[i, :j] = fv["fv_node"]; off[i, :j + 1] = fv["fv_off"]; feat[i, :len(fv["fv_feat"])] = fv["fv_feat"]
    return desc, angle, fl, n, node, off, feat, nn, cap


def search_by_bow(ex, keyframes, nnratio=0.7, check_orientation=True, flags=0, out=None):
    """ORBmatcher::SearchByBoW(pKF, F, vpMapPointMatches) for every frame of the extractor's last batch (orb_compute_bow must have
    run on it) against one keyframe each. keyframes: one dict per frame with desc [n, 32], angle [n], flags [n] (keypoint holds a
    good map point) and fv = dict(fv_node, fv_off, fv_feat) as compute_bow returns it - or, with ORB_SRC_DEVICE, the tuple of
    pack_bow_keyframes with device pointers in place of the arrays (and out = (nmatches_ptr, match_ptr) with ORB_DST_DEVICE).
    Returns (nmatches[B], match[B, kcap])."""
    if flags & ORB_SRC_DEVICE:
        *ptrs, cap = keyframes
        kf = _BowKeyframes(*ptrs, cap)
        B = ex.cur_batch
    else:
        *arrs, cap = pack_bow_keyframes(keyframes)
        kf = _BowKeyframes(*[a.ctypes.data for a in arrs], cap)
        B = len(keyframes)
    if out is None:
        out = (np.zeros(B, np.int32), np.full((B, ex.kcap), -1, np.int32))
    nm, match = out
    ex._check(ex.L.orb_search_by_bow(ex.h, C.byref(kf), float(nnratio), int(check_orientation), _p(match), _p(nm), flags))
    return out


# ---- two-camera frames: ComputeBoW over both cameras' descriptors and SearchByBoW against the combined frame ----
def compute_bow_stereo(exL, exR, voc, levelsup=4, flags=0):
    """Frame::ComputeBoW of a two-camera frame (features = left keypoints followed by the right ones). One dict per frame with the
    keys of compute_bow; feature indices are in the combined index space."""
    L = lib()
    if not getattr(L, "_bow2_typed", False):
        L.orb_compute_bow_stereo.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orb_search_by_bow_stereo.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L._bow2_typed = True
    B, k = exL.cur_batch, exL.kcap + exR.kcap
    bn, fn = np.zeros(B, np.int32), np.zeros(B, np.int32)
    bw, bv = np.zeros((B, k), np.uint32), np.zeros((B, k), np.float64)
    fnode, foff, ffeat = np.zeros((B, k), np.uint32), np.zeros((B, k + 1), np.int32), np.zeros((B, k), np.uint32)
    fw, fnd = np.zeros((B, k), np.int32), np.zeros((B, k), np.int32)
    o = _BowOut(*[a.ctypes.data for a in (bn, bw, bv, fn, fnode, foff, ffeat, fw, fnd)])
    exL._check(L.orb_compute_bow_stereo(exL.h, exR.h, voc.h, levelsup, C.byref(o), flags))
    out = []
    for f in range(B):
        nb, nn = int(bn[f]), int(fn[f])
        out.append(dict(bow_word=bw[f, :nb].copy(), bow_val=bv[f, :nb].copy(), fv_node=fnode[f, :nn].copy(), fv_off=foff[f, :nn + 1].copy(),
                        fv_feat=ffeat[f, :foff[f, nn]].copy(), feat_word=fw[f], feat_node=fnd[f]))
    return out


def search_by_bow_stereo(exL, exR, keyframes, nnratio=0.7, check_orientation=True, flags=0):
    """ORBmatcher::SearchByBoW(pKF, F, ...) with a two-camera F = (exL, exR) after compute_bow_stereo.
    Returns (nmatches[B], match_left[B, kcapL], match_right[B, kcapR])."""
    *arrs, cap = pack_bow_keyframes(keyframes)
    kf = _BowKeyframes(*[a.ctypes.data for a in arrs], cap)
    B = len(keyframes)
    nm = np.zeros(B, np.int32)
    mL = np.full((B, exL.kcap), -1, np.int32); mR = np.full((B, exR.kcap), -1, np.int32)
    exL._check(lib().orb_search_by_bow_stereo(exL.h, exR.h, C.byref(kf), float(nnratio), int(check_orientation), _p(mL), _p(mR), _p(nm), flags))
    return nm, mL, mR


# ---- LocalMapping / Relocalization matchers (include/orb_b200.h: orb_load_frames, orb_fuse_search, orb_search_by_projection_kf,
#      orb_search_for_triangulation, orb_distinctive_descriptors) ----
FQ_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("ur", "<f4"), ("level", "<i4"), ("flags", "<i4")])  # orb_fuse_query
assert FQ_DTYPE.itemsize == 20


def _map_lib():
    L = lib()
    if not getattr(L, "_map_typed", False):
        vp, i, f = C.c_void_p, C.c_int, C.c_float
        L.orb_load_frames.argtypes = [vp, vp, vp, vp, vp, i, i, i]
        L.orb_fuse_search.argtypes = [vp, vp, vp, vp, i, f, i, vp, vp, i]
        L.orb_search_by_projection_kf.argtypes = [vp, vp, vp, vp, i, vp, f, i, i, vp, vp, i]
        L.orb_search_for_triangulation.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, i, vp, vp, i]
        L.orb_distinctive_descriptors.argtypes = [vp, vp, vp, i, vp, vp, i]
        L.orb_search_by_bow_kf.argtypes = [vp, vp, vp, vp, i, f, i, vp, vp, i]
        L.orb_search_by_projection_sim3.argtypes = [vp, vp, vp, vp, i, vp, f, f, vp, vp, i]
        L.orb_search_for_initialization.argtypes = [vp, vp, vp, vp, i, i, f, i, vp, vp, vp, i]
        L.orb_search_for_triangulation_fisheye.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, i, vp, vp, i]
        L._map_typed = True
    return L


def load_frames(ex, kps_list, desc_list, uright_list=None):
    """Keyframes of the host's map become the extractor's resident batch (orb_load_frames): lists of KP_DTYPE keypoints (mvKeysUn),
    uint8 [n, 32] descriptors and optionally float32 mvuRight, one entry per keyframe."""
    B = len(kps_list)
    cap = max([len(k) for k in kps_list] + [1])
    kps = np.zeros((B, cap), KP_DTYPE); desc = np.zeros((B, cap, 32), np.uint8); n = np.zeros(B, np.int32)
    ur = np.full((B, cap), -1, np.float32) if uright_list is not None else None
    for f in range(B):
        m = len(kps_list[f]); n[f] = m
        kps[f, :m] = kps_list[f]; desc[f, :m] = desc_list[f]
        if ur is not None:
            ur[f, :m] = uright_list[f]
    ex._check(_map_lib().orb_load_frames(ex.h, _p(kps), _p(desc), _p(ur) if ur is not None else None, _p(n), B, cap, 0))
    ex.cur_batch = B
    return n


def fuse_search(ex, queries, qdesc, nq, th, mode=0, flags=0):
    """The search of ORBmatcher::Fuse for every resident frame. queries: FQ_DTYPE [B, qcap], qdesc: uint8 [B, qcap, 32], nq: int32 [B].
    Returns (best_idx[B, qcap], best_dist[B, qcap])."""
    queries = np.ascontiguousarray(queries, dtype=FQ_DTYPE); qdesc = np.ascontiguousarray(qdesc, dtype=np.uint8)
    nq = np.ascontiguousarray(nq, dtype=np.int32)
    B, qcap = queries.shape
    bi = np.full((B, qcap), -1, np.int32); bd = np.full((B, qcap), 256, np.int32)
    ex._check(_map_lib().orb_fuse_search(ex.h, _p(queries), _p(qdesc), _p(nq), qcap, float(th), int(mode), _p(bi), _p(bd), flags))
    return bi, bd


def search_by_projection_kf(ex, queries, qdesc, nq, locked0, th, orb_dist, check_orientation=True, flags=0):
    """ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) for every resident frame. queries: Q_DTYPE
    [B, qcap], locked0: uint8 [B, kcap] or None. Returns (nmatches[B], match[B, kcap])."""
    queries = np.ascontiguousarray(queries, dtype=Q_DTYPE); qdesc = np.ascontiguousarray(qdesc, dtype=np.uint8)
    nq = np.ascontiguousarray(nq, dtype=np.int32)
    B, qcap = queries.shape
    lk = None
    if locked0 is not None:
        locked0 = np.ascontiguousarray(locked0, dtype=np.uint8)
        assert locked0.shape == (B, ex.kcap)
        lk = _p(locked0)
    nm = np.zeros(B, np.int32); match = np.full((B, ex.kcap), -1, np.int32)
    ex._check(_map_lib().orb_search_by_projection_kf(ex.h, _p(queries), _p(qdesc), _p(nq), qcap, lk, float(th), int(orb_dist),
                                                     int(check_orientation), _p(match), _p(nm), flags))
    return nm, match


IQ_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("angle", "<f4"), ("octave", "<i4")])  # orb_init_query
assert IQ_DTYPE.itemsize == 16


def search_for_initialization(ex, queries, qdesc, nq, window_size=100, nnratio=0.9, check_orientation=True, flags=0):
    """ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) for every resident frame = F2. queries:
    IQ_DTYPE [B, qcap] (vbPrevMatched + angle / octave of F1's keypoints), qdesc uint8 [B, qcap, 32] = F1.mDescriptors.
    Returns (nmatches[B], vnMatches12[B, qcap], vbPrevMatched[B, qcap, 2])."""
    queries = np.ascontiguousarray(queries, dtype=IQ_DTYPE); qdesc = np.ascontiguousarray(qdesc, dtype=np.uint8)
    nq = np.ascontiguousarray(nq, dtype=np.int32)
    B, qcap = queries.shape
    nm = np.zeros(B, np.int32); m12 = np.full((B, qcap), -1, np.int32); prev = np.zeros((B, qcap, 2), np.float32)
    ex._check(_map_lib().orb_search_for_initialization(ex.h, _p(queries), _p(qdesc), _p(nq), qcap, int(window_size), float(nnratio),
                                                       int(check_orientation), _p(m12), _p(prev), _p(nm), flags))
    return nm, m12, prev


class _KfSet(C.Structure):   # orb_kf_set
    _fields_ = [(n, C.c_void_p) for n in ("kps", "desc", "uright", "has_mp", "n", "fv_node", "fv_off", "fv_feat", "fv_n")] + \
               [("count", C.c_int32), ("cap", C.c_int32)]


def _pack_kf_set(keyframes):
    K = len(keyframes)
    cap = max([len(k["kps"]) for k in keyframes] + [1])
    kps = np.zeros((K, cap), KP_DTYPE); desc = np.zeros((K, cap, 32), np.uint8); hm = np.zeros((K, cap), np.uint8)
    have_ur = any(k.get("uright") is not None for k in keyframes)
    ur = np.full((K, cap), -1, np.float32)
    n = np.zeros(K, np.int32); nn = np.zeros(K, np.int32)
    node = np.zeros((K, cap), np.uint32); off = np.zeros((K, cap + 1), np.int32); feat = np.zeros((K, cap), np.uint32)
    for i, k in enumerate(keyframes):
        m = len(k["kps"]); n[i] = m
        kps[i, :m] = k["kps"]; desc[i, :m] = k["desc"]; hm[i, :m] = k["has_mp"]
        if k.get("uright") is not None:
            ur[i, :m] = k["uright"]
        fv = k["fv"]; j = len(fv["fv_node"]); nn[i] = j
        node[i, :j] = fv["fv_node"]; off[i, :j + 1] = fv["fv_off"]; feat[i, :len(fv["fv_feat"])] = fv["fv_feat"]
    keep = (kps, desc, ur, hm, n, node, off, feat, nn)
    S = _KfSet(kps.ctypes.data, desc.ctypes.data, ur.ctypes.data if have_ur else None, hm.ctypes.data, n.ctypes.data, node.ctypes.data,
               off.ctypes.data, feat.ctypes.data, nn.ctypes.data, K, cap)
    return S, cap, keep


def search_for_triangulation(ex, keyframes, pairs, F12, ep, only_stereo=False, coarse=False, check_orientation=True, flags=0):
    """ORBmatcher::SearchForTriangulation for keyframe pairs. keyframes: list of dicts (kps KP_DTYPE, desc, uright or None, has_mp
    uint8, fv = dict fv_node / fv_off / fv_feat); pairs: [(i1, i2)]; F12: float32 [npairs, 9]; ep: float32 [npairs, 2].
    Returns (nmatches[npairs], match12[npairs, cap])."""
    S, cap, _keep = _pack_kf_set(keyframes)
    pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
    k1 = np.ascontiguousarray(pairs[:, 0]); k2 = np.ascontiguousarray(pairs[:, 1])
    P = len(pairs)
    F12 = np.ascontiguousarray(F12, dtype=np.float32).reshape(P, 9); ep = np.ascontiguousarray(ep, dtype=np.float32).reshape(P, 2)
    nm = np.zeros(P, np.int32); m12 = np.full((P, cap), -1, np.int32)
    ex._check(_map_lib().orb_search_for_triangulation(ex.h, C.byref(S), _p(k1), _p(k2), _p(F12), _p(ep), P, int(only_stereo), int(coarse),
                                                      int(check_orientation), _p(m12), _p(nm), flags))
    return nm, m12


def search_for_triangulation_fisheye(ex, keyframes, pairs, rigs, only_stereo=False, coarse=False, check_orientation=True, flags=0):
    """ORBmatcher::SearchForTriangulation between two-camera keyframes (mpCamera2 != NULL). keyframes: dicts as for
    search_for_triangulation with kps = the left keypoints followed by the right ones and nleft; rigs: synth.RIG_DTYPE [npairs, 4]
    (orb_kb8_rig for the combinations ll, lr, rl, rr). Returns (nmatches[npairs], match12[npairs, cap])."""
    S, cap, _keep = _pack_kf_set(keyframes)
    nleft = np.ascontiguousarray([k["nleft"] for k in keyframes], dtype=np.int32)
    pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
    k1 = np.ascontiguousarray(pairs[:, 0]); k2 = np.ascontiguousarray(pairs[:, 1])
    P = len(pairs)
    rigs = np.ascontiguousarray(rigs).reshape(P, 4)
    assert rigs.dtype.itemsize == 120      # sizeof(orb_kb8_rig)
    nm = np.zeros(P, np.int32); m12 = np.full((P, cap), -1, np.int32)
    ex._check(_map_lib().orb_search_for_triangulation_fisheye(ex.h, C.byref(S), _p(nleft), _p(k1), _p(k2), _p(rigs), P, int(only_stereo),
                                                              int(coarse), int(check_orientation), _p(m12), _p(nm), flags))
    return nm, m12


def search_by_bow_kf(ex, keyframes, pairs, nnratio=0.75, check_orientation=True, flags=0):
    """ORBmatcher::SearchByBoW(pKF1, pKF2, vpMatches12) for keyframe pairs (has_mp = map point present and not bad).
    Returns (nmatches[npairs], match12[npairs, cap])."""
    S, cap, _keep = _pack_kf_set(keyframes)
    pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
    k1 = np.ascontiguousarray(pairs[:, 0]); k2 = np.ascontiguousarray(pairs[:, 1])
    P = len(pairs)
    nm = np.zeros(P, np.int32); m12 = np.full((P, cap), -1, np.int32)
    ex._check(_map_lib().orb_search_by_bow_kf(ex.h, C.byref(S), _p(k1), _p(k2), P, float(nnratio), int(check_orientation), _p(m12), _p(nm), flags))
    return nm, m12


def search_by_projection_sim3(ex, queries, qdesc, nq, matched0, th, ratio_hamming=1.0, flags=0):
    """ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) for every resident frame. queries: Q_DTYPE
    [B, qcap], matched0: uint8 [B, kcap] or None. Returns (nmatches[B], match[B, kcap])."""
    queries = np.ascontiguousarray(queries, dtype=Q_DTYPE); qdesc = np.ascontiguousarray(qdesc, dtype=np.uint8)
    nq = np.ascontiguousarray(nq, dtype=np.int32)
    B, qcap = queries.shape
    lk = None
    if matched0 is not None:
        matched0 = np.ascontiguousarray(matched0, dtype=np.uint8)
        assert matched0.shape == (B, ex.kcap)
        lk = _p(matched0)
    nm = np.zeros(B, np.int32); match = np.full((B, ex.kcap), -1, np.int32)
    ex._check(_map_lib().orb_search_by_projection_sim3(ex.h, _p(queries), _p(qdesc), _p(nq), qcap, lk, float(th), float(ratio_hamming),
                                                       _p(match), _p(nm), flags))
    return nm, match


def distinctive_descriptors(ex, desc_lists):
    """MapPoint::ComputeDistinctiveDescriptors for a list of map points (each a uint8 [N_p, 32] array of observed descriptors).
    Returns (best[npoints], median[npoints])."""
    P = len(desc_lists)
    off = np.zeros(P + 1, np.int32)
    for p, d in enumerate(desc_lists):
        off[p + 1] = off[p] + len(d)
    allv = np.zeros((max(int(off[-1]), 1), 32), np.uint8)
    for p, d in enumerate(desc_lists):
        if len(d):
            allv[off[p]:off[p + 1]] = d
    best = np.zeros(P, np.int32); med = np.zeros(P, np.int32)
    ex._check(_map_lib().orb_distinctive_descriptors(ex.h, _p(allv), _p(off), P, _p(best), _p(med), 0))
    return best, med
