#!/usr/bin/env python3
"""Benchmark of the ORB front-end hot path (BASELINE.json metric): stereo frames/s for ORB extraction of
both images + Frame::ComputeStereoMatches on EuRoC-shape 752x480 pairs, 1200 features (configs[1]).

  python bench.py --gpus N --steps K --warmup W            own arm (CUDA, through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU implementation of the path
                                                           (oracle/_ref: its own sources compiled unmodified)

One "step" = one pass of the hot path over one batch of `--batch` synthetic stereo pairs per GPU.
`value` is measured with the inputs already resident in HBM; `e2e` goes through the public C-ABI call
with pinned HOST buffers, host->device and device->host copies inside the timed region. Frames are
independent, so N GPUs each process their own batch (weak scaling, no data-path collective).
One JSON line is printed by rank 0.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from morb_slam_b200 import synth  # noqa: E402

METRIC = "stereo_frames_per_s_orb_extract_plus_stereo_match"
UNIT = "frames/s"
CFG = "euroc"
WORKLOAD_NAME = {"euroc": "EuRoC", "kitti": "KITTI", "tumvi": "TUM-VI"}
# the fisheye rig of the tumvi workload: synth.stereo_pair shifts the scene horizontally, which is what two parallel cameras see
FISHEYE_RIG = synth.kb8_rig("parallel")
CONFIG_INDEX = {"euroc": 1, "kitti": 3, "tumvi": 2}
MATCHER = {"euroc": "ComputeStereoMatches", "kitti": "ComputeStereoMatches", "tumvi": "ComputeStereoFishEyeMatches (knnMatch k=2 + ratio 0.7 + KannalaBrandt8::TriangulateMatches)"}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML every ~2 ms in a thread; falls back
    to one `nvidia-smi` query when NVML is unavailable)."""

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        self._nvml = None

    def prepare(self):
        """NVML initialisation and one query of each kind OUTSIDE the timed region (the first calls of a fresh
        process take tens of milliseconds, longer than a short timed region)."""
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(self._h, pynvml.NVML_CLOCK_SM)
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")
            get_reasons(self._h)
        except Exception:
            self._nvml = None

    def start(self):
        if self._nvml is None:
            self.prepare()
        if self._nvml is None:
            return
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()

    def _run(self):
        nv = self._nvml
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                r = int(get_reasons(self._h))
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join(timeout=1)
        if self.samples:
            return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "samples": len(self.samples),
                    "reasons": sorted(self.reasons), "source": "nvml"}
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=clocks.sm,clocks.max.sm",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
            a, b = [float(x) for x in out.strip().split(",")]
            return {"sm_mhz": a, "sm_max_mhz": b, "samples": 1, "reasons": [], "source": "nvidia-smi after the region"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["clock query unavailable"]}


_PAIR_CACHE = {}


def make_pairs(n, w, h, base_seed):
    key = (n, w, h, base_seed)
    if key in _PAIR_CACHE:
        return _PAIR_CACHE[key]
    L, R = _make_pairs(n, w, h, base_seed)
    _PAIR_CACHE[key] = (L, R)
    return L, R


def _make_pairs(n, w, h, base_seed):
    L = np.empty((n, h, w), np.uint8)
    R = np.empty((n, h, w), np.uint8)
    for i in range(n):
        L[i], R[i] = synth.stereo_pair(base_seed + i, w, h)
    return L, R


# --------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU code (oracle/_ref), all host threads, bounded sample per step
# --------------------------------------------------------------------------------------------------
def cpu_reference_run(n_pairs, threads, w, h, nf, lap, fx, b, seeds_from=2000, distinct=8):
    """Times extract(left) + extract(right) + ComputeStereoMatches for n_pairs pairs on `threads` host threads.
    Returns (pairs_per_second, kind)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle_py as op
    op.build()
    use_ref = op.ref_available()
    Ext = op.RefExtractor if use_ref else op.OracleExtractor
    L, R = make_pairs(distinct, w, h, seeds_from)
    mbf = float(np.float32(fx * b))
    mb = float(np.float32(mbf) / np.float32(fx))

    kb8 = sigma2 = None
    if CFG == "tumvi":
        from oracle import oracle_kb8_py as okb
        kb8 = okb.reference() if os.path.exists(okb.REF_KB8_SO) else okb.oracle()
        sigma2 = ((np.float32(1.2) ** np.arange(8, dtype=np.float32)) ** 2).astype(np.float32)

    def worker(t):
        eL, eR = Ext(nf), Ext(nf)
        cnt = 0
        for i in range(t, n_pairs, threads):
            monoL, kL, dL = eL(L[i % distinct], lap)
            monoR, kR, dR = eR(R[i % distinct], lap)
            _ = (monoL, monoR)
            if CFG == "tumvi":
                # ComputeStereoFishEyeMatches (src/Frame.cc:1222-1250): BFMatcher.knnMatch(k = 2) of the lapping-area descriptors +
                # ratio test (the restatement of cv::BFMatcher, pinned against cv2: the reference calls OpenCV here)
                mLq, mRq = _[0], _[1]
                io, do = op.oracle_knn2(dL[mLq:], dR[mRq:]) if len(dL) > mLq and len(dR) > mRq else (None, None)
                if do is not None:
                    # ... and the triangulation / acceptance loop (:1244-1273; the reference's own lines when oracle/_ref has them)
                    kb8.fisheye_accept(FISHEYE_RIG, kL, mLq, kR, mRq, sigma2, io, do)
            elif use_ref:
                op.ref_stereo(eL, eR, kL, dL, kR, dR, mbf, mb)
            else:
                op.oracle_stereo(eL, eR, kL, dL, kR, dR, mbf, float(np.float32(fx)))
            cnt += 1
        return cnt

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        done = sum(ex.map(worker, range(threads)))
    dt = time.perf_counter() - t0
    return done / dt, ("reference" if use_ref else "port"), dt


def cv2_primitives_ms(w, h, nf):
    """Context only: single-thread time of the real OpenCV kernels the reference calls per image (resize chain,
    per-cell FAST with both thresholds, 8 Gaussian blurs), via the cv2 wheel when it is installed. Excludes the
    quad-tree, orientation and descriptor code of the reference. None when cv2 is absent."""
    try:
        import cv2
    except Exception:
        return None
    cv2.setNumThreads(1)
    img = synth.mono_frame(77, w, h)
    det20 = cv2.FastFeatureDetector_create(threshold=20, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    det7 = cv2.FastFeatureDetector_create(threshold=7, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)

    def once():
        levels = [img]
        for l in range(1, 8):
            inv = np.float32(1.0) / (np.float32(1.2) ** l)
            dw, dh = int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))
            levels.append(cv2.resize(levels[-1], (dw, dh), interpolation=cv2.INTER_LINEAR))
        for lv in levels:
            H, W = lv.shape
            wd, hd = W - 32, H - 32
            nc, nr = int(wd / 35), int(hd / 35)
            wc, hc = int(np.ceil(wd / nc)), int(np.ceil(hd / nr))
            for i in range(nr):
                for j in range(nc):
                    roi = lv[16 + i * hc:min(16 + i * hc + hc + 6, H - 16), 16 + j * wc:min(16 + j * wc + wc + 6, W - 16)]
                    if roi.shape[0] < 7 or roi.shape[1] < 7:
                        continue
                    if not det20.detect(roi, None):
                        det7.detect(roi, None)
            cv2.GaussianBlur(lv, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)

    once()
    t0 = time.perf_counter()
    for _ in range(3):
        once()
    return (time.perf_counter() - t0) / 3 * 1e3


def shim_primitives_ms(w, h):
    """The same primitive calls as cv2_primitives_ms, through the scalar restatement the reference arm is compiled against
    (oracle/shim: shim_resize / shim_fast / shim_gauss7), one thread."""
    from oracle import oracle_py as op
    lib = op.oracle_lib()
    img = synth.mono_frame(77, w, h)
    out = np.zeros(3 * 4096, np.int32)

    def once():
        levels = [img]
        for l in range(1, 8):
            inv = np.float32(1.0) / (np.float32(1.2) ** l)
            dw, dh = int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))
            dst = np.empty((dh, dw), np.uint8)
            s_ = levels[-1]
            lib.shim_resize(op._p(s_), s_.shape[1], s_.shape[0], s_.strides[0], op._p(dst), dw, dh)
            levels.append(dst)
        for lv in levels:
            H, W = lv.shape
            wd, hd = W - 32, H - 32
            nc, nr = int(wd / 35), int(hd / 35)
            wc, hc = int(np.ceil(wd / nc)), int(np.ceil(hd / nr))
            for i in range(nr):
                for j in range(nc):
                    y0, x0 = 16 + i * hc, 16 + j * wc
                    y1, x1 = min(y0 + hc + 6, H - 16), min(x0 + wc + 6, W - 16)
                    if y1 - y0 < 7 or x1 - x0 < 7:
                        continue
                    roi = lv[y0:y1, x0:x1]
                    if lib.shim_fast(op._p(roi), x1 - x0, y1 - y0, lv.strides[0], 20, op._p(out), 4096) == 0:
                        lib.shim_fast(op._p(roi), x1 - x0, y1 - y0, lv.strides[0], 7, op._p(out), 4096)
            dst = np.empty_like(lv)
            lib.shim_gauss7(op._p(lv), W, H, lv.strides[0], op._p(dst))

    once()
    t0 = time.perf_counter()
    for _ in range(3):
        once()
    return (time.perf_counter() - t0) / 3 * 1e3


def real_opencv_estimate(value, w, h, nf, lap):
    """What the reference arm would reach with the real OpenCV kernels in place of the scalar restatement: per image, the time of
    the reference's code on one thread minus the restated primitives plus the same primitives through cv2 (SIMD). An ESTIMATE
    (cv2 is only available as the Python wheel here: the reference cannot be linked against it)."""
    try:
        from oracle import oracle_py as op
        Ext = op.RefExtractor if op.ref_available() else op.OracleExtractor
        img = synth.mono_frame(77, w, h)
        e = Ext(nf)
        e(img, lap)
        t0 = time.perf_counter()
        for _ in range(3):
            e(img, lap)
        t_ref = (time.perf_counter() - t0) / 3 * 1e3
        t_shim = shim_primitives_ms(w, h)
        t_cv2 = cv2_primitives_ms(w, h, nf)
        if t_cv2 is None:
            return None
        t_est = max(t_ref - t_shim, 0.0) + t_cv2
        return {"value": value * t_ref / t_est, "unit": UNIT,
                "ms_per_image_reference_on_scalar_primitives": t_ref, "ms_per_image_scalar_primitives": t_shim,
                "ms_per_image_cv2_primitives": t_cv2, "ms_per_image_estimated": t_est,
                "how": "value x t_ref / (t_ref - t_scalar_primitives + t_cv2_primitives), single-thread times per image"}
    except Exception as e:  # context only
        return {"error": repr(e)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w, h, nf, lap, fx, b = synth.CONFIGS[CFG]
    cores = os.cpu_count() or 1
    per_step = 16 * cores          # bounded sample: about one second of wall time per step
    for _ in range(max(args.warmup, 0) and 1):
        cpu_reference_run(min(per_step, cores), cores, w, h, nf, lap, fx, b)
    t_total, n_total, kind = 0.0, 0, "reference"
    for _ in range(args.steps):
        fps, kind, dt = cpu_reference_run(per_step, cores, w, h, nf, lap, fx, b)
        t_total += dt
        n_total += per_step
    value = n_total / t_total
    sample = "%d stereo pairs per step x %d steps, %d host threads, extract(L)+extract(R)+%s" % (
        per_step, args.steps, cores, MATCHER[CFG].split(" ")[0])
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * t_total / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "%s-shape stereo %dx%d, %d features/image, ORB extract x2 + %s (CPU)" % (WORKLOAD_NAME[CFG], w, h, nf, MATCHER[CFG]),
                       "pairs_per_step": per_step},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                             "cv2_primitives_ms_per_image": cv2_primitives_ms(w, h, nf),
                             "real_opencv_estimate": real_opencv_estimate(value, w, h, nf, lap)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))



def bind_to_gpu_numa_node(dev):
    """Pin this rank's host threads to the cores of the NUMA node its GPU hangs off (sysfs), so the staging copies and the DMA
    descriptors stay on the local memory controller. Returns a short description; a box that exposes one node is left alone."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pci = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(dev))
        bus = pci.busId
        if isinstance(bus, bytes):
            bus = bus.decode()
        if not isinstance(bus, str):     # some nvidia-ml-py versions hand back a number: rebuild the sysfs name from the fields
            bus = "%04x:%02x:%02x.0" % (pci.domain, pci.bus, pci.device)
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if len(nodes) < 2 or node < 0:
            return "one NUMA node exposed (%d found, GPU reports node %d): nothing to bind" % (len(nodes), node)
        cpus = []
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, set(cpus) & os.sched_getaffinity(0) or os.sched_getaffinity(0))
        return "bound to NUMA node %d (%d cores)" % (node, len(cpus))
    except Exception as e:
        return "not bound: %r" % (e,)


def pcie_concurrent_probe(capi, torch, dist, P0, hostL, hostR, dL, dR, dev, reps=4):
    """What the host can move while EVERY rank copies at once: each rank uploads its step's input bytes (both cameras) on one
    stream while it downloads the step's result bytes on another, barrier before, wall clock around, max over ranks.
    ceiling_frames_per_s = the frame rate at which the step's copies alone would saturate that (all ranks together)."""
    L_ = capi.lib()
    B = hostL.shape[0]
    h2d = hostL.nbytes + hostR.nbytes
    d2h = int(P0.d2h_bytes())
    dst = torch.empty(d2h, dtype=torch.uint8, device="cuda:%d" % dev)
    hdst = capi.pinned_empty((d2h,), np.uint8)

    def once():
        L_.orb_memcpy_h2d_async(P0.exL.h, capi._p(dL.data_ptr()), capi._p(hostL), hostL.nbytes)
        L_.orb_memcpy_h2d_async(P0.exL.h, capi._p(dR.data_ptr()), capi._p(hostR), hostR.nbytes)
        L_.orb_memcpy_d2h_async(P0.exR.h, capi._p(hdst), capi._p(dst.data_ptr()), d2h)
    once()
    P0.exL.sync(); P0.exR.sync()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    P0.exL.sync(); P0.exR.sync()
    dt = (time.perf_counter() - t0) / reps
    world = 1
    if dist is not None:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda:%d" % dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t[0])
        world = dist.get_world_size()
    return {"ms_per_step_copies": dt * 1e3, "h2d_gbs_per_gpu": h2d / dt / 1e9, "d2h_gbs_per_gpu": d2h / dt / 1e9,
            "ceiling_frames_per_s": world * B / dt,
            "what": "all ranks copy one step's inputs (H2D) and results (D2H) at the same time, nothing else running"}

# --------------------------------------------------------------------------------------------------
# own arm
# --------------------------------------------------------------------------------------------------
class Pair:
    """Left/right extractor handles of one in-flight batch plus its pinned host buffers."""

    def __init__(self, capi, B, w, h, nf, device):
        self.exL = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B, device=device)
        self.exR = capi.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B, device=device)
        k = self.exL.kcap
        pe = capi.pinned_empty
        self.outL = (pe((B,), np.int32), pe((B,), np.int32), pe((B, k), capi.KP_DTYPE), pe((B, k, 32), np.uint8))
        self.outR = (pe((B,), np.int32), pe((B,), np.int32), pe((B, k), capi.KP_DTYPE), pe((B, k, 32), np.uint8))
        self.st = (pe((B, k), np.float32), pe((B, k), np.float32))
        # fisheye: mvLeftToRightMatch, mvRightToLeftMatch, mvDepth, mvStereo3Dpoints, accept / reject code
        self.fe = (pe((B, k), np.int32), pe((B, k), np.int32), pe((B, k), np.float32), pe((B, k, 3), np.float32), pe((B, k), np.int8))

    def d2h_bytes(self):
        return sum(a.nbytes for a in self.outL) + sum(a.nbytes for a in self.outR) + sum(a.nbytes for a in (self.fe if CFG == "tumvi" else self.st))

    def launches(self):
        return self.exL.launch_count() + self.exR.launch_count()


def quick_workload(cfg, capi, torch, dist, dev, rank, world, B, steps, host_mem):
    """Short run of another BASELINE.json configuration at this GPU count, same procedure as the main line (device-resident value
    and e2e through host buffers, two batches in flight, barrier + max over ranks): configs[3] KITTI and configs[2] TUM-VI appear in
    every default line, so the driver's 1 / 2 / 4 / 8-GPU runs carry them too."""
    global CFG
    saved = CFG
    CFG = cfg
    pairs = []
    try:
        w, h, nf, lap, fx, b = synth.CONFIGS[cfg]
        mbf, maxD = float(np.float32(fx * b)), float(np.float32(fx))
        distinct = min(B, 8)
        Ls, Rs = make_pairs(distinct, w, h, 2000 + 1000 * rank)
        hostL = capi.pinned_empty((B, h, w), np.uint8, write_combined=host_mem == "wc")
        hostR = capi.pinned_empty((B, h, w), np.uint8, write_combined=host_mem == "wc")
        for i in range(B):
            hostL[i] = Ls[i % distinct]
            hostR[i] = Rs[i % distinct]
        pairs = [Pair(capi, B, w, h, nf, dev) for _ in range(2)]
        dL = torch.from_numpy(np.ascontiguousarray(Ls[np.arange(B) % distinct])).to("cuda:%d" % dev)
        dR = torch.from_numpy(np.ascontiguousarray(Rs[np.arange(B) % distinct])).to("cuda:%d" % dev)
        NO, AS = capi.ORB_NO_OUTPUT, capi.ORB_ASYNC
        fisheye = cfg == "tumvi"
        rig_c = capi.kb8_rig(FISHEYE_RIG)

        def step(p, e2e):
            if e2e:
                p.exL.extract_batch(hostL, lap, out=p.outL, flags=AS)
                p.exR.extract_batch(hostR, lap, out=p.outR, flags=AS)
            else:
                p.exL.extract_batch((dL.data_ptr(), B, h, w), lap, out=p.outL, flags=NO | AS)
                p.exR.extract_batch((dR.data_ptr(), B, h, w), lap, out=p.outR, flags=NO | AS)
            if fisheye:
                capi.compute_stereo_fisheye_matches_batch(p.exL, p.exR, flags=AS, want=False)
                if e2e:
                    p.exL._check(p.exL.L.orb_stereo_fisheye_triangulate_batch(p.exL.h, p.exR.h, ctypes.byref(rig_c), *[capi._p(a) for a in p.fe],
                                                                              p.exL.kcap, AS))
                else:
                    capi.compute_stereo_fisheye_triangulation_batch(p.exL, p.exR, rig_c, flags=AS, want=False)
            elif e2e:
                capi.compute_stereo_matches_batch(p.exL, p.exR, mbf, maxD, out=p.st, flags=AS)
            else:
                capi.compute_stereo_matches_batch(p.exL, p.exR, mbf, maxD, out=(None, None), flags=NO | AS)

        def barrier():
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()

        ms = []
        for e2e in (False, True):
            for i in range(3):
                step(pairs[i % 2], e2e)
            for p in pairs:
                p.exR.sync(); p.exL.sync()
            barrier()
            pairs[0].exL.timer_start()
            for i in range(steps):
                p = pairs[i % 2]
                if e2e and i >= 2:
                    p.exR.sync(); p.exL.sync()
                step(p, e2e)
            for p in (pairs[1], pairs[0]):
                p.exR.sync(); p.exL.sync()
            ms.append(pairs[0].exL.timer_stop())
            barrier()
        t = torch.tensor(ms, dtype=torch.float64, device="cuda:%d" % dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        frames = world * B * steps
        K = float(np.mean(pairs[0].outL[0][:B]))
        return {"workload": "%s-shape stereo %dx%d, %d features/image, ORB extract x2 + %s (BASELINE.json configs[%d])" % (
                    WORKLOAD_NAME[cfg], w, h, nf, MATCHER[cfg].split(" ")[0], CONFIG_INDEX[cfg]),
                "value": frames / (float(t[0]) * 1e-3), "unit": UNIT, "n_gpus": world, "stereo_pairs_per_step_per_gpu": B, "steps": steps,
                "ms_per_step": float(t[0]) / steps, "keypoints_per_image": K,
                "e2e": {"value": frames / (float(t[1]) * 1e-3), "unit": UNIT, "ms_per_step": float(t[1]) / steps,
                        "h2d_bytes_per_step": int(hostL.nbytes + hostR.nbytes), "d2h_bytes_per_step": int(pairs[0].d2h_bytes())}}
    except Exception as e:  # context line: never takes the main line down
        return {"error": repr(e)}
    finally:
        CFG = saved
        for p in pairs:
            p.exL.close(); p.exR.close()


def mapping_lines(args, capi, synth, torch, dist, P0, B, distinct, w, h, mbf, dev, world):
    """Fuse search, SearchByProjection(Frame, KeyFrame), SearchForTriangulation, ComputeDistinctiveDescriptors on the resident left
    frames; every line = max over ranks of the mean wall time of `reps` calls through host buffers."""
    import numpy as np
    from oracle import oracle_map_py as omap
    ex = P0.exL
    nLh, _, kLh, dLh = P0.outL
    uRh = P0.st[0]
    nd = min(B, distinct)
    t = ex.tables()
    gp = capi.grid_params(w, h)
    capi.assign_features_to_grid(ex, gp)
    reps = 5

    def timed(fn):
        for _ in range(2):
            fn()
        ex.sync()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        ex.sync()
        ms = (time.perf_counter() - t0) * 1e3 / reps
        tm = torch.tensor([ms], dtype=torch.float64, device="cuda:%d" % dev)
        if dist is not None:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        return float(tm[0])
    out = {}
    ref = omap.reference() if (omap.have_reference() and world == 1 and not args.no_cpu_baseline) else None
    # Fuse: ~0.8 candidate map points per keypoint
    sets = [synth.synth_fuse_points(900 + i, kLh[i, :int(nLh[i])], dLh[i, :int(nLh[i])], w, h) for i in range(nd)]
    qs = [omap.fuse_queries(s_[0], mbf) for s_ in sets]
    qcap = max(len(q) for q in qs)
    FQ = np.zeros((B, qcap), capi.FQ_DTYPE); FQD = np.zeros((B, qcap, 32), np.uint8); fnq = np.zeros(B, np.int32)
    for i in range(B):
        q = qs[i % nd]
        FQ[i, :len(q)], FQD[i, :len(q)], fnq[i] = q, sets[i % nd][1], len(q)
    res = {}
    ms = timed(lambda: res.__setitem__("f", capi.fuse_search(ex, FQ, FQD, fnq, 3.0, 0)))
    out["fuse"] = {"what": "the search of ORBmatcher::Fuse(pKF, vpMapPoints, th=3) per keyframe, host buffers", "keyframes": B * world,
                   "map_points_per_keyframe": float(fnq.mean()), "ms_per_batch": ms, "keyframes_per_s": B * world / (ms * 1e-3),
                   "found_frame0": int((res["f"][1][0] <= 50).sum())}
    if ref is not None:
        t0 = time.perf_counter()
        for i in range(nd):
            n = int(nLh[i]); pts, pdesc, kn, kb = sets[i]
            ref.fuse(kLh[i, :n], dLh[i, :n], uRh[i, :n], gp, t["scale"], t["sigma2"], mbf, kn, kb, pts, pdesc, 3.0)
        out["fuse"]["cpu_baseline"] = {"keyframes_per_s": nd / (time.perf_counter() - t0), "cores": 1, "kind": "reference",
                                       "sample": "%d keyframes, grid build + Fuse on 1 host thread" % nd}
    # SearchByProjection(Frame, KeyFrame): the keyframe's map points = the frame's own keypoints seen with noise
    rng = np.random.default_rng(3)
    kq = []
    for i in range(nd):
        n = int(nLh[i])
        q, qd = synth.synth_queries(950 + i, kLh[i, :n], dLh[i, :n], None, None, w, h, jitter=2.0)
        q["flags"] &= 1
        kq.append((q, synth.flip_bits(rng, qd, 40)))
    kcap_q = max(len(q) for q, _ in kq)
    KQ = np.zeros((B, kcap_q), capi.Q_DTYPE); KQD = np.zeros((B, kcap_q, 32), np.uint8); knq = np.zeros(B, np.int32)
    for i in range(B):
        q, qd = kq[i % nd]
        KQ[i, :len(q)], KQD[i, :len(q)], knq[i] = q, qd, len(q)
    ms = timed(lambda: res.__setitem__("k", capi.search_by_projection_kf(ex, KQ, KQD, knq, None, 10.0, 100, True)))
    out["search_by_projection_keyframe"] = {"what": "ORBmatcher::SearchByProjection(CurrentFrame, pKF, found, th=10, ORBdist=100) per frame, host buffers",
                                            "frames": B * world, "ms_per_batch": ms, "frames_per_s": B * world / (ms * 1e-3),
                                            "matches_frame0": int(res["k"][0][0])}
    if ref is not None:
        t0 = time.perf_counter()
        for i in range(nd):
            n = int(nLh[i])
            ref.search_by_projection_kf(kLh[i, :n], dLh[i, :n], None, t["scale"], gp, kq[i][0], kq[i][1], 10.0, 100, True)
        out["search_by_projection_keyframe"]["cpu_baseline"] = {"frames_per_s": nd / (time.perf_counter() - t0), "cores": 1, "kind": "reference",
                                                                "sample": "%d frames, grid build + SearchByProjection(KeyFrame) on 1 host thread" % nd}
    # SearchForTriangulation: every distinct frame against a moved copy of itself
    kfs, pairs, Fs = [], [], []
    for i in range(nd):
        n = int(nLh[i])
        k1, k2 = synth.synth_triangulation_pair(980 + i, kLh[i, :n], dLh[i, :n], uRh[i, :n], w, h)
        kfs += [k1, k2]
        Fs.append(synth.synth_fundamental(980 + i))
    pairs = [(2 * (p % nd), 2 * (p % nd) + 1) for p in range(B)]
    F12 = np.array([Fs[p % nd] for p in range(B)], np.float32); eps = np.tile(np.array([[5000.0, 240.0]], np.float32), (B, 1))
    ms = timed(lambda: res.__setitem__("t", capi.search_for_triangulation(ex, kfs, pairs, F12, eps, False, False, True)))
    out["search_for_triangulation"] = {"what": "ORBmatcher::SearchForTriangulation per keyframe pair (Pinhole epipolar test), host buffers incl. the packing of the keyframe set",
                                       "pairs": B * world, "ms_per_batch": ms, "pairs_per_s": B * world / (ms * 1e-3),
                                       "matches_pair0": int(res["t"][0][0])}
    if ref is not None:
        t0 = time.perf_counter()
        for i in range(nd):
            ref.search_for_triangulation(kfs[2 * i], kfs[2 * i + 1], gp, t["scale"], t["sigma2"], Fs[i], (5000.0, 240.0), False, False, True)
        out["search_for_triangulation"]["cpu_baseline"] = {"pairs_per_s": nd / (time.perf_counter() - t0), "cores": 1, "kind": "reference",
                                                           "sample": "%d pairs on 1 host thread" % nd}
    # ComputeDistinctiveDescriptors: 20 000 map points with 1 .. 40 observations
    obs = synth.synth_observations(5, 2000)
    obs = [obs[i % len(obs)] for i in range(20000)]
    ms = timed(lambda: res.__setitem__("d", capi.distinctive_descriptors(ex, obs)))
    out["distinctive_descriptors"] = {"what": "MapPoint::ComputeDistinctiveDescriptors for 20 000 map points (1 - 100 observations each), host buffers incl. packing",
                                      "map_points": 20000 * world, "ms_per_batch": ms, "map_points_per_s": 20000 * world / (ms * 1e-3)}
    if ref is not None:
        t0 = time.perf_counter()
        for d in obs[:2000]:
            if len(d) <= 100:
                ref.distinctive(d)
        out["distinctive_descriptors"]["cpu_baseline"] = {"map_points_per_s": 2000 / (time.perf_counter() - t0), "cores": 1, "kind": "reference",
                                                          "sample": "2000 map points on 1 host thread"}
    return out


def run_own_arm(args):
    import torch
    from morb_slam_b200 import capi
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        # stdout carries the one JSON line only: NCCL prints its version banner there while the communicator comes up, so file
        # descriptor 1 points at stderr until the first collective has run
        sys.stdout.flush()
        saved_out = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_out, 1)
            os.close(saved_out)
    dev = local_rank
    torch.cuda.set_device(dev)
    w, h, nf, lap, fx, b = synth.CONFIGS[CFG]
    B = args.batch
    mbf, maxD = float(np.float32(fx * b)), float(np.float32(fx))

    # synthetic input: `distinct` different pairs tiled to the batch (every frame is still processed in full)
    distinct = min(B, args.distinct)
    Ls, Rs = make_pairs(distinct, w, h, 2000 + 1000 * rank)
    bind = bind_to_gpu_numa_node(dev) if args.numa_bind else None
    wc = args.host_mem == "wc"
    hostL = capi.pinned_empty((B, h, w), np.uint8, write_combined=wc)
    hostR = capi.pinned_empty((B, h, w), np.uint8, write_combined=wc)
    for i in range(B):
        hostL[i] = Ls[i % distinct]
        hostR[i] = Rs[i % distinct]
    pairs = [Pair(capi, B, w, h, nf, dev) for _ in range(2)]
    P0 = pairs[0]
    dL = torch.from_numpy(hostL).to("cuda:%d" % dev)
    dR = torch.from_numpy(hostR).to("cuda:%d" % dev)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    NO, AS = capi.ORB_NO_OUTPUT, capi.ORB_ASYNC

    fisheye = CFG == "tumvi"
    rig_c = capi.kb8_rig(FISHEYE_RIG)

    def step_resident(p):
        p.exL.extract_batch((dL.data_ptr(), B, h, w), lap, out=p.outL, flags=NO | AS)
        p.exR.extract_batch((dR.data_ptr(), B, h, w), lap, out=p.outR, flags=NO | AS)
        if fisheye:
            capi.compute_stereo_fisheye_matches_batch(p.exL, p.exR, flags=AS, want=False)
            capi.compute_stereo_fisheye_triangulation_batch(p.exL, p.exR, rig_c, flags=AS, want=False)
        else:
            capi.compute_stereo_matches_batch(p.exL, p.exR, mbf, maxD, out=(None, None), flags=NO | AS)

    def step_e2e(p):
        p.exL.extract_batch(hostL, lap, out=p.outL, flags=AS)
        p.exR.extract_batch(hostR, lap, out=p.outR, flags=AS)
        if fisheye:   # the kNN lists stay on the device (the reference's local `matches`); the Frame's members come back
            capi.compute_stereo_fisheye_matches_batch(p.exL, p.exR, flags=AS, want=False)
            p.exL._check(p.exL.L.orb_stereo_fisheye_triangulate_batch(p.exL.h, p.exR.h, ctypes.byref(rig_c), *[capi._p(a) for a in p.fe],
                                                                      p.exL.kcap, AS))
        else:
            capi.compute_stereo_matches_batch(p.exL, p.exR, mbf, maxD, out=p.st, flags=AS)

    def finish(p):
        p.exR.sync()
        p.exL.sync()

    # ---- device-resident throughput (`value`)
    for i in range(max(args.warmup, 3)):
        step_resident(pairs[i % 2])
    for p in pairs:
        finish(p)
    sampler = ClockSampler(dev)
    sampler.prepare()
    barrier()
    launches0 = sum(p.launches() for p in pairs)
    sampler.start()
    P0.exL.timer_start()
    for i in range(args.steps):
        step_resident(pairs[i % 2])      # two batches in flight, like a production loop
    finish(pairs[1])
    finish(pairs[0])
    ms_resident = P0.exL.timer_stop()    # recorded on pair 0's stream after everything has finished
    clocks = sampler.stop()
    launches = sum(p.launches() for p in pairs) - launches0
    barrier()

    # ---- end to end through the public API with host buffers, two batches in flight
    for i in range(max(args.warmup, 3)):
        step_e2e(pairs[i % 2])
    for p in pairs:
        finish(p)
    barrier()
    P0.exL.timer_start()
    for i in range(args.steps):
        p = pairs[i % 2]
        if i >= 2:
            finish(p)          # the previous batch of this pair has been consumed
        step_e2e(p)
    finish(pairs[1])
    finish(pairs[0])
    ms_e2e = P0.exL.timer_stop() if args.steps % 2 == 1 or args.steps < 2 else None
    if ms_e2e is None:
        # the last batch ran on pair 1: close the interval on pair 0's stream after everything finished
        ms_e2e = P0.exL.timer_stop()
    n_matches = int((pairs[0].st[0][0, :pairs[0].outL[0][0]] >= 0).sum())
    barrier()

    # ---- PCIe context: pinned host <-> device copy bandwidth on this box (linear 128 MiB copies)
    pcie = None
    try:
        conc = pcie_concurrent_probe(capi, torch, dist, P0, hostL, hostR, dL, dR, dev)
    except Exception as e:  # context only
        conc = {"error": repr(e)}
    try:
        nb = 128 << 20
        hb = capi.pinned_empty((nb,), np.uint8)
        db_ = torch.empty(nb, dtype=torch.uint8, device="cuda:%d" % dev)
        L_ = capi.lib()
        for fn in ("h2d", "d2h"):
            for rep in range(2):
                P0.exL.timer_start()
                if fn == "h2d":
                    L_.orb_memcpy_h2d(P0.exL.h, capi._p(db_.data_ptr()), capi._p(hb), nb)
                else:
                    L_.orb_memcpy_d2h(P0.exL.h, capi._p(hb), capi._p(db_.data_ptr()), nb)
                ms = P0.exL.timer_stop()
            pcie = dict(pcie or {}, **{fn + "_gbs": nb / (ms * 1e-3) / 1e9})
        del db_
    except Exception as e:  # context only
        pcie = {"error": str(e)}
    pcie = dict(pcie or {}, concurrent=conc, host_mem=args.host_mem, numa_bind=bind)

    # ---- per-stage device time (CUDA events between the stages, on the launching stream)
    P0.exL.set_stage_timing(True)
    reps = 5
    stage_reps = []
    for _ in range(reps):
        P0.exL.extract_batch((dL.data_ptr(), B, h, w), lap, out=P0.outL, flags=NO)
        P0.exR.extract_batch((dR.data_ptr(), B, h, w), lap, out=P0.outR, flags=NO)
        capi.compute_stereo_matches_batch(P0.exL, P0.exR, mbf, maxD, out=(None, None), flags=NO)
        stage_reps.append(np.array(P0.exL.stage_times(), dtype=np.float64))
    stage = np.median(np.stack(stage_reps), axis=0)   # median: the first repetition of a fresh process can hiccup
    P0.exL.set_stage_timing(False)
    cand = P0.exL.level_counts(B).sum(axis=1).mean()
    K = float(np.mean(P0.outL[0][:B]))

    # ---- kNN line (BASELINE.json configs[4]): 1200 queries vs a database row-sharded over the ranks
    #      (1.25 M rows per GPU = 10 M rows on 8 GPUs); local top-2 -> one NCCL all-gather -> merge kernel
    knn = None
    if not args.no_knn:
        from morb_slam_b200 import sharding
        nq, ndb = 1200, args.knn_rows
        q = torch.from_numpy(synth.random_descriptors(10, nq)).to("cuda:%d" % dev)   # same queries on every rank
        g = torch.Generator(device="cuda:%d" % dev)
        g.manual_seed(1234 + rank)
        dbt = torch.randint(0, 256, (ndb, 32), dtype=torch.uint8, device="cuda:%d" % dev, generator=g)
        base = rank * ndb
        # the exchange is fused into the kernels over peer memory (sharding.ShardedKnn: NVLink peer stores + flag wait); the plain
        # route (one NCCL all-gather + merge kernel) is timed next to it and must give the same lists
        sk = sharding.ShardedKnn(P0.exL, nq)
        kout = torch.empty((2, nq, 2), dtype=torch.int32, device="cuda:%d" % dev)
        for _ in range(2):
            oi, od = sk.search(q, dbt, base, out=kout)
            ni, nd = sharding.sharded_knn2(P0.exL, q, dbt, base)
        torch.cuda.synchronize()
        routes_equal = bool(torch.equal(oi, ni) and torch.equal(od, nd))
        barrier()
        kreps = 10
        # every search = scan + merge-and-push + wait-and-merge enqueued on the handle's stream, one host synchronisation at the end;
        # wall-clock between device synchronisations, max over ranks
        t0 = time.perf_counter()
        for _ in range(kreps):
            sk.search(q, dbt, base, out=kout, flags=AS)
        sk.check()
        torch.cuda.synchronize()
        kms = (time.perf_counter() - t0) * 1e3 / kreps
        barrier()
        t0 = time.perf_counter()
        for _ in range(kreps):
            ni, nd = sharding.sharded_knn2(P0.exL, q, dbt, base)
        torch.cuda.synchronize()
        kms_nccl = (time.perf_counter() - t0) * 1e3 / kreps
        # scan kernel alone (device-timed on the handle's stream)
        fl = capi.ORB_SRC_DEVICE | capi.ORB_DST_DEVICE | AS
        ti = torch.empty((nq, 2), dtype=torch.int32, device="cuda:%d" % dev)
        td = torch.empty((nq, 2), dtype=torch.int32, device="cuda:%d" % dev)
        P0.exL.timer_start()
        for _ in range(kreps):
            capi.hamming_knn2(P0.exL, q.data_ptr(), dbt.data_ptr(), base, fl, ndb=ndb, nq=nq, out=(ti.data_ptr(), td.data_ptr()))
        scan_ms = P0.exL.timer_stop() / kreps
        tk = torch.tensor([kms, scan_ms, kms_nccl, 0.0 if routes_equal else 1.0], dtype=torch.float64, device="cuda:%d" % dev)
        if dist is not None:
            dist.all_reduce(tk, op=dist.ReduceOp.MAX)
        kms, scan_ms, kms_nccl, routes_equal = float(tk[0]), float(tk[1]), float(tk[2]), float(tk[3]) == 0.0
        sk.close()
        knn = {"queries": nq, "db_rows_total": ndb * world, "db_rows_per_gpu": ndb, "ms_sharded_search": kms,
               "exchange": "peer-memory stores + epoch flags inside the merge kernels (no collective call)",
               "ms_sharded_search_nccl_allgather_route": kms_nccl, "routes_equal": routes_equal,
               "ms_scan_kernel": scan_ms, "pairs_per_s": nq * ndb * world / (kms * 1e-3),
               "pairs_per_s_scan_kernel_per_gpu": nq * ndb / (scan_ms * 1e-3),
               "queries_per_s_at_db": nq / (kms * 1e-3),
               # the plain formulation's bound, 8 POPC per pair at 16 lanes/clk/SM; the kernel trades two of them for
               # four LOP3 on the ALU pipe (carry-save adders) and may exceed it
               "popc_roofline_pairs_per_s_per_gpu": 148 * 16 / 8 * 1.965e9,
               "int_pipe_bound_pairs_per_s_per_gpu": 148 * min(16 / 6.0, 64 / 28.0) * 1.965e9,   # 6 POPC (XU) | ~28 ALU ops per pair
               "checksum": int(oi.sum().item())}

    # ---- windowed matcher line: Frame::AssignFeaturesToGrid + ORBmatcher::SearchByProjection on the
    #      device-resident left frames of the last batch; one query per right-image keypoint (projected map points)
    match = None
    if not args.no_match:
        nRh, _, kRh, dRh = P0.outR      # host copies of the right extraction (filled by the e2e phase)
        qcap = P0.exR.kcap
        Q = np.zeros((B, qcap), capi.Q_DTYPE)
        QD = np.zeros((B, qcap, 32), np.uint8)
        nq = np.zeros(B, np.int32)
        for i in range(min(B, distinct)):
            n = int(nRh[i])
            q, qd = synth.synth_queries(500 + i, kRh[i, :n], dRh[i, :n], None, None, w, h)
            Q[i, :n], QD[i, :n], nq[i] = q, qd, n
        for i in range(distinct, B):
            Q[i], QD[i], nq[i] = Q[i % distinct], QD[i % distinct], nq[i % distinct]
        dQ = torch.from_numpy(Q.view(np.uint8).reshape(B, -1)).to("cuda:%d" % dev)
        dQD = torch.from_numpy(QD).to("cuda:%d" % dev)
        dnq = torch.from_numpy(nq).to("cuda:%d" % dev)
        dtz = torch.zeros(B, dtype=torch.float32, device="cuda:%d" % dev)
        dm = torch.empty((B, P0.exL.kcap), dtype=torch.int32, device="cuda:%d" % dev)
        dnm = torch.empty(B, dtype=torch.int32, device="cuda:%d" % dev)
        gp = capi.grid_params(w, h)
        fl = capi.ORB_SRC_DEVICE | capi.ORB_DST_DEVICE | AS
        src = (dQ.data_ptr(), dQD.data_ptr(), dnq.data_ptr(), dtz.data_ptr(), B, qcap)

        def match_step():
            capi.assign_features_to_grid(P0.exL, gp, AS)
            capi.search_by_projection(P0.exL, src, None, None, 7.0, False, None, float(np.float32(b)), mbf, True,
                                      out=(dnm.data_ptr(), dm.data_ptr()), flags=fl)
        for _ in range(3):
            match_step()
        P0.exL.sync()
        mreps = 10
        P0.exL.timer_start()
        for _ in range(mreps):
            match_step()
        mms = P0.exL.timer_stop() / mreps
        tm = torch.tensor([mms], dtype=torch.float64, device="cuda:%d" % dev)
        if dist is not None:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        mms = float(tm[0])
        match = {"what": "AssignFeaturesToGrid + SearchByProjection(th=7, stereo) per frame, device-resident", "frames": B * world,
                 "queries_per_frame": float(nq.mean()), "ms_per_batch": mms, "frames_per_s": B * world / (mms * 1e-3),
                 "queries_per_s": float(nq.sum()) * world / (mms * 1e-3), "matches_frame0": int(dnm[0].item())}
        if world == 1 and not args.no_cpu_baseline:
            # the reference's own SearchByProjection (oracle/_ref, else the restatement) on one host core, same inputs
            try:
                from oracle import oracle_match_py as om
                impl, kind = (om.reference(), "reference") if om.have_reference() else (om.oracle(), "port")
                nLh, _, kLh, dLh = P0.outL
                uRh = P0.st[0]
                scale_t = P0.exL.tables()["scale"]
                t0 = time.perf_counter()
                nfr = min(B, distinct)
                for i in range(nfr):
                    nC, n = int(nLh[i]), int(nq[i])
                    impl.search_by_projection(kLh[i, :nC], dLh[i, :nC], uRh[i, :nC], scale_t, gp, float(np.float32(b)), mbf, Q[i, :n], QD[i, :n],
                                              7.0, False, 0.0, True)
                dtm = time.perf_counter() - t0
                match["cpu_baseline"] = {"frames_per_s": nfr / dtm, "cores": 1, "kind": kind,
                                         "sample": "%d frames, grid build + SearchByProjection, 1 host thread" % nfr}
            except Exception as e:  # context only
                match["cpu_baseline"] = {"error": str(e)}

        # local-map search (ORBmatcher::SearchByProjection(F, vpMapPoints, th = 3), Tracking::SearchLocalPoints): 1.5 map points
        # per keypoint of the left frames (their own keypoints as tracked map points + wrong associations), 30 % locked before
        nLh, _, kLh, dLh = P0.outL
        uRh = P0.st[0]
        tqs = [synth.synth_track_queries(700 + i, kLh[i, :int(nLh[i])], dLh[i, :int(nLh[i])], uRh[i, :int(nLh[i])], w, h, mbf=mbf)
               for i in range(min(B, distinct))]
        tqcap = max(len(q) for q, _ in tqs)
        TQ = np.zeros((B, tqcap), capi.TQ_DTYPE)
        TQD = np.zeros((B, tqcap, 32), np.uint8)
        tnq = np.zeros(B, np.int32)
        for i in range(B):
            q, qd = tqs[i % len(tqs)]
            TQ[i, :len(q)], TQD[i, :len(q)], tnq[i] = q, qd, len(q)
        lk0 = (np.random.default_rng(9).random((B, P0.exL.kcap)) < 0.3).astype(np.uint8)
        for i in range(distinct, B):
            lk0[i] = lk0[i % distinct]
        dTQ = torch.from_numpy(TQ.view(np.uint8).reshape(B, -1)).to("cuda:%d" % dev)
        dTQD = torch.from_numpy(TQD).to("cuda:%d" % dev)
        dtnq = torch.from_numpy(tnq).to("cuda:%d" % dev)
        dlk = torch.from_numpy(lk0).to("cuda:%d" % dev)
        tsrc = (dTQ.data_ptr(), dTQD.data_ptr(), dtnq.data_ptr(), dlk.data_ptr(), B, tqcap)

        def local_step():
            capi.search_local_points(P0.exL, tsrc, None, None, None, 3.0, 0.8, out=(dnm.data_ptr(), dm.data_ptr()), flags=fl)
        for _ in range(3):
            local_step()
        P0.exL.sync()
        P0.exL.timer_start()
        for _ in range(mreps):
            local_step()
        lms = P0.exL.timer_stop() / mreps
        tm = torch.tensor([lms], dtype=torch.float64, device="cuda:%d" % dev)
        if dist is not None:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        lms = float(tm[0])
        match["local_map"] = {"what": "SearchByProjection(F, vpMapPoints, th=3, nnratio=0.8) per frame on the resident grid, device-resident",
                              "frames": B * world, "map_points_per_frame": float(tnq.mean()), "ms_per_batch": lms,
                              "frames_per_s": B * world / (lms * 1e-3), "map_points_per_s": float(tnq.sum()) * world / (lms * 1e-3),
                              "matches_frame0": int(dnm[0].item())}
        if world == 1 and not args.no_cpu_baseline:
            try:
                from oracle import oracle_match_py as om
                impl, kind = (om.reference(), "reference") if om.have_reference() else (om.oracle(), "port")
                scale_t = P0.exL.tables()["scale"]
                nfr = min(B, distinct)
                t0 = time.perf_counter()
                for i in range(nfr):
                    nC, n = int(nLh[i]), int(tnq[i])
                    impl.search_local_points(kLh[i, :nC], dLh[i, :nC], uRh[i, :nC], lk0[i, :nC], scale_t, gp, TQ[i, :n], TQD[i, :n], 3.0, 0.8)
                dtm = time.perf_counter() - t0
                match["local_map"]["cpu_baseline"] = {"frames_per_s": nfr / dtm, "cores": 1, "kind": kind,
                                                      "sample": "%d frames, grid build + SearchByProjection(local map), 1 host thread" % nfr}
            except Exception as e:  # context only
                match["local_map"]["cpu_baseline"] = {"error": str(e)}

    # ---- bag-of-words line: Frame::ComputeBoW on the device-resident left descriptors of the last
    #      batch against a synthetic vocabulary of ORBvoc.txt size (k = 10, L = 6: 10^6 words; the real file is not in the mount)
    bow = None
    if not args.no_match:
        voc_arrays = synth.synth_vocabulary_full(77, 10, 6)
        voc = capi.ORBVocabulary(voc_arrays, device=dev)

        def bow_step():
            capi.compute_bow(P0.exL, voc, 4, flags=AS, want=False)
        for _ in range(3):
            bow_step()
        P0.exL.sync()
        breps = 10
        P0.exL.timer_start()
        for _ in range(breps):
            bow_step()
        bms = P0.exL.timer_stop() / breps
        tb = torch.tensor([bms], dtype=torch.float64, device="cuda:%d" % dev)
        if dist is not None:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        bms = float(tb[0])
        nLh = P0.outL[0]
        g0 = capi.compute_bow(P0.exL, voc, 4)[0]
        bow = {"what": "Frame::ComputeBoW (transform, levelsup 4) per frame, vocabulary k=10 L=6 (1 111 111 nodes), device-resident",
               "frames": B * world, "features_per_frame": float(nLh.mean()), "ms_per_batch": bms, "frames_per_s": B * world / (bms * 1e-3),
               "features_per_s": float(nLh.sum()) * world / (bms * 1e-3), "words_frame0": int(len(g0["bow_word"])),
               "nodes_frame0": int(len(g0["fv_node"]))}
        if world == 1 and not args.no_cpu_baseline:
            try:
                from oracle import oracle_bow_py as ob
                orc = ob.OracleVocabulary(voc_arrays)      # the reference's text loader needs a 150 MB file for this size: port
                dLh = P0.outL[3]
                nfr = min(B, distinct)
                t0 = time.perf_counter()
                for i in range(nfr):
                    orc.transform(dLh[i, :int(nLh[i])], 4)
                dtb = time.perf_counter() - t0
                bow["cpu_baseline"] = {"frames_per_s": nfr / dtb, "cores": 1, "kind": "port",
                                       "sample": "%d frames, transform on 1 host thread (restatement equal to the reference's DBoW2)" % nfr}
            except Exception as e:  # context only
                bow["cpu_baseline"] = {"error": str(e)}
        # ORBmatcher::SearchByBoW (TrackReferenceKeyFrame): every frame against a keyframe made of the NEXT distinct frame's keypoints
        # and FeatureVector (both produced by this library: extraction + orb_compute_bow), i.e. another view of a similar scene
        try:
            fv_all = capi.compute_bow(P0.exL, voc, 4)
            dLh, kLh = P0.outL[3], P0.outL[2]
            kfs = []
            nd = min(B, distinct)
            for i in range(nd):
                j = (i + 1) % nd
                m = int(nLh[j])
                kfs.append(dict(desc=dLh[j, :m], angle=kLh[j, :m]["angle"], flags=np.ones(m, np.uint8),
                                fv={k: fv_all[j][k] for k in ("fv_node", "fv_off", "fv_feat")}))
            kfs = [kfs[i % len(kfs)] for i in range(B)]
            *arrs, kcap_kf = capi.pack_bow_keyframes(kfs)
            darrs = [torch.from_numpy(a.view(np.uint8).reshape(B, -1) if a.ndim > 1 else a).to("cuda:%d" % dev) for a in arrs]
            ksrc = tuple(t.data_ptr() for t in darrs) + (kcap_kf,)
            dm2 = torch.empty((B, P0.exL.kcap), dtype=torch.int32, device="cuda:%d" % dev)
            dnm2 = torch.empty(B, dtype=torch.int32, device="cuda:%d" % dev)
            fl2 = capi.ORB_SRC_DEVICE | capi.ORB_DST_DEVICE | AS

            def sbow_step():
                capi.search_by_bow(P0.exL, ksrc, 0.7, True, flags=fl2, out=(dnm2.data_ptr(), dm2.data_ptr()))
            for _ in range(3):
                sbow_step()
            P0.exL.sync()
            P0.exL.timer_start()
            for _ in range(breps):
                sbow_step()
            sms = P0.exL.timer_stop() / breps
            ts = torch.tensor([sms], dtype=torch.float64, device="cuda:%d" % dev)
            if dist is not None:
                dist.all_reduce(ts, op=dist.ReduceOp.MAX)
            sms = float(ts[0])
            bow["search_by_bow"] = {"what": "ORBmatcher::SearchByBoW(0.7, orientation check) per frame against one keyframe, device-resident",
                                    "frames": B * world, "ms_per_batch": sms, "frames_per_s": B * world / (sms * 1e-3),
                                    "matches_frame0": int(dnm2[0].item())}
            if world == 1 and not args.no_cpu_baseline:
                from oracle import oracle_match_py as om
                impl, kind = (om.reference(), "reference") if om.have_reference() else (om.oracle(), "port")
                nfr = min(B, distinct)
                t0 = time.perf_counter()
                for i in range(nfr):
                    m = int(nLh[i])
                    impl.search_by_bow(kfs[i]["desc"], kfs[i]["angle"], kfs[i]["flags"], kfs[i]["fv"], dLh[i, :m], kLh[i, :m]["angle"], fv_all[i], 0.7, True)
                bow["search_by_bow"]["cpu_baseline"] = {"frames_per_s": nfr / (time.perf_counter() - t0), "cores": 1, "kind": kind,
                                                        "sample": "%d frames, SearchByBoW on 1 host thread" % nfr}
        except Exception as e:  # context only
            bow["search_by_bow"] = {"error": repr(e)}
        voc.close()

    # ---- LocalMapping-side matchers (widening beyond SURVEY.md 8): the resident left frames act as keyframes. Timed through the
    #      public calls with HOST buffers (queries in, results out inside the timed region); CPU column = the reference's own lines
    #      (oracle/_ref/libmorb_ref_map.so) on one host core.
    mapping = None
    if not args.no_match:
        try:
            mapping = mapping_lines(args, capi, synth, torch, dist, P0, B, distinct, w, h, mbf, dev, world)
        except Exception as e:  # context only
            mapping = {"error": repr(e)}

    # ---- input rectification (System::TrackStereo's cv::remap, src/System.cc:254-261): the same left
    #      batch treated as raw frames and rectified on the device before the extraction; cost = the difference
    rectify = None
    if not args.no_match:
        mx, my = synth.rectify_maps(w, h, seed=1)
        P0.exL.set_rectify_maps(mx, my)

        def ext_step(fl):
            P0.exL.extract_batch((dL.data_ptr(), B, h, w), lap, out=P0.outL, flags=NO | AS | fl)
        tms = []
        for fl in (0, capi.ORB_INPUT_REMAP):
            for _ in range(3):
                ext_step(fl)
            P0.exL.sync()
            P0.exL.timer_start()
            for _ in range(10):
                ext_step(fl)
            tms.append(P0.exL.timer_stop() / 10)
        P0.exL.set_rectify_maps(None, None)
        tr = torch.tensor(tms, dtype=torch.float64, device="cuda:%d" % dev)
        if dist is not None:
            dist.all_reduce(tr, op=dist.ReduceOp.MAX)
        rms = max(float(tr[1] - tr[0]), 1e-6)
        rectify = {"what": "cv::remap INTER_LINEAR (float maps, radial-tangential rectification) of the raw frames into level 0, device-resident",
                   "images": B * world, "ms_per_batch": rms, "images_per_s": B * world / (rms * 1e-3),
                   "algorithmic_gbs": B * (2 * w * h) / (rms * 1e-3) / 1e9, "ms_extract_plain": float(tr[0]), "ms_extract_with_remap": float(tr[1])}
        if world == 1 and not args.no_cpu_baseline:
            try:
                import cv2
                cv2.setNumThreads(1)
                raw0 = np.ascontiguousarray(hostL[0])
                t0 = time.perf_counter()
                for _ in range(50):
                    cv2.remap(raw0, mx, my, cv2.INTER_LINEAR)
                rectify["cpu_baseline"] = {"images_per_s": 50 / (time.perf_counter() - t0), "cores": 1, "kind": "reference",
                                           "sample": "cv2.remap (the real OpenCV %s kernel the reference calls), 50 images, 1 thread" % cv2.__version__}
            except Exception as e:  # context only
                rectify["cpu_baseline"] = {"error": str(e)}

    # ---- single-pair latency through the C ABI (how Tracking calls the front-end: one stereo pair at a time, host image in,
    #      host keypoints / descriptors / mvuRight out); context next to the batched throughput
    latency = None
    if world == 1 and not args.no_match:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import latency as _lat
            r = _lat.main(100, quiet=True)
            latency = {"what": "one EuRoC stereo pair: operator() x 2 + ComputeStereoMatches, host buffers, batch 1",
                       "ms_median_sync_calls": r["sync"][0], "ms_p90_sync_calls": r["sync"][1],
                       "ms_median_async_calls": r["async"][0], "ms_p90_async_calls": r["async"][1]}
            for c in ("kitti", "tumvi"):   # the other image configurations, same measurement
                r = _lat.main(60, quiet=True, cfg=c)
                latency[c] = {"ms_median_sync_calls": r["sync"][0], "ms_p90_sync_calls": r["sync"][1],
                              "ms_median_async_calls": r["async"][0], "ms_p90_async_calls": r["async"][1]}
        except Exception as e:  # context only
            latency = {"error": str(e)}
        # the same pair through the C++ drop-in class (pageable cv::Mat in, std::vector<cv::KeyPoint> / cv::Mat out, the reference's
        # call pattern src/Frame.cc:194-217): sequential calls, and the two extractions on two std::threads like the reference
        try:
            import re
            import tempfile
            drv = os.path.join(ROOT, "tests", "cpp", "dropin_driver")
            cpp = os.path.join(ROOT, "morb_slam_b200", "cpp")
            subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I" + os.path.join(ROOT, "oracle", "shim"), "-I" + os.path.join(ROOT, "include"),
                            "-I" + cpp, os.path.join(ROOT, "tests", "cpp", "dropin_driver.cc"), os.path.join(cpp, "ORBextractor.cc"),
                            "-L" + os.path.join(ROOT, "morb_slam_b200", "lib"), "-lorb_b200",
                            "-Wl,-rpath," + os.path.join(ROOT, "morb_slam_b200", "lib"), "-o", drv], check=True, capture_output=True)
            with tempfile.TemporaryDirectory() as td:
                w_, h_, nf_, _lap, fx_, b_ = synth.CONFIGS["euroc"]
                Lp, Rp = synth.stereo_pair(9000, w_, h_)
                Lp.tofile(os.path.join(td, "l.raw")); Rp.tofile(os.path.join(td, "r.raw"))
                out = subprocess.run([drv, "--latency", str(w_), str(h_), str(nf_), os.path.join(td, "l.raw"), os.path.join(td, "r.raw"), "200",
                                      repr(float(np.float32(fx_ * b_))), repr(float(np.float32(fx_)))], check=True, capture_output=True, text=True).stdout
            m = {k: (float(a), float(b2)) for k, a, b2 in re.findall(r"(seq|threads|pair)\s+median ([0-9.]+) ms p90 ([0-9.]+) ms", out)}
            latency["dropin_cpp"] = {"what": "ORB_SLAM3::ORBextractor::operator() x 2 + ComputeStereoMatchesB200 of the drop-in class, pageable cv::Mat buffers",
                                     "ms_median_sequential": m["seq"][0], "ms_p90_sequential": m["seq"][1],
                                     "ms_median_two_threads": m["threads"][0], "ms_p90_two_threads": m["threads"][1],
                                     "ms_median_extract_pair": m["pair"][0], "ms_p90_extract_pair": m["pair"][1]}
        except Exception as e:  # context only
            if isinstance(latency, dict):
                latency["dropin_cpp"] = {"error": str(e)[:200]}

    # ---- the other image configurations of BASELINE.json at this GPU count (short runs, every rank takes part)
    workloads = None
    if CFG == "euroc" and not args.no_workloads:
        workloads = {c: quick_workload(c, capi, torch, dist, dev, rank, world, min(B, 128), 8, args.host_mem) for c in ("kitti", "tumvi")}

    # ---- reduce over ranks: max time
    t = torch.tensor([ms_resident, ms_e2e], dtype=torch.float64, device="cuda:%d" % dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_resident_max, ms_e2e_max = float(t[0]), float(t[1])
    frames = world * B * args.steps
    value = frames / (ms_resident_max * 1e-3)
    e2e_value = frames / (ms_e2e_max * 1e-3)

    line = None
    if rank == 0:
        peaks, peak_kind = _peaks()
        # algorithmic bytes per image: P = sum of level pixels, C = FAST candidates, K = keypoints
        Ppix = 0
        for l in range(8):
            wl, hl = P0.exL.level_size(l)
            Ppix += wl * hl
        names = ["pyramid", "blur", "fast_cells", "octree", "assemble", "orient_describe", "stereo_match", "stereo_gate"]
        # ONE definition of the algorithmic bytes per image, SURVEY.md 8(d): S1 pyramid P; S2 FAST P + 8 C; S3 quad-tree 8 C + 28 K;
        # S4 orientation 749 K and S5 descriptor 32 K (k_orient_describe); the P that S5 reads is the blur kernel's input (the blurred
        # levels are materialised here, their write is this design's choice and not counted). Sum = 3 P + 16 C + 809 K.
        algo = {"pyramid": Ppix, "blur": Ppix, "fast_cells": Ppix + 8 * cand, "octree": 8 * cand + 28 * K,
                "assemble": 0, "orient_describe": 749 * K + 32 * K, "stereo_match": 0, "stereo_gate": 0}
        dom = int(np.argmax(stage[:6]))
        dom_name = names[dom]
        # stage times above are for ONE image stream (left handle): B images per launch
        dom_bytes = algo[dom_name] * B
        achieved = dom_bytes / (stage[dom] * 1e-3) / 1e9 if stage[dom] > 0 else 0.0
        peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic = None
        kmap = {"pyramid": "k_resize_tiles", "blur": "k_blur7", "fast_cells": "k_fast_cells", "octree": "k_octree_passes",
                "assemble": "k_assemble", "orient_describe": "k_orient_describe"}
        tj = {}
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic_r2q.json")))
            traffic = float(tj[kmap[dom_name]]["dram_bytes_per_image"]) * B
        except Exception:
            pass
        per_kernel = {}
        for i, n in enumerate(names[:6]):
            if stage[i] > 0 and n in kmap:
                gbs = algo[n] * B / (stage[i] * 1e-3) / 1e9
                tk_ = tj.get(kmap[n], {})
                per_kernel[kmap[n]] = {"ms_per_launch_group": float(stage[i]), "achieved_gbs": gbs, "frac_hbm": gbs / peak,
                                       "algorithmic_bytes_per_image": float(algo[n]),
                                       # ncu --set full of this round's build (profiles/traffic_r2q.json): what actually bounds the kernel
                                       "issue_frac": None if tk_.get("issue_slots_busy_pct") is None else tk_["issue_slots_busy_pct"] / 100.0,
                                       "alu_pipe_frac": None if tk_.get("alu_pipe_pct") is None else tk_["alu_pipe_pct"] / 100.0,
                                       "dram_bytes_per_image_ncu": tk_.get("dram_bytes_per_image")}
                # the resource ncu shows as the kernel's binding one (the larger of issue slots and the busiest integer pipe), as a
                # fraction of ITS peak: the "ncu-measured roofline" of a kernel that is not memory-bound
                pk = per_kernel[kmap[n]]
                if pk["issue_frac"] is not None and pk["alu_pipe_frac"] is not None:
                    pk["binding"] = "alu_pipe" if pk["alu_pipe_frac"] >= pk["issue_frac"] else "issue_slots"
                    pk["frac_binding"] = max(pk["alu_pipe_frac"], pk["issue_frac"])
        roofline = {"bound": "hbm", "kernel": kmap.get(dom_name, "k_" + dom_name), "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_kind,
                    "note": "frac = algorithmic bytes (SURVEY.md 8(d)) / measured HBM peak; the image kernels are bound by the INT "
                            "ALU pipe / instruction issue on B200, not by HBM (per_kernel.issue_frac / alu_pipe_frac from the "
                            "committed ncu --set full page of this build, DRAM traffic = algorithmic bytes): profiles/README_r2.md",
                    "algorithmic_bytes_per_launch": dom_bytes, "ms_per_launch": float(stage[dom]),
                    "stage_ms_left_images": {n: float(v) for n, v in zip(names, stage)}, "per_kernel": per_kernel}
        cores = os.cpu_count() or 1
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            npairs = 64 * cores    # bounded sample: roughly 10-30 s of CPU work in total
            fps, kind, dt = cpu_reference_run(npairs, cores, w, h, nf, lap, fx, b)
            cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": "%d stereo pairs of the same workload on %d host threads (%.1f s)" % (npairs, cores, dt),
                   "note": "reference sources compiled unmodified against a scalar restatement of the OpenCV primitives "
                           "(no OpenCV C++ in this image); cv2_primitives_ms_per_image = the real OpenCV kernels alone, 1 thread",
                   "cv2_primitives_ms_per_image": cv2_primitives_ms(w, h, nf),
                   "real_opencv_estimate": real_opencv_estimate(fps, w, h, nf, lap)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_resident_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic",
                "config": {"workload": "%s-shape stereo %dx%d, %d features/image, ORB extract x2 + %s, "
                                       "batched frames (BASELINE.json configs[%d])" % (WORKLOAD_NAME[CFG], w, h, nf, MATCHER[CFG], CONFIG_INDEX[CFG]),
                           "stereo_pairs_per_step_per_gpu": B, "distinct_pairs": distinct,
                           "l2_policy": "inputs larger than L2 (%.0f MB of images + %.0f MB of pyramids per step)" % (
                               2 * B * w * h / 1e6, 2 * 2 * B * Ppix / 1e6),
                           "keypoints_per_image": K, "fast_candidates_per_image": float(cand),
                           "stereo_matches_frame0": n_matches},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(hostL.nbytes + hostR.nbytes),
                        "d2h_bytes_per_step": int(P0.d2h_bytes()), "ms_per_step": ms_e2e_max / args.steps, "pcie": pcie},
                "gpu_launches": int(launches),
                "roofline": roofline, "match": match, "bow": bow, "mapping": mapping, "rectify": rectify, "latency": latency,
                "workloads": workloads, "cpu_baseline": cpu}
        if knn:
            line["knn"] = knn
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="stereo pairs per step per GPU")
    ap.add_argument("--distinct", type=int, default=32, help="distinct synthetic pairs generated per rank")
    ap.add_argument("--knn-rows", type=int, default=1250000)
    ap.add_argument("--workload", default="euroc", choices=["euroc", "kitti", "tumvi"],
                    help="euroc = BASELINE.json configs[1] (the metric's configuration); kitti = configs[3] (1241x376, 2000 features); "
                         "tumvi = configs[2] (512x512 fisheye, 1500 features, lapping area, BF kNN + ratio instead of the row-band matcher)")
    ap.add_argument("--host-mem", default="pinned", choices=["pinned", "wc"],
                    help="input host buffers of the e2e leg: page-locked (default) or page-locked + write-combined")
    ap.add_argument("--numa-bind", action="store_true", help="bind every rank to the cores of its GPU's NUMA node")
    ap.add_argument("--no-workloads", action="store_true", help="skip the short KITTI / TUM-VI lines of the default run")
    ap.add_argument("--no-knn", action="store_true")
    ap.add_argument("--no-match", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    global CFG
    CFG = args.workload
    if CFG == "tumvi":
        args.no_match = True   # the next-row lines (grid search, bag of words, rectification) are quoted on the EuRoC workload
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_own_arm(args)


if __name__ == "__main__":
    main()
