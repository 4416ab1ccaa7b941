"""The bench lines committed under profiles/ (written by bench.py on the B200 box) carry every key of the measurement contract:
one JSON object per run with metric / value / unit / n_gpus / steps / warmup / ms_per_step / higher_is_better / scaling / vs_baseline /
dtype / data / config.workload, clocks without thermal slowdown, e2e with the host <-> device bytes, gpu_launches > 0, the roofline
of the dominant kernel and (N = 1) the CPU baseline; the reference arm's lines say so. CPU only: guards the format, not the numbers."""
import glob
import json
import os

import pytest

from tests.conftest import ROOT

P = os.path.join(ROOT, "profiles")


def _line(path):
    lines = [l for l in open(path).read().splitlines() if l.startswith("{")]
    assert len(lines) == 1, path
    return json.loads(lines[0])


OWN = sorted(f for f in glob.glob(os.path.join(P, "bench_r1j_*.json")) + glob.glob(os.path.join(P, "bench_r2*.json")) if "reference" not in f)
REF = sorted(glob.glob(os.path.join(P, "bench_r1j_*reference*.json")) + glob.glob(os.path.join(P, "bench_r2*reference*.json")))


@pytest.mark.parametrize("path", OWN, ids=[os.path.basename(f) for f in OWN])
def test_own_arm_line(path):
    d = _line(path)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert k in d, (path, k)
    assert d["metric"] == "stereo_frames_per_s_orb_extract_plus_stereo_match" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "u8" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "BASELINE.json configs[" in d["config"]["workload"] and "model" not in d["config"]
    assert d["steps"] >= 1 and d["warmup"] >= 3 and d["value"] > 0 and d["gpu_launches"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["clocks"]["sm_mhz"] >= 0.9 * d["clocks"]["sm_max_mhz"]
    e = d["e2e"]
    assert e["unit"] == "frames/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.02
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    if d["n_gpus"] == 1 and d.get("cpu_baseline"):
        c = d["cpu_baseline"]
        assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == "frames/s" and c["sample"]
    if d.get("knn"):
        assert d["knn"]["routes_equal"] is True


@pytest.mark.parametrize("path", REF, ids=[os.path.basename(f) for f in REF])
def test_reference_arm_line(path):
    d = _line(path)
    assert d["impl"] == "reference" and d["gpu_launches"] == 0 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


R2 = sorted(f for f in glob.glob(os.path.join(P, "bench_r2*.json")) if "reference" not in f)


@pytest.mark.parametrize("path", R2, ids=[os.path.basename(f) for f in R2])
def test_round2_line_extras(path):
    """round 2: issue / ALU-pipe fractions next to frac_hbm per kernel, the concurrent-copy ceiling of the e2e leg, and (EuRoC line)
    the KITTI / TUM-VI configurations measured at the same GPU count"""
    d = _line(path)
    pk = d["roofline"]["per_kernel"]
    assert "k_fast_cells" in pk and "k_octree_passes" in pk
    for k in ("k_fast_cells", "k_blur7", "k_resize_tiles", "k_orient_describe"):
        assert 0 < pk[k]["frac_hbm"] < 1.2 and pk[k]["algorithmic_bytes_per_image"] > 0
        assert pk[k]["issue_frac"] is None or 0 < pk[k]["issue_frac"] <= 1
    c = d["e2e"]["pcie"]["concurrent"]
    assert c["ceiling_frames_per_s"] > 0 and d["e2e"]["value"] <= 1.1 * c["ceiling_frames_per_s"]
    if "configs[1]" in d["config"]["workload"] and d.get("workloads"):
        for w in ("kitti", "tumvi"):
            assert d["workloads"][w]["value"] > 0 and d["workloads"][w]["e2e"]["value"] > 0 and d["workloads"][w]["n_gpus"] == d["n_gpus"]


def test_default_line_has_the_cpu_baseline():
    d = _line(os.path.join(P, "bench_r1j_default.json"))
    assert d["n_gpus"] == 1 and d["cpu_baseline"]["kind"] == "reference" and "configs[1]" in d["config"]["workload"]
