"""bench.py prints exactly one JSON line with the keys the driver reads — checked for the CPU reference arm here and for
the GPU arm (small batch, few steps) on the B200."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(args, timeout):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, f"expected ONE line on stdout, got {len(lines)}"
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "1"], 600)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "sliced face-detect images/sec" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["value"] > 0 and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.gpu
def test_gpu_arm_line():
    d = _run(["--steps", "3", "--warmup", "3", "--batch", "8", "--images", "64", "--cpu-sample", "1"], 1500)
    assert BASE_KEYS | {"roofline", "clocks", "gpu_launches"} <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["dtype"] == "f16" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["gpu_launches"] > 0 and d["vs_baseline"] is None and d["scaling"] == "weak"
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and 0 < r["frac"] < 1.2 and abs(r["achieved"] / r["peak"] - r["frac"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 8 * 768 * 1024 * 3 and e["d2h_bytes_per_step"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["value"] > 0 and c["cores"] >= 1 and "sample" in c
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    # round 2: parity gate before timing, the other north-star kernels in the same line, measured shares, extra legs
    assert d["parity_gate"] == "pass" and set(d["parity"]["checks"]) >= {"C2", "C1", "K3_greedynmm_ios_1024", "K3_nms_iou_9900", "K4_crop_stitch"}
    names = " ".join(k["kernel"] for k in r["other_kernels"])
    for part in ("k2_pose_decode", "k3_merge", "k2_finalize", "attach_keypoints", "k5_bias_act"):
        assert part in names
    hp = d["hot_path"]
    assert 0 < hp["backbone_share_of_step"] < 1.5 and hp["backbone_ms_per_step"] > 0 and hp["north_star_kernels_ms_per_step"] > 0
    x = d["extra"]
    assert x["f32"]["value"] > 0 and x["c1"]["value"] > 0 and x["c5"]["value"] > 0 and x["c1"]["cpu_port_seconds_per_image"] > 0
    assert {row["N"] for row in x["c3_merge_us"]["rows"]} == {256, 1024, 4096, 9900}
    assert 0 < x["c5"]["k4_crop"]["frac"] < 1.2 and 0 < x["c5"]["k4_stitch"]["frac"] < 1.2
